"""CPU oracle for the learn-nerf render-and-train hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU yard-stick.  The product path (``learn-nerf_b200/``) never imports it and
has no CPU fallback.

PARITY UNPINNED.  The reference (``/root/reference/learn_nerf``) is JAX/Flax;
``jax``/``flax``/``optax`` are not installed in this image and cannot be
installed (no network, not in the wheelhouse), and the reference holds no test,
golden vector or fixture for render.py / model.py / instant_ngp.py /
ref_nerf.py / train.py (its one test, ``learn_nerf/test_dataset.py``, covers
dataset shuffling only).  So this restatement could not be checked against the
reference's own outputs.  It follows the reference source line by line (each
function cites file:line) and is pinned only by self-derived known-answer
tests (``tests/test_oracle.py``) and by the committed fixtures in
``tests/golden/`` that this oracle itself generated.

Modules
-------
``expf``          deterministic fp32 ``exp`` (mul/add only) shared bit-for-bit
                  with the CUDA kernels, used where the renderer calls jnp.exp
``render_np``     numpy fp32, op-for-op: ray_t_range, stratified sampling,
                  termination probs, compositing, inverse-CDF fine sampling
``models_torch``  torch-CPU NeRFModel / InstantNGPModel / RefNERFModel with
                  Flax parameter names, differentiable
``render_torch``  the same renderer in torch (any dtype) for autograd oracles
``train_torch``   TrainLoop.losses + optax.adam restated
"""
