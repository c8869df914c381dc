"""The split-fp16 tcgen05 GEMM engine (csrc/gemm_tc.cu) against fp64 numpy: the fp32-accurate contraction
behind nn.Dense (reference model.py:51-60) and its transposes (what jax.grad derives, train.py:90)."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import *  # noqa: F401,F403  (sys.path setup)
from learn_nerf import _native

pytestmark = pytest.mark.gpu


def _call(mode, epi, M, N, A0, lda0, K0, A1, lda1, K1, B, ldb, C, ldc, bias=None, aux=None, ldaux=0, r1s=None,
          r1w=None, db=None, a_amax=None, b_amax=None, c_amax=None, mask_in=None, mask_out=None):
    _native.ensure_init(C.device)
    p = _native._p
    rc = _native.load().lnrf_tcgemm(mode, epi, M, N, p(A0), lda0, K0, p(A1), lda1, K1, p(B), ldb, p(C), ldc, p(bias),
                                    p(aux), ldaux, p(r1s), p(r1w), p(db), p(a_amax), p(b_amax), p(c_amax),
                                    p(mask_in), p(mask_out), _native._stream())
    _native._check(rc, "lnrf_tcgemm")


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("M,N,K0,K1,epi,btrans", [
    (1000, 256, 256, 0, 0, False),
    (40000 + 77, 256, 256, 0, 0, False),  # several tiles per CTA: ring slots and accumulator buffers wrap
    (30000, 256, 256, 60, 0, False),
    (50000, 128, 256, 24, 1, False),
    (25000 + 3, 256, 128, 0, 3, True),
    (33000, 60, 256, 0, 5, True),
    (4099, 256, 256, 60, 0, False),   # skip layer: two K segments, ragged last tile
    (777, 256, 60, 0, 0, False),      # first layer: one partial chunk
    (2048, 128, 256, 24, 1, False),   # colour layer shape
    (3000, 256, 128, 0, 3, True),     # g8 = dc @ W10^T + spre (x) w9
    (3000, 256, 256, 0, 2, True),     # dX chain with ReLU mask
    (500, 60, 256, 0, 5, True),       # d x_emb (N = 60)
    (129, 16, 64, 0, 1, False),       # Instant-NGP Ref-NeRF head shapes
    (129, 64, 16, 20, 0, False),
])
def test_rows_vs_fp64(M, N, K0, K1, epi, btrans):
    g = torch.Generator().manual_seed(M + N + K0)
    K = K0 + K1
    A0 = torch.randn(M, K0, generator=g).cuda()
    A1 = torch.randn(M, K1, generator=g).cuda() if K1 else None
    W = (torch.randn(K, N, generator=g) / np.sqrt(K)).cuda()
    Wd = W.t().contiguous() if btrans else W
    bias = torch.randn(N, generator=g).cuda()
    aux = torch.randn(M, N, generator=g).cuda()
    r1s, r1w = torch.randn(M, generator=g).cuda(), torch.randn(N, generator=g).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    c_amax = torch.zeros(1, device="cuda")
    _call(1 if btrans else 0, epi, M, N, A0, K0, K0, A1, K1, K1, Wd, Wd.shape[1], C, N, bias, aux, N, r1s, r1w,
          c_amax=c_amax)
    torch.cuda.synchronize()
    A = np.concatenate([A0.cpu().double().numpy()] + ([A1.cpu().double().numpy()] if K1 else []), axis=1)
    ref = A @ W.cpu().double().numpy()
    if epi in (0, 1):
        ref = ref + bias.cpu().double().numpy()
    if epi == 0:
        ref = np.maximum(ref, 0)
    if epi == 2:
        ref = ref * (aux.cpu().numpy() > 0)
    if epi == 3:
        ref = ref + np.outer(r1s.cpu().double().numpy(), r1w.cpu().double().numpy())
    out = C.cpu().double().numpy()
    assert np.isfinite(out).all()
    err = np.abs(out - ref).max()
    fp32 = np.abs((torch.cat([A0] + ([A1] if K1 else []), 1).cpu() @ W.cpu()).double().numpy()
                  - A @ W.cpu().double().numpy()).max()
    print(f"max abs err {err:.3e} (torch fp32 CPU matmul: {fp32:.3e})")
    assert err < max(4e-6, 1.5 * fp32), err   # |C| ~ 1: a few fp32 ulps, no worse than an fp32 matmul
    assert abs(float(c_amax) - np.abs(out).max()) <= 1e-6 * np.abs(out).max()


@pytest.mark.parametrize("scale", [1.0, 1e-7, 3e4])
def test_rows_scaled_operand(scale):
    """An operand far from O(1) (gradients) keeps fp32 accuracy through its amax-derived power-of-two scale."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 1500, 256, 256
    A = (torch.randn(M, K, generator=g) * scale).cuda()
    A[::7] *= 1e-3   # rows of very different magnitude
    W = (torch.randn(K, N, generator=g) / 16).cuda()
    amax = torch.zeros(1, device="cuda")
    _call(3, 0, A.numel(), 0, A, 0, 0, None, 0, 0, None, 0, amax, 0, c_amax=amax)
    assert float(amax) == float(A.abs().max())
    C = torch.empty(M, N, device="cuda")
    _call(0, 5, M, N, A, K, K, None, 0, 0, W, N, C, N, a_amax=amax)
    ref = A.cpu().double().numpy() @ W.cpu().double().numpy()
    rows = np.linalg.norm(ref, axis=1)
    err = np.linalg.norm(C.cpu().double().numpy() - ref, axis=1)
    big = rows > 1e-2 * rows.max()
    print("worst row rel err (large rows)", (err[big] / rows[big]).max(), "overall", _rel(C.cpu().double().numpy(), ref))
    assert (err[big] / rows[big]).max() < 2e-6
    assert _rel(C.cpu().double().numpy(), ref) < 1e-6


@pytest.mark.parametrize("Ksamp,M,N", [(5000, 256, 256), (4096 + 17, 60, 256), (3001, 256, 128), (777, 24, 128),
                                       (200, 64, 64), (100000, 256, 256),
                                       # M, N <= 64: the 192-sample-chunk variant (the Instant-NGP head shapes)
                                       (100001, 64, 64), (50000, 40, 64), (7777, 64, 16), (191, 32, 64),
                                       (193, 12, 64), (768, 64, 4),
                                       # M > 128 and an even number of 64-column blocks: the CTA-pair (cta_group::2) variant
                                       (7000, 192, 128), (300, 132, 256), (64, 256, 256), (31, 256, 128), (50001, 252, 244)])
def test_tn_vs_fp64(Ksamp, M, N):
    g = torch.Generator().manual_seed(Ksamp)
    H = torch.relu(torch.randn(Ksamp, M, generator=g)).cuda()
    G = (torch.randn(Ksamp, N, generator=g) * 1e-6).cuda()
    C0 = torch.randn(M, N, generator=g).cuda() * 1e-6
    C = C0.clone()
    db = torch.zeros(N, device="cuda")
    b_amax = torch.zeros(1, device="cuda")
    _call(3, 0, G.numel(), 0, G, 0, 0, None, 0, 0, None, 0, b_amax, 0, c_amax=b_amax)
    _call(2, 0, Ksamp, N, H, M, M, None, 0, 0, G, N, C, N, db=db, b_amax=b_amax)
    ref = C0.cpu().double().numpy() + H.cpu().double().numpy().T @ G.cpu().double().numpy()
    refdb = G.cpu().double().numpy().sum(0)
    r = _rel(C.cpu().double().numpy(), ref)
    f32 = _rel((C0.cpu() + H.cpu().t() @ G.cpu()).double().numpy(), ref)
    print(f"dW rel-L2 {r:.3e} (torch fp32 CPU: {f32:.3e}); db {_rel(db.cpu().double().numpy(), refdb):.3e}")
    assert r < 2e-6
    assert _rel(db.cpu().double().numpy(), refdb) < 1e-5


@pytest.mark.parametrize("M,N", [(1000, 256), (4099, 128), (77, 64), (45000, 256), (4130, 64), (300, 16)])
def test_relu_bit_masks_roundtrip(M, N):
    """epi 0 writes the bits [relu output > 0]; epi 6 applies them: same result as masking by the fp32 activations.
    The mask buffer is exactly tcg_mask_words long; a sentinel region behind it must stay untouched (the last tile
    of a ragged M spans row blocks that have no mask words)."""
    g = torch.Generator().manual_seed(M)
    K = 256
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(K, N, generator=g) / 16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    Hh = torch.empty(M, N, device="cuda")
    words = ((M + 31) // 32) * ((N + 31) // 32) * 32
    guarded = torch.full((words + 8192,), -1, dtype=torch.int32, device="cuda")
    bits = guarded[:words]
    _call(0, 0, M, N, A, K, K, None, 0, 0, W, N, Hh, N, bias, mask_out=bits)
    assert bool((guarded[words:] == -1).all()), "mask_out written past tcg_mask_words"
    G = torch.randn(M, N, generator=g).cuda()
    Wt = (torch.randn(N, N, generator=g) / 16).cuda()   # [N rows (outputs), K = N cols]
    C_bits = torch.empty(M, N, device="cuda")
    C_aux = torch.empty(M, N, device="cuda")
    _call(1, 6, M, N, G, N, N, None, 0, 0, Wt, N, C_bits, N, mask_in=bits)
    _call(1, 2, M, N, G, N, N, None, 0, 0, Wt, N, C_aux, N, aux=Hh, ldaux=N)
    assert torch.equal(C_bits, C_aux)
    assert (C_aux == 0).float().mean() > 0.3


def test_rows_strided_operands():
    """Leading dimensions larger than the logical widths (views into wider matrices) for A, W and C: the TMA
    tensor map of A, the weight conversion and the epilogue all honour them."""
    g = torch.Generator().manual_seed(11)
    M, N, K0, K1 = 2500, 192, 128, 60
    A0w = torch.randn(M, 384, generator=g).cuda()   # A0 = columns 64 .. 191 of a 384-wide matrix
    A1w = torch.randn(M, 64, generator=g).cuda()    # A1 = columns 0 .. 59 of a 64-wide matrix
    Ww = (torch.randn(K0 + K1, 256, generator=g) / 14).cuda()  # W = columns 0 .. 191 of a 256-wide matrix
    Cw = torch.full((M, 200), 7.0, device="cuda")
    bias = torch.randn(N, generator=g).cuda()
    A0 = A0w[:, 64:192]
    _native.ensure_init(Cw.device)
    p = _native._p
    a0_ptr = ctypes.c_void_p(A0w.data_ptr() + 64 * 4)
    rc = _native.load().lnrf_tcgemm(0, 1, M, N, a0_ptr, 384, K0, p(A1w), 64, K1, p(Ww), 256, p(Cw), 200, p(bias), None, 0,
                                    None, None, None, None, None, None, None, None, _native._stream())
    _native._check(rc, "lnrf_tcgemm")
    A = np.concatenate([A0.cpu().double().numpy(), A1w[:, :K1].cpu().double().numpy()], axis=1)
    ref = A @ Ww[:, :N].cpu().double().numpy() + bias.cpu().double().numpy()
    out = Cw.cpu().double().numpy()
    assert np.abs(out[:, :N] - ref).max() < 4e-6
    assert (out[:, N:] == 7.0).all()   # nothing written beyond the N columns
