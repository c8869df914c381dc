"""B200-native drop-in for the hot path of unixpickle/learn-nerf.

Same module and class names as the reference package (``learn_nerf.render``,
``learn_nerf.model``, ``learn_nerf.instant_ngp``, ``learn_nerf.train``) so that the
reference's ``train_nerf.py`` / ``render_nerf.py`` flows port by switching imports;
arrays are torch CUDA tensors and all arithmetic runs in liblnrf.so (hand-written
sm_100a CUDA behind the C ABI in include/lnrf.h).
"""
