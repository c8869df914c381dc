timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nerf.py tests/test_gpu_fullsize.py tests/test_gpu_refnerf.py -x -q 2>&1 | tail -4
python profiles/stage_bench.py 2>&1 | tail -12
