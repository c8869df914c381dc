timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -3
python profiles/run_gemm.py 2>&1 | tail -7
