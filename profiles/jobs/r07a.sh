python -m pytest tests/test_gpu_ngp.py -x -q -k "fp32_heads_train_path" -s 2>&1 | grep -v "^$" | tail -12
LNRF_FP32_FFMA=1 python -m pytest tests/test_gpu_ngp.py tests/test_gpu_ngpref.py -x -q 2>&1 | tail -3
