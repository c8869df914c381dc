timeout 600 python -m pytest tests/test_gpu_ngpref.py tests/test_gpu_refnerf.py tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -6
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "ngpref or ref" 2>&1 | tail -3
B="python bench.py --model ngpref --precision fp32 --steps 5 --warmup 3 --no_cpu_baseline --no_extra --no_cuda_graph"
$B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('ngpref tc  ', d['dtype'], d['ms_per_step'], d['value'])"
LNRF_FP32_FFMA=1 $B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('ngpref ffma', d['dtype'], d['ms_per_step'], d['value'])"
