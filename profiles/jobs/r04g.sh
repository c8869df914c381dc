timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_refnerf.py tests/test_gpu_ngpref.py -x -q 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_nerf.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -3
for mdl in nerf refnerf; do
B="python bench.py --model $mdl --precision fp32 --steps 5 --warmup 3 --no_cpu_baseline --no_extra --no_cuda_graph"
$B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('$mdl tc  ', d['dtype'], d['ms_per_step'], d['value'])"
done
