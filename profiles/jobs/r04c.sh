timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_nerf.py -q -x -k "fp32 or not bf16" 2>&1 | tail -3
B="python bench.py --precision fp32 --steps 5 --warmup 3 --no_cpu_baseline --no_extra --no_cuda_graph"
$B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('tc  ', d['dtype'], d['ms_per_step'], d['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r04_launches_nerf_train_fp32.csv $B > gpurun_out/ncu_fp32.log 2>&1
tail -n 1 gpurun_out/ncu_fp32.log | cut -c1-200
