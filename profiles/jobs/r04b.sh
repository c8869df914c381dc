timeout 600 python -m pytest tests/test_gpu_nerf.py tests/test_gpu_fullsize.py -q -x -k "fp32 or not bf16" 2>&1 | tail -15
B="python bench.py --precision fp32 --steps 5 --warmup 3 --no_cpu_baseline --no_extra --no_cuda_graph"
$B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('tc  ', d['dtype'], d['ms_per_step'], d['value'])"
LNRF_FP32_FFMA=1 $B 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('ffma', d['dtype'], d['ms_per_step'], d['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r04_launches_nerf_train_fp32.csv $B > gpurun_out/ncu_fp32.log 2>&1
tail -n 2 gpurun_out/ncu_fp32.log
