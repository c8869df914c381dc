timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r05_bench1.json 2> gpurun_out/r05_bench1.err; tail -2 gpurun_out/r05_bench1.err
B="python bench.py --precision fp32 --steps 2 --warmup 1 --no_cpu_baseline --no_extra --no_cuda_graph"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_nerf_train_fp32.csv $B > gpurun_out/ncu_fp32.log 2>&1
R="python bench.py --model refnerf --precision fp32 --steps 2 --warmup 1 --no_cpu_baseline --no_extra --no_cuda_graph"
$R > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_refnerf_train_fp32.csv $R > gpurun_out/ncu_ref.log 2>&1
tail -n 1 gpurun_out/ncu_fp32.log | cut -c1-150; tail -n 1 gpurun_out/ncu_ref.log | cut -c1-150
