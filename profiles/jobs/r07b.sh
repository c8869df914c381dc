G="python bench.py --model ngp --precision fp32 --steps 2 --warmup 1 --no_extra --no_cpu_baseline --no_cuda_graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r07_launches_ngp_train_fp32.csv $G > gpurun_out/ncu7.log 2>&1
G="python bench.py --model ngpref --precision fp32 --steps 2 --warmup 1 --no_extra --no_cpu_baseline --no_cuda_graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r07_launches_ngpref_train_fp32.csv $G > gpurun_out/ncu8.log 2>&1
tail -n 1 gpurun_out/ncu7.log | cut -c1-100; tail -n 1 gpurun_out/ncu8.log | cut -c1-100
