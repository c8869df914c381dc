G="python bench.py --model ngpref --precision fp32 --steps 2 --warmup 1 --no_extra --no_cpu_baseline --no_cuda_graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r08_launches_ngpref_train_fp32.csv $G > gpurun_out/ncu8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'jtv_bwd_pt_kernel|tcg_tn_kernel' -s 6 -c 4 -o gpurun_out/r08_ngpref_kernels -f $G > gpurun_out/ncu9.log 2>&1
tail -n 1 gpurun_out/ncu8.log | cut -c1-100; tail -n 1 gpurun_out/ncu9.log | cut -c1-100
