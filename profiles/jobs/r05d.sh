R="env SAVE=0 python profiles/run_fwd.py"
$R > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none -k regex:fwd_cta2 -s 2 -c 1 -o gpurun_out/r02_nerf_render_kernel -f $R > gpurun_out/ncu3.log 2>&1
G="python bench.py --model ngp --precision bf16 --steps 2 --warmup 1 --no_extra --no_cpu_baseline --no_cuda_graph"
$G > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ngp_train_bf16.csv $G > gpurun_out/ncu4.log 2>&1
N="env PREC=bf16 python profiles/run_ngp.py"
$N > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none -k regex:'ngp_|hashgrid' -s 8 -c 4 -o gpurun_out/r02_ngp_kernels -f $N > gpurun_out/ncu5.log 2>&1
for f in gpurun_out/ncu3.log gpurun_out/ncu4.log gpurun_out/ncu5.log; do tail -n 1 $f | cut -c1-160; done
