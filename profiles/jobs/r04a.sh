set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/r04_bench1.json 2> gpurun_out/r04_bench1.err; tail -2 gpurun_out/r04_bench1.err
F="env SAVE=1 BWD=1 python profiles/run_fwd.py"
$F > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'cta2|dw_kernel' -s 2 -c 3 -o gpurun_out/r02_nerf_train_kernels -f $F > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log
