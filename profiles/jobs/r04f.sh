ncu --set full --clock-control none --import-source on -k regex:'tcg_tn|tcg_rows_kernel<\(int\)0' -s 3 -c 1 -o gpurun_out/r04_gemm_rows -f python profiles/run_gemm.py > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'tcg_tn' -s 3 -c 1 -o gpurun_out/r04_gemm_tn -f python profiles/run_gemm.py > gpurun_out/ncu_gemm2.log 2>&1
tail -n 1 gpurun_out/ncu_gemm.log | cut -c1-200; tail -n 1 gpurun_out/ncu_gemm2.log | cut -c1-200
