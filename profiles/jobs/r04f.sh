python profiles/run_gemm.py 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:'tcg_' -s 9 -c 5 -o gpurun_out/r04_gemm_tc -f python profiles/run_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -n 2 gpurun_out/ncu_gemm.log | cut -c1-200
