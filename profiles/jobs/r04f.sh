python profiles/run_gemm.py > gpurun_out/gemm_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'tcg_' -s 4 -c 16 -o gpurun_out/r04_gemm_tc -f python profiles/run_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -n 2 gpurun_out/ncu_gemm.log | cut -c1-200
