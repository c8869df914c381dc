for k in 64 128 192 256; do echo "K=$k"; K=$k python profiles/run_gemm.py 2>&1 | grep "rows NN bias"; K=$k TCG_DBG=31 python profiles/run_gemm.py 2>&1 | grep "rows NN bias"; done
