timeout 300 python -m pytest tests/test_gpu_ngp.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
python bench.py --model ngp --precision fp32 --no_extra --no_cpu_baseline --steps 10 --warmup 3 2>>gpurun_out/r09c.err | grep -o '"ms_per_step": [0-9.]*' | head -1
