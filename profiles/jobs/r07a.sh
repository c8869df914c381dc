python -m pytest tests/test_gpu_ngpref.py tests/test_gpu_ngp.py -x -q 2>&1 | tail -3
python bench.py --model ngpref --precision fp32 --no_extra --no_cpu_baseline --steps 10 --warmup 3 2>>gpurun_out/r07a.err | cut -c1-200
