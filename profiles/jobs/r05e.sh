timeout 600 python -m pytest tests/test_gpu_ngp.py -x -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -k "ngp" 2>&1 | tail -2
python bench.py --model ngp --precision bf16 --steps 10 --warmup 3 --no_cpu_baseline --no_extra 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('ngp bf16', d['ms_per_step'], d['value'])"
G="python bench.py --model ngp --precision bf16 --steps 2 --warmup 1 --no_extra --no_cpu_baseline --no_cuda_graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ngp_train_bf16.csv $G > gpurun_out/ncu4.log 2>&1
