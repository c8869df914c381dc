python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in ngp ngpref refnerf nerf; do python bench.py --model $m --precision fp32 --no_extra --no_cpu_baseline --steps 10 --warmup 3 2>>gpurun_out/r07a.err | cut -c1-190; done
