"""Run pytest / bench.py in-process with lnrf_set_debug_flags(FLAGS) applied first (tuning helper).
usage: FLAGS=32 python profiles/with_flags.py pytest tests -m gpu -q -k bf16
       FLAGS=32 python profiles/with_flags.py bench --no_cpu_baseline"""
import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from learn_nerf import _native
_native.load().lnrf_set_debug_flags(int(os.environ.get("FLAGS", "0")))
what, rest = sys.argv[1], sys.argv[2:]
if what == "pytest":
    import pytest
    sys.exit(pytest.main(rest))
sys.argv = [os.path.join(ROOT, "bench.py")] + rest
runpy.run_path(sys.argv[0], run_name="__main__")
