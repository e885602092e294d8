"""Writes the worker script of tests/test_gpu_sharded.py to a file so that it can be launched by hand (torchrun, optionally
under compute-sanitizer): python tools/run_sharded_worker.py <out_dir>  ->  <out_dir>/worker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_sharded import WORKER  # noqa: E402

out = os.path.abspath(sys.argv[1])
os.makedirs(out, exist_ok=True)
with open(os.path.join(out, "worker.py"), "w") as f:
    f.write(WORKER % (ROOT, out))
print(os.path.join(out, "worker.py"))
