"""A few cooperative cases: python scripts/dev/coop_cases.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from miniframe_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), sys.argv[1])
sys.argv = sys.argv[:1]
import importlib.util
spec = importlib.util.spec_from_file_location("sb", os.path.join(os.path.dirname(__file__), "..", "small_batch.py"))
src = open(spec.origin).read().split("from miniframe_b200._engine import Engine\nrows = []")[0]
exec(src)
for N, B in [(500, 64), (2000, 64), (10000, 1), (10000, 8)]:
    ms, tf, ll, _ = run(N, B, reps=2)
    print(json.dumps({"N": N, "B": B, "ms": ms, "tflops": tf, "ll0": float(ll[0])}), flush=True)
