"""FP64 throughput probes on the B200: mma.sync shapes x occupancy, DFMA, cuBLAS DGEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from miniframe_b200._engine import Engine
eng = Engine.get()
names = {0: "mma m8n8k4", 1: "dfma", 2: "mma m16n8k4", 3: "mma m16n8k8", 4: "mma m16n8k16", 5: "m8n8k4 ilp1", 6: "m8n8k4 ilp2", 7: "m8n8k4 ilp4", 8: "m8n8k4 ilp8"}
occs = {0: "4x256/SM", 1: "1x128/SM", 2: "1x256/SM", 3: "2x256/SM"}
for occ in (1, 2, 3):
    for shape in (5, 6, 7, 8, 0):
        tf = eng.probe_fp64(occ * 10 + shape, 4000)
        print("%-14s %-9s %7.2f TFLOP/s" % (names[shape], occs[occ], tf), flush=True)
for n in (2048, 4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("cublas dgemm %d^3: %.2f TFLOP/s" % (n, 2 * n ** 3 / best / 1e9))
