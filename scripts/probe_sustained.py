"""How long can the FP64 tensor pipe run at its burst rate?  The register-resident DMMA probe
(mf_probe_fp64 mode 0, and 30 = two 256-thread CTAs per SM) for launches of increasing length,
with nvidia-smi clocks / power sampled during the long ones."""
import os, sys, subprocess, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from miniframe_b200._engine import Engine
eng = Engine.get()
for mode in (0, 30):
    for iters in (2000, 4000, 8000, 16000, 32000, 64000, 128000, 256000, 512000, 1024000):
        path = tempfile.mktemp()
        f = open(path, "w")
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                                "--format=csv,noheader,nounits", "-lms", "50"], stdout=f) if iters >= 256000 else None
        if smi: time.sleep(1.2)
        tf = eng.probe_fp64(mode, iters)
        line = ""
        if smi:
            smi.terminate(); smi.wait(); f.close()
            rows = [r.strip().split(", ") for r in open(path) if r.strip()]
            busy = [r for r in rows if float(r[1]) > 250]
            if busy:
                flags = sorted(set(x for row in busy for x in row[2:5]))
                line = "  clocks min/max %s/%s MHz, power max %.0f W, temp max %s C, flags %s" % (
                    min(row[0] for row in busy), max(row[0] for row in busy), max(float(row[1]) for row in busy),
                    max(row[5] for row in busy), flags)
        print("mode %2d iters %7d: %6.2f TFLOP/s%s" % (mode, iters, tf, line), flush=True)
