#!/usr/bin/env python
"""Per-step time of N back-to-back style_transfer steps (batch 32 @512^2) for growing N, with the SM clock sampled
during each run: separates inter-kernel idle time (independent of N) from power-capped clocks (grows with N)."""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ccst_b200
from ccst_b200 import synth

dev = torch.device("cuda:0")
vgg, dec = synth.make_models(0)
eng = ccst_b200.engine_for(vgg, dec, dev)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
xs = [synth.images(32, 512, 512, 1 + i).to(dev) for i in range(2)]
g = torch.Generator().manual_seed(7)
stat = [torch.randn((1, 512, 1, 1), generator=g).abs().to(dev), (torch.rand((1, 512, 1, 1), generator=g) + 0.1).to(dev)]
out = torch.empty((32, 3, 512, 512), device=dev)
for i in range(5):
    eng.transfer(xs[i & 1], stat, 1.0, prec, out=out)
torch.cuda.synchronize()


class Clk(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.s = []

    def run(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not self.stop:
            self.s.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            time.sleep(0.002)


for n in (1, 2, 5, 10, 20, 50, 100, 300):
    time.sleep(1.0)  # let the chip cool to its idle state
    c = Clk()
    c.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        eng.transfer(xs[i & 1], stat, 1.0, prec, out=out)
    t_enq = time.perf_counter() - t0
    b.record()
    torch.cuda.synchronize()
    c.stop = True
    c.join()
    ms = a.elapsed_time(b)
    clk = sorted(x[0] for x in c.s)
    pw = max(x[1] for x in c.s)
    print(f"N={n:4d}: {ms / n:7.4f} ms/step  (host enqueue {t_enq / n * 1e3:6.3f} ms/step)  sm_mhz min/med/max "
          f"{clk[0]}/{clk[len(clk) // 2]}/{clk[-1]}  power max {pw:.0f} W")
