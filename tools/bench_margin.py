#!/usr/bin/env python
"""tools/bench_margin.py [n ...] - throughput of svx_margin_scores (step 6.7) at collection sizes n (dim 1024, k 16, both
directions; whole call incl. the normalisation pass): ms per call, TFLOP/s of the two fp16 tcgen05 GEMMs, pairs/s."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speech_vecalign_b200 import score_align
torch.manual_seed(0)
for n in [int(a) for a in sys.argv[1:]] or [10000, 100000]:
    x = torch.randn(n, 1024, device="cuda", dtype=torch.float16)
    y = (x.float() + 0.5 * torch.randn(n, 1024, device="cuda")).half()
    s = score_align.compute_sim(x, y, 16, "ratio", as_numpy=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if n <= 200000 else 1
    e0.record()
    for _ in range(reps):
        s = score_align.compute_sim(x, y, 16, "ratio", as_numpy=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2 * 2.0 * n * n * 1024
    print(f"n={n}: {ms:.2f} ms per call, {flops / ms / 1e9:.1f} TFLOP/s (fp16 tcgen05), {n / ms * 1e3:.0f} pairs/s, score[0]={float(s[0]):.5f}", flush=True)
