import sys, time, math, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import speech_vecalign_b200 as svb
from speech_vecalign_b200 import synth
n0s, n1s = synth.batch_sizes(64, seed=1234)
k = 5
types = svb.make_alignment_types(6)
host = []
for n0, n1 in zip(n0s, n1s):
    a = torch.randn((k, int(n0), 1024), dtype=torch.float32).pin_memory(); b = torch.randn((k, int(n1), 1024), dtype=torch.float32).pin_memory()
    host.append((a, b))
kw = dict(final_alignment_types=types, del_percentile_frac=0.2, width_over2=8, max_size_full_dp=300, costs_sample_size=20000, num_samps_for_norm=100, output="records", seeds=list(range(64)))
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = svb.vecalign_batch(host, **kw)
    torch.cuda.synchronize(); print("total %.1f ms" % (1e3 * (time.perf_counter() - t)))
pr = cProfile.Profile(); pr.enable()
out = svb.vecalign_batch(host, **kw)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
