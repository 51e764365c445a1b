#!/usr/bin/env python
"""bench.py — throughput of the B200 alignment path on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload default|cfg1|cfg2|cfg3|cfg4|cfg5|all] [--pairs P]
    python bench.py --impl reference ...       # the reference's CPU implementation, all host cores

A step = one pass of the hot path (normalise -> downsample -> norms -> knob -> dense level ->
banded levels -> traceback) over one batch of synthetic document pairs.

default  BASELINE.json configs[1] (2000 x 2000 segments, dim 1024, max overlap 4: K=4, a=5) as the headline line -
         a batch of --pairs independent pairs per GPU, weak scaling (pairs are closed computations: no data-path
         collective) - plus `all_configs`: short runs of configs[0] (the shipped example pair),
         configs[2] (20000 x 20000), configs[3] (ONE corpus of 8192 pairs with 200-800 segments, a=6, LPT-partitioned
         over the ranks: strong scaling, results gathered on rank 0) and configs[4] (5000 x 5000, a=8).
cfgN     that configuration alone as the headline line.

Printed (rank 0, ONE JSON line): value = whole-job aligned pairs/s with the embeddings already in HBM (CUDA events,
max over ranks); e2e = the same through the public API (speech_vecalign_b200.vecalign_batch) from pinned HOST
tensors, H2D and the D2H of the packed results inside the timed region, next to a probe of the box's aggregate
pinned host->device bandwidth; roofline of the longest launcher and every launcher's numbers under `kernels`;
cpu_baseline = the reference's CPU path (oracle/_ref core when present, else the oracle port) on a bounded sample.
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, a, fixed (n0,n1) or None, default pairs per GPU (cfg4: pairs of the whole corpus))
    "cfg1": ("BASELINE configs[0]: example/voxpopuli en-de document pair (237 x 217 segments), alignment_max_size=4", 4, (237, 217), 64),
    "cfg2": ("BASELINE configs[1]: synthetic pairs 2000x2000 segments, dim 1024, max overlap 4 (a=5)", 5, (2000, 2000), 256),
    "cfg3": ("BASELINE configs[2]: synthetic long-session pairs 20000x20000, dim 1024, a=5, search_buffer_size=5", 5, (20000, 20000), 32),
    "cfg4": ("BASELINE configs[3]: ONE corpus of synthetic doc pairs with 200-800 segments each, a=6, LPT-partitioned over the GPUs", 6, None, 8192),
    "cfg5": ("BASELINE configs[4]: synthetic pairs 5000x5000, dim 1024, alignment_max_size=8", 8, (5000, 5000), 64),
}
DIM = 1024
PARAMS = dict(del_percentile_frac=0.2, search_buffer_size=5, max_size_full_dp=300, costs_sample_size=20000,
              num_samps_for_norm=100)
CFG4_CHUNK = 1024          # pairs per launch chain of the config-4 corpus
CFG4_E2E_PAIRS = 512       # pairs of the corpus the end-to-end arm streams from pinned host memory

# what bounds each launcher: the cost kernels are FP32-pipe / shared-memory bound (arithmetic
# intensity 17-36 flop/B; exact mode = separately rounded multiply and add), the DPs are serial latency chains
BOUND = {"svx_level_prologue": "hbm", "svx_normalize_rows": "hbm", "svx_downsample": "hbm", "svx_sample_norms": "hbm",
         "svx_score_pairs": "l2-gather", "svx_del_knob": "latency", "svx_dense_costs": "fp32", "svx_dense_dp": "latency",
         "svx_banded_costs_level0": "fp32", "svx_banded_costs_coarse": "smem", "svx_banded_dp_level0": "latency",
         "svx_banded_dp_coarse": "latency", "svx_widen_fp16": "hbm"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="default", choices=["default", "all"] + sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="document pairs per GPU per step (cfg4: pairs of the corpus; 0 = workload default)")
    ap.add_argument("--cost-mode", default="exact", choices=["exact", "fast", "tc"])
    ap.add_argument("--streams", type=int, default=0,
                    help="pair groups on separate CUDA streams (1 = one serial chain; 0 = auto: 4 below 128 pairs, else 1 - "
                         "large batches already amortise the latency-bound wavefront kernels)")
    ap.add_argument("--unfused-prologue", action="store_true", help="A/B: the three separate prologue launchers")
    ap.add_argument("--widen-pass", action="store_true",
                    help="A/B (cfg4): widen the resident fp16 rows to fp32 in a separate pass instead of reading them from the prologue")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="default workload without the all_configs block")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = max(steps, 10)")
    ap.add_argument("--h2d-probe", action="store_true", help="only probe the aggregate pinned host->device bandwidth of the ranks")
    return ap.parse_args()


def workload_sizes(name, pairs, rank):
    from speech_vecalign_b200 import synth
    _, a, fixed, _ = WORKLOADS[name]
    if fixed is not None:
        return np.full(pairs, fixed[0], dtype=np.int64), np.full(pairs, fixed[1], dtype=np.int64), a
    n0, n1 = synth.batch_sizes(pairs, seed=1234 + rank)
    return n0, n1, a


def example_pair(a):
    """BASELINE configs[0]: the overlap tensors of the shipped en-de pair (tests/golden/example, copied from the
    reference's example/voxpopuli by tests/golden/make_golden.py), built as align() builds them."""
    from speech_vecalign_b200 import embedding_utils as eu
    from speech_vecalign_b200.vecalign import load_ignore_index_file
    ex = os.path.join(ROOT, "tests", "golden", "example")
    out = []
    for lang, side in (("en", "src"), ("de", "tgt")):
        key_to_row, rows = eu.read_in_embeddings(f"{ex}/{lang}.cat_segs.txt", f"{ex}/{lang}.embed", True, False)
        lines = open(f"{ex}/{lang}.segments.txt", "rt", encoding="utf-8").readlines()
        ign = load_ignore_index_file(f"{ex}/ignore.{side}.txt")
        out.append(np.ascontiguousarray(eu.make_doc_embedding(key_to_row, rows, lines, a - 1, ignore_indices=ign, overlap_segments=True)))
    return out[0], out[1]


# ------------------------------------------------------------------------------------------------
# CPU reference arm (no CUDA anywhere on this path)
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(name, a, n0, n1, seed0, counter):
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)    # inf -> float32 casts inside the reference core
    from threadpoolctl import threadpool_limits
    from oracle import ref_loader, vecalign_oracle as vo
    from speech_vecalign_b200 import synth
    with counter.get_lock():
        wid = counter.value
        counter.value += 1
    _W["limits"] = threadpool_limits(1)
    core = ref_loader.ref_core()
    _W["core"], _W["kind"] = (core, "reference") if core is not None else (None, "port")
    _W["vo"] = vo
    i = wid % len(n0)
    k = a - 1
    _W["pair"] = example_pair(a) if name == "cfg1" else synth.synth_pair(int(n0[i]), int(n1[i]), k, dim=DIM, seed=seed0 + wid)
    _W["args"] = (vo.alignment_types(a), PARAMS["del_percentile_frac"], math.ceil(k / 2) + PARAMS["search_buffer_size"],
                  PARAMS["max_size_full_dp"], PARAMS["costs_sample_size"], PARAMS["num_samps_for_norm"])
    _W["a"] = a


def _cpu_worker_step(step):
    vo = _W["vo"]
    v0, v1 = _W["pair"]
    np.random.seed(step)
    t = time.perf_counter()
    st = vo.vecalign(v0.copy(), v1.copy(), *_W["args"], core=_W["core"])
    dt = time.perf_counter() - t
    w = _W["args"][2]
    return dt, vo.dp_cells(v0.shape[1], v1.shape[1], w, PARAMS["max_size_full_dp"]), len(st[0]["final_alignments"]), _W["kind"]


class CpuReference:
    """The reference's CPU implementation of the path on all host cores: one worker process per
    core (the reference itself is a serial loop, seg_align/align.py:206; pairs are independent, so
    this is its fair multi-core form), OpenBLAS pinned to one thread per worker.  One step = every
    worker aligns one pair of the workload."""

    def __init__(self, name, nproc=None):
        import multiprocessing as mp
        from oracle import core as ocore
        ocore.build()
        self.nproc = nproc or len(os.sched_getaffinity(0))
        _, a, _, _ = WORKLOADS[name]
        n0, n1, a = workload_sizes(name, max(self.nproc, 8), 0)
        ctx = mp.get_context("fork")
        counter = ctx.Value("i", 0)
        self.pool = ctx.Pool(self.nproc, initializer=_cpu_worker_init, initargs=(name, a, n0, n1, 777, counter))
        self.name = name

    def step(self, i):
        t = time.perf_counter()
        res = self.pool.map(_cpu_worker_step, [i] * self.nproc, chunksize=1)
        wall = time.perf_counter() - t
        return wall, sum(r[1] for r in res), res[0][3], sum(r[0] for r in res) / len(res)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = "cfg2" if args.workload in ("default", "all") else args.workload
    desc, a, _, _ = WORKLOADS[name]
    ref = CpuReference(name)
    for i in range(args.warmup):
        ref.step(i)
    t_tot, cells, kind, per_pair = 0.0, 0, "port", 0.0
    for i in range(args.steps):
        w, c, kind, pp = ref.step(args.warmup + i)
        t_tot += w
        cells += c
        per_pair += pp
    ref.close()
    value = ref.nproc * args.steps / t_tot
    sample = f"{ref.nproc} worker processes x 1 pair of the workload per step ({args.steps} timed steps), OpenBLAS 1 thread/worker; " \
             f"mean single-core time per pair {per_pair / args.steps:.2f} s"
    line = {
        "impl": "reference", "metric": "aligned doc pairs/sec (DP cells/sec alongside)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "pairs_per_step": ref.nproc, "alignment_max_size": a, **PARAMS},
        "dp_cells_per_sec": cells / t_tot,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": ref.nproc, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class Ctx:
    """Process-wide state of the GPU arm: ranks, device, peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG", "WARN")          # a caller's NCCL_DEBUG=INFO is kept; NCCL's output goes to stderr (StdoutToStderr)
            dist.init_process_group("nccl", device_id=self.dev)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak, self.peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else \
            (6650.0, "fallback (B200_PROFILING.md)")
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            self.traffic = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        if self.world == 1:
            return [int(x) for x in xs]
        t = self.torch.tensor([int(x) for x in xs], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return [int(v) for v in t.tolist()]


def h2d_probe(ctx, gb=4.0, reps=3):
    """Aggregate pinned host -> device bandwidth of the ranks: every rank streams `gb` GB in 256 MB cudaMemcpyAsync
    pieces at the same time; GB/s = all ranks' bytes / the slowest rank's time.  The ceiling the end-to-end arm
    (pinned host inputs) can reach on this box, whatever the kernels do."""
    torch = ctx.torch
    n = int(gb * 2 ** 30)
    piece = 256 << 20
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.zero_()
    devb = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    best = None
    for _ in range(reps + 1):
        ctx.barrier()
        t0 = time.perf_counter()
        for o in range(0, n, piece):
            devb[o:o + piece].copy_(host[o:o + piece], non_blocking=True)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)      # first repetition warms up
    del host, devb
    return {"gbs_aggregate": ctx.world * n / best / 1e9, "gbs_per_gpu": n / best / 1e9, "bytes_per_rank": n,
            "how": f"{ctx.world} ranks x {gb:.0f} GB pinned -> device in 256 MB cudaMemcpyAsync pieces, concurrently; best of {reps}"}


def launcher_report(ctx, ktimes, alg, flops, workload, cost_mode, sm_mhz):
    """kernels = every launcher's {ms, GB/s, fraction of the HBM peak, limiter, TFLOP/s}; roofline = the LONGEST one."""
    fp32_peak = 148 * 128 * sm_mhz * 1e6 * (1 if cost_mode == "exact" else 2) / 1e12     # TFLOP/s; exact mode: multiply and add are rounded separately
    serial_ms = float(sum(ktimes.values())) or 1e-9
    kernels = {}
    for nm, ms in sorted(ktimes.items(), key=lambda kv: -kv[1]):
        gbps = (alg.get(nm, 0) / (ms * 1e-3) / 1e9) if ms > 0 else None
        lim = BOUND.get(nm)
        if nm == "svx_dense_costs" and cost_mode == "tc":
            lim = "tensor"
        kernels[nm] = {"ms_per_step": ms, "GBps": gbps, "hbm_frac": (gbps / ctx.hbm_peak) if gbps else None, "limiter": lim,
                       "share_of_step": ms / serial_ms}
        if nm in flops and ms > 0:
            tf = flops[nm] / (ms * 1e-3) / 1e12
            kernels[nm]["fp32_tflops"] = tf
            kernels[nm]["fp32_frac"] = tf / fp32_peak
    dom = max(ktimes, key=ktimes.get)
    dom_ms = ktimes[dom]
    achieved = alg.get(dom, 0) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s",
                "frac": achieved / ctx.hbm_peak, "traffic": ctx.traffic.get(workload, {}).get(dom), "peak_source": ctx.peak_src,
                "algorithmic_bytes_per_launch": alg.get(dom, 0), "ms_per_launch": dom_ms,
                "share_of_step": dom_ms / serial_ms, "limiter": kernels[dom]["limiter"],
                "timing": "CUDA events around every launcher, extra passes of the same step as one serial chain"}
    if dom in flops and dom_ms > 0:
        tf = flops[dom] / (dom_ms * 1e-3) / 1e12
        roofline["fp32"] = {"achieved_tflops": tf, "peak_tflops": fp32_peak, "frac": tf / fp32_peak,
                            "peak_source": f"148 SMs x 128 lanes x {sm_mhz:.0f} MHz, " +
                                           ("1 flop per lane and cycle (exact mode: separately rounded multiply and add)" if cost_mode == "exact" else "FMA")}
    # the launcher is FP32-bound, not HBM-bound: the largest HBM-bound launcher is reported beside it
    hbm_doms = [k for k in ktimes if BOUND.get(k) == "hbm"]
    if BOUND.get(dom) != "hbm" and hbm_doms:
        hk = max(hbm_doms, key=ktimes.get)
        hg = alg.get(hk, 0) / (ktimes[hk] * 1e-3) / 1e9
        roofline["largest_hbm_bound_launcher"] = {"kernel": hk, "bound": "hbm", "achieved": hg, "peak": ctx.hbm_peak, "unit": "GB/s",
                                                  "frac": hg / ctx.hbm_peak, "traffic": ctx.traffic.get(workload, {}).get(hk),
                                                  "algorithmic_bytes_per_launch": alg.get(hk, 0), "ms_per_launch": ktimes[hk],
                                                  "share_of_step": ktimes[hk] / serial_ms}
    return kernels, roofline, serial_ms


def oracle_check(v0, v1, types, w, seed, recs, global_stream_seed=None):
    """Alignments / scores of one pair against the CPU oracle (outside every timed region)."""
    try:
        from oracle import vecalign_oracle as vo
        from speech_vecalign_b200.engine import records_to_alignments
        np.random.seed(seed if global_stream_seed is None else global_stream_seed)
        ref = vo.vecalign(v0, v1, types, PARAMS["del_percentile_frac"], w, PARAMS["max_size_full_dp"],
                          PARAMS["costs_sample_size"], PARAMS["num_samps_for_norm"], fast_host=True)
        al, sc = records_to_alignments(recs)
        same = al == [(list(x), list(y)) for x, y in ref[0]["final_alignments"]]
        diff = float(np.max(np.abs(sc - ref[0]["alignment_scores"]))) if len(sc) == len(ref[0]["alignment_scores"]) and len(sc) else None
        return same, diff
    except Exception as e:  # the checker is optional here
        return None, repr(e)


def bench_weak(ctx, name, pairs, steps, warmup, e2e_steps, with_cpu, with_clocks):
    """One workload with a fixed (n0, n1) per pair (or the shipped example pair), `pairs` pairs per GPU per step."""
    torch = ctx.torch
    import speech_vecalign_b200 as svb
    from speech_vecalign_b200 import capi, synth
    from speech_vecalign_b200.engine import BatchRun
    args, rank, world, dev = ctx.args, ctx.rank, ctx.world, ctx.dev
    lib = capi.lib()
    desc, a, _, _ = WORKLOADS[name]
    streams = args.streams if args.streams > 0 else (4 if pairs < 128 else 1)
    n0, n1, a = workload_sizes(name, pairs, rank)
    k = a - 1
    types = svb.make_alignment_types(a)
    w = math.ceil(k / 2) + PARAMS["search_buffer_size"]
    mode = {"exact": capi.SVX_COST_EXACT, "fast": capi.SVX_COST_FAST, "tc": capi.SVX_COST_TC}[args.cost_mode]

    # ---- inputs born in HBM: a pristine copy and the working copy the path mutates ----
    off0 = np.concatenate([[0], np.cumsum(k * (n0 + n1) * DIM)]).astype(np.int64)
    total = int(off0[-1])
    pristine = torch.empty(total, dtype=torch.float32, device=dev)
    work = torch.empty_like(pristine)

    def views(buf, cnt=pairs):
        out = []
        for p in range(cnt):
            b = int(off0[p])
            m0 = k * int(n0[p]) * DIM
            out.append((buf[b:b + m0].view(k, int(n0[p]), DIM), buf[b + m0:int(off0[p + 1])].view(k, int(n1[p]), DIM)))
        return out

    pv, wv = views(pristine), views(work)
    note(f"{name}: generating {pairs} pairs per GPU")
    if name == "cfg1":
        e0, e1 = example_pair(a)
        t0, t1 = torch.from_numpy(e0).to(dev), torch.from_numpy(e1).to(dev)
        for p in range(pairs):
            pv[p][0].copy_(t0)
            pv[p][1].copy_(t1)
        data = "the shipped example pair (tests/golden/example), replicated"
    else:
        for p in range(pairs):
            synth.synth_pair_torch(int(n0[p]), int(n1[p]), k, dim=DIM, seed=100000 * rank + p, device=dev,
                                   out0=pv[p][0], out1=pv[p][1])
        data = "synthetic"
    torch.cuda.synchronize()

    np.random.seed(4242 + rank)
    run = BatchRun([t0.data_ptr() for t0, _ in wv], [t1.data_ptr() for _, t1 in wv], n0, n1, k, k, DIM, types,
                   PARAMS["del_percentile_frac"], w, PARAMS["max_size_full_dp"], PARAMS["costs_sample_size"],
                   PARAMS["num_samps_for_norm"], dev, cost_mode=mode, fused_prologue=not args.unfused_prologue)
    cells = run.dp_cells()

    def one_step(timing, ngroups):
        work.copy_(pristine)                       # untimed: the path normalises its input in place
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run.run(timing=timing, ngroups=ngroups)
        e1.record()
        return e0, e1

    note(f"{name}: planned, timing {warmup} + {steps} steps")
    for _ in range(warmup):
        one_step(False, streams)
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if (with_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    lib.svx_launch_count(1)
    t_wall = time.perf_counter()
    evs = [one_step(False, streams) for _ in range(steps)]
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(lib.svx_launch_count(0))
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    # per-launcher device times: the same step again as ONE serial chain (one stream), so that a
    # kernel's duration is not inflated by the other groups' kernels running beside it
    ktimes, kpasses = {}, min(steps, 3)
    for _ in range(kpasses):
        one_step(True, 1)
        for nm, ms in run.kernel_times().items():        # synchronises
            ktimes[nm] = ktimes.get(nm, 0.0) + ms / kpasses
    total_ms = ctx.max_over_ranks(sum(step_ms))
    res = run.results()
    n_align = sum(r["nrecs"] for r in res)
    bad = sum(1 for r in res if r["status"])

    # ---- e2e: public API, pinned host tensors in, packed records out -----------------------------
    e2e = None
    note(f"{name}: value done, e2e arm ({e2e_steps} steps)")
    if e2e_steps > 0:
        # the end-to-end arm streams batches of at most 64 pairs / ~4.3 GB of pinned host memory per rank (the
        # arm is PCIe-bound, so the batch size does not change pairs/s)
        ep = min(pairs, 64)
        while ep > 1 and int(off0[ep]) * 4 > 4.3e9:           # long documents: keep the pinned buffer near 4 GB
            ep -= 1
        etotal = int(off0[ep])
        host = torch.empty(etotal, dtype=torch.float32, pin_memory=True)
        host.copy_(pristine[:etotal])
        hv = views(host, ep)
        kw = dict(final_alignment_types=types, del_percentile_frac=PARAMS["del_percentile_frac"], width_over2=w,
                  max_size_full_dp=PARAMS["max_size_full_dp"], costs_sample_size=PARAMS["costs_sample_size"],
                  num_samps_for_norm=PARAMS["num_samps_for_norm"], cost_mode=args.cost_mode, output="records",
                  streams=streams, seeds=[1000003 * rank + p for p in range(ep)])
        out = svb.vecalign_batch(hv, **kw)         # warm-up (allocator, page-locking paths)
        d2h = sum(o["recs"].nbytes + 8 * len(o["del_penalty"]) + 8 for o in out)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = svb.vecalign_batch(hv, **kw)
        ctx.barrier()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ep * e2e_steps / dt, "unit": "pairs/s",
               "h2d_bytes_per_step": int(etotal * 4 + (run.host_init_bytes * ep) // pairs), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps, "pairs_per_gpu_per_step": ep,
               "h2d_gbs_achieved_aggregate": world * etotal * 4 * e2e_steps / dt / 1e9,
               "api": "speech_vecalign_b200.vecalign_batch(pinned host fp32 tensors, seeds=per pair, output='records')"}
        # the same arm with the embeddings in fp16, the dtype of the reference's .embed files (--fp16_embed / stopes):
        # half the PCIe bytes, widened on the device.  Reported beside e2e, not instead of it.
        host16 = torch.empty(etotal, dtype=torch.float16, pin_memory=True)
        host16.copy_(pristine[:etotal])
        hv16 = views(host16, ep)
        out16 = svb.vecalign_batch(hv16, **kw)
        ctx.barrier()
        t0 = time.perf_counter()
        s16 = max(3, e2e_steps // 2)
        for _ in range(s16):
            out16 = svb.vecalign_batch(hv16, **kw)
        ctx.barrier()
        dt16 = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e["fp16_inputs"] = {"value": world * ep * s16 / dt16, "unit": "pairs/s", "h2d_bytes_per_step": int(etotal * 2),
                              "ms_per_step": 1e3 * dt16 / s16, "records": int(sum(o["nrecs"] for o in out16))}
        del host, host16, hv, hv16

    clocks = sampler.stop() if sampler else None      # sampled over the timed loop, the kernel passes and e2e
    n_align, bad = ctx.sum_over_ranks([n_align, bad])
    alg, flops = run.algorithmic_bytes(), run.cost_flops()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    kernels, roofline, serial_ms = launcher_report(ctx, ktimes, alg, flops, name, args.cost_mode, sm_mhz)

    parity = None
    note(f"{name}: oracle spot check / cpu baseline")
    if rank == 0:
        from speech_vecalign_b200.engine import records_to_alignments
        al0, _ = records_to_alignments(res[0]["recs"])
        parity = {"pair0_is_a_monotone_partition": [i for x, _ in al0 for i in x] == list(range(int(n0[0]))) and
                  [j for _, y in al0 for j in y] == list(range(int(n1[0])))}
        if int(n0[0]) * int(n1[0]) <= 2000 * 2000 and a <= 6:       # the CPU oracle finishes in seconds; the larger configurations are
            same, diff = oracle_check(pv[0][0].cpu().numpy(), pv[0][1].cpu().numpy(), types, w, 0, res[0]["recs"], global_stream_seed=4242)
            parity.update({"pair0_identical_alignments": same, "pair0_max_score_diff": diff})
        else:                                                        # pinned to it by tests/test_gpu_full_configs.py
            parity["oracle"] = "full-size parity of this configuration: tests/test_gpu_full_configs.py"

    cpu = None            # filled in by run_ours after every GPU arm has finished

    value = world * pairs * steps / (total_ms * 1e-3)
    rec = {
        "metric": "aligned doc pairs/sec (DP cells/sec alongside)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": data,
        "config": {"workload": desc, "pairs_per_gpu_per_step": pairs, "alignment_max_size": a, "dim": DIM,
                   "cost_mode": args.cost_mode, "streams": streams,
                   "l2": f"inputs {total * 4 / 2**30:.2f} GiB per step > 126 MB L2 (restored from a pristine copy before every step)",
                   **PARAMS},
        "dp_cells_per_sec": world * cells * steps / (total_ms * 1e-3),
        "alignments_per_step": n_align, "pairs_with_device_error": bad,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "wall_s_timed_loop": t_wall,
        "roofline": roofline, "serial_chain_ms_per_step": serial_ms, "kernels": kernels, "cpu_baseline": cpu, "parity": parity,
    }
    del run, pristine, work, pv, wv
    torch.cuda.empty_cache()
    return rec


def bench_cfg4(ctx, corpus_pairs, steps, warmup, e2e_steps, with_cpu, with_clocks):
    """BASELINE configs[3]: ONE corpus of `corpus_pairs` document pairs (200-800 segments, a=6), LPT-partitioned over
    the ranks (strong scaling), each shard aligned in chunks of <= 1024 pairs; the embeddings are resident in HBM in
    the reference's on-disk dtype (fp16, .embed files) and widened to fp32 working tensors chunk by chunk inside the
    timed step; the records of every pair are gathered on rank 0 in input order."""
    torch = ctx.torch
    import speech_vecalign_b200 as svb
    from speech_vecalign_b200 import capi, synth
    from speech_vecalign_b200.engine import BatchRun, WidenJobs, make_params, row_sources, workspace_bytes
    from speech_vecalign_b200.sharding import estimate_work, gather_in_order, lpt_partition
    args, rank, world, dev = ctx.args, ctx.rank, ctx.world, ctx.dev
    lib = capi.lib()
    desc, a, _, _ = WORKLOADS["cfg4"]
    k = a - 1
    types = svb.make_alignment_types(a)
    w = math.ceil(k / 2) + PARAMS["search_buffer_size"]
    mode = {"exact": capi.SVX_COST_EXACT, "fast": capi.SVX_COST_FAST, "tc": capi.SVX_COST_TC}[args.cost_mode]
    N0, N1 = synth.batch_sizes(corpus_pairs, seed=1234)              # the corpus is the same whatever the number of ranks
    shards = lpt_partition(estimate_work(N0, N1, a, PARAMS["search_buffer_size"]), world)
    mine = shards[rank]
    n0, n1 = N0[mine], N1[mine]
    P = len(mine)

    # ---- this rank's shard, resident in fp16 ------------------------------------------------------
    off0 = np.concatenate([[0], np.cumsum(k * (n0 + n1) * DIM)]).astype(np.int64)
    store = torch.empty(int(off0[-1]), dtype=torch.float16, device=dev)

    def views(buf, lo, hi, base=0):
        out = []
        for p in range(lo, hi):
            b = int(off0[p]) - base
            m0 = k * int(n0[p]) * DIM
            out.append((buf[b:b + m0].view(k, int(n0[p]), DIM), buf[b + m0:int(off0[p + 1]) - base].view(k, int(n1[p]), DIM)))
        return out

    sv = views(store, 0, P)
    note(f"cfg4: generating {P} of {corpus_pairs} pairs on rank 0")
    for p in range(P):
        v0, v1 = synth.synth_pair_torch(int(n0[p]), int(n1[p]), k, dim=DIM, seed=7_000_000 + int(mine[p]), device=dev)
        sv[p][0].copy_(v0)
        sv[p][1].copy_(v1)
    torch.cuda.synchronize()
    note("cfg4: planning the chunks")

    # ---- chunks: one launch chain each, sharing one fp32 work buffer and one arena -------------------
    bounds = list(range(0, P, CFG4_CHUNK)) + [P]
    prm = make_params(k, k, DIM, types, PARAMS["del_percentile_frac"], w, PARAMS["max_size_full_dp"], PARAMS["costs_sample_size"],
                      PARAMS["num_samps_for_norm"], mode, False, args.unfused_prologue)
    work_elems = max(int(off0[bounds[c + 1]] - off0[bounds[c]]) for c in range(len(bounds) - 1)) if P else 0
    arena_bytes = max(workspace_bytes(prm, n0[bounds[c]:bounds[c + 1]], n1[bounds[c]:bounds[c + 1]])[0] for c in range(len(bounds) - 1)) if P else 0
    work = torch.empty(work_elems, dtype=torch.float32, device=dev)
    arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
    chunks = []
    for c in range(len(bounds) - 1):
        lo, hi = bounds[c], bounds[c + 1]
        wv = views(work, lo, hi, base=int(off0[lo]))
        # the resident fp16 rows are the SOURCES of the level-0 prologue (read once, 2 bytes per element; the normalised
        # fp32 rows land in the shared work buffer); --widen-pass restores the separate widening pass (fp16 -> fp32
        # tensor -> prologue) that the unfused prologue still needs
        widen = WidenJobs([t for pr in sv[lo:hi] for t in pr], [t for pr in wv for t in pr], DIM, dev) if args.widen_pass else None
        srcs = None if args.widen_pass else (row_sources([t0 for t0, _ in sv[lo:hi]]), row_sources([t1 for _, t1 in sv[lo:hi]]))
        run = BatchRun([t0.data_ptr() for t0, _ in wv], [t1.data_ptr() for _, t1 in wv], n0[lo:hi], n1[lo:hi], k, k, DIM, types,
                       PARAMS["del_percentile_frac"], w, PARAMS["max_size_full_dp"], PARAMS["costs_sample_size"],
                       PARAMS["num_samps_for_norm"], dev, cost_mode=mode, seeds=[int(g) for g in mine[lo:hi]], arena=arena,
                       fused_prologue=not args.unfused_prologue, sources=srcs)
        run.keep_init_on_device()                 # the chunks take turns in one arena: descriptors + draws are restored device to device
        chunks.append((lo, hi, widen, run))
    cells = sum(ch[3].dp_cells() for ch in chunks)

    def one_step(timing=None, collect=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for lo, hi, widen, run in chunks:
            if timing is not None:
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
            if widen is not None:
                widen.run()                       # fp16 -> fp32 working tensors (the path normalises them in place)
            if timing is not None:
                b_.record()
            run.upload()                          # descriptors + draws of this chunk (the arena is shared)
            run.run(timing=timing is not None, ngroups=1)
            if timing is not None:
                kt = run.kernel_times()
                if widen is not None:
                    kt["svx_widen_fp16"] = a_.elapsed_time(b_)
                for nm, ms in kt.items():
                    timing[nm] = timing.get(nm, 0.0) + ms
            if collect is not None:
                collect.extend(run.results())     # synchronises; before the next chunk overwrites the arena
        e1.record()
        return e0, e1

    note(f"cfg4: timing {warmup} + {steps} steps")
    for _ in range(warmup):
        one_step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if (with_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    lib.svx_launch_count(1)
    t_wall = time.perf_counter()
    evs = [one_step() for _ in range(steps)]
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(lib.svx_launch_count(0))
    total_ms = ctx.max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs))
    ktimes = {}
    one_step(timing=ktimes)
    # results of every pair -> rank 0, input order (host-side gather of a few KB per pair: the only exchange)
    res = []
    one_step(collect=res)
    t_g = time.perf_counter()
    gathered = gather_in_order([(r["recs"].tobytes(), r["nrecs"], r["status"]) for r in res], mine, corpus_pairs)
    t_g = time.perf_counter() - t_g
    n_align, bad = ctx.sum_over_ranks([sum(r["nrecs"] for r in res), sum(1 for r in res if r["status"])])

    # ---- e2e: a sub-corpus from pinned host fp32 tensors through the public API, sharded the same way ------
    e2e = None
    note(f"cfg4: value done, e2e arm ({e2e_steps} steps)")
    if e2e_steps > 0:
        ne = min(CFG4_E2E_PAIRS, corpus_pairs)
        sub = lpt_partition(estimate_work(N0[:ne], N1[:ne], a, PARAMS["search_buffer_size"]), world)[rank]
        pos = {int(g): i for i, g in enumerate(mine)}
        sel = [pos[int(g)] for g in sub if int(g) in pos]
        extra = [int(g) for g in sub if int(g) not in pos]     # pairs of the sub-corpus that live on another rank: regenerate
        hv, hseeds = [], []
        for g in sub:
            g = int(g)
            if g in pos:
                s0_, s1_ = sv[pos[g]]
                h0 = torch.empty(s0_.shape, dtype=torch.float32, pin_memory=True); h0.copy_(s0_)
                h1 = torch.empty(s1_.shape, dtype=torch.float32, pin_memory=True); h1.copy_(s1_)
            else:
                v0, v1 = synth.synth_pair_torch(int(N0[g]), int(N1[g]), k, dim=DIM, seed=7_000_000 + g, device=dev)
                h0 = torch.empty(v0.shape, dtype=torch.float32, pin_memory=True); h0.copy_(v0.half())
                h1 = torch.empty(v1.shape, dtype=torch.float32, pin_memory=True); h1.copy_(v1.half())
            hv.append((h0, h1))
            hseeds.append(g)
        ebytes = sum(h0.numel() * 4 + h1.numel() * 4 for h0, h1 in hv)
        kw = dict(final_alignment_types=types, del_percentile_frac=PARAMS["del_percentile_frac"], width_over2=w,
                  max_size_full_dp=PARAMS["max_size_full_dp"], costs_sample_size=PARAMS["costs_sample_size"],
                  num_samps_for_norm=PARAMS["num_samps_for_norm"], cost_mode=args.cost_mode, output="records", seeds=hseeds)

        def e2e_step():
            out = svb.vecalign_batch(hv, **kw) if hv else []
            return gather_in_order([(o["recs"].tobytes(), o["nrecs"], o["status"]) for o in out], sub, ne)

        got = e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = e2e_step()
        ctx.barrier()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        tot_bytes = ctx.sum_over_ranks([ebytes])[0]
        e2e = {"value": ne * e2e_steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(tot_bytes),
               "d2h_bytes_per_step": int(sum(len(g_[0]) for g_ in got) if got else 0), "steps": e2e_steps,
               "ms_per_step": 1e3 * dt / e2e_steps, "corpus_pairs": ne, "h2d_gbs_achieved_aggregate": tot_bytes * e2e_steps / dt / 1e9,
               "api": "sharding.lpt_partition + speech_vecalign_b200.vecalign_batch(pinned host fp32 tensors, seeds=pair id, "
                      "output='records') + sharding.gather_in_order to rank 0"}
        if rank == 0 and got is not None and gathered is not None:
            e2e["records_equal_resident_arm"] = all(got[i] == gathered[i] for i in range(ne))
        del hv

    clocks = sampler.stop() if sampler else None
    alg, flops = {}, {}
    for _, _, widen, run in chunks:
        for nm, v in run.algorithmic_bytes().items():
            alg[nm] = alg.get(nm, 0) + v
        for nm, v in run.cost_flops().items():
            flops[nm] = flops.get(nm, 0) + v
    if args.widen_pass:
        alg["svx_widen_fp16"] = int(off0[-1]) * 6
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    kernels, roofline, serial_ms = launcher_report(ctx, ktimes, alg, flops, "cfg4", args.cost_mode, sm_mhz)

    parity = None
    note("cfg4: checksum / oracle spot check")
    if rank == 0:
        # one checksum over every pair's records in input order: equal across --gpus N <=> the sharded run reproduces the
        # single-GPU records bit for bit; plus the CPU oracle on the first pairs of this rank's shard
        h = hashlib.sha1()
        for g_ in gathered:
            h.update(g_[0])
        nchk = min(4, P)
        same, worst = True, 0.0
        for p in range(nchk):
            v0, v1 = sv[p][0].float().cpu().numpy(), sv[p][1].float().cpu().numpy()
            ok, diff = oracle_check(v0, v1, types, w, int(mine[p]), res[p]["recs"])
            same = same and bool(ok)
            worst = max(worst, diff if isinstance(diff, float) else 0.0)
        parity = {"records_sha1": h.hexdigest(), "oracle_pairs_checked": nchk, "identical_alignments": same, "max_score_diff": worst}

    cpu = None            # filled in by run_ours after every GPU arm has finished

    cells_all = ctx.sum_over_ranks([cells])[0]
    value = corpus_pairs * steps / (total_ms * 1e-3)
    rec = {
        "metric": "aligned doc pairs/sec (DP cells/sec alongside)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "corpus_pairs": corpus_pairs, "pairs_on_rank0": P, "chunk_pairs": CFG4_CHUNK, "alignment_max_size": a,
                   "dim": DIM, "cost_mode": args.cost_mode, "partition": "sharding.lpt_partition on sharding.estimate_work",
                   "resident_dtype": "fp16 (.embed on-disk dtype), " + ("widened to fp32 by a separate pass inside the step" if args.widen_pass else
                                                                      "read by the level-0 prologue inside the step (row sources)"),
                   "l2": f"rank 0 streams {int(off0[-1]) * 2 / 2**30:.2f} GiB of fp16 embeddings per step > 126 MB L2", **PARAMS},
        "dp_cells_per_sec": cells_all * steps / (total_ms * 1e-3),
        "alignments_per_step": n_align, "pairs_with_device_error": bad, "gather_to_rank0_s": t_g,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "wall_s_timed_loop": t_wall,
        "roofline": roofline, "serial_chain_ms_per_step": serial_ms, "kernels": kernels, "cpu_baseline": cpu, "parity": parity,
    }
    del chunks, store, work, arena, sv
    torch.cuda.empty_cache()
    return rec


def cpu_baseline(name):
    """The reference's CPU path on all host cores, one pair per worker (bounded sample).  Runs LAST: the workers are
    forked, and a forked child tearing down its copy of the CUDA state must not precede further GPU work here."""
    try:
        ref = CpuReference(name)
        ref.step(0)
        wall, c, kind, pp = ref.step(1)
        ref.close()
        return {"value": ref.nproc / wall, "unit": "pairs/s", "cores": ref.nproc, "kind": kind,
                "sample": f"{ref.nproc} worker processes x 1 pair of the workload (1 warm-up + 1 timed round), "
                          f"OpenBLAS 1 thread/worker; single-core {pp:.2f} s/pair",
                "dp_cells_per_sec": c / wall}
    except Exception as e:
        return {"error": repr(e)}


def brief(rec):
    """A config's entry of `all_configs`: the numbers, without the headline line's boilerplate."""
    keys = ("value", "unit", "scaling", "ms_per_step", "steps", "dp_cells_per_sec", "alignments_per_step", "pairs_with_device_error",
            "e2e", "roofline", "kernels", "parity", "gpu_launches", "config")
    return {k: rec[k] for k in keys if k in rec}


def run_ours(args):
    ctx = Ctx(args)
    if args.h2d_probe:
        p = h2d_probe(ctx)
        if ctx.rank == 0:
            emit(json.dumps({"h2d_probe": p, "n_gpus": ctx.world}))
        return
    e2e_steps = 0 if args.no_e2e else (args.e2e_steps or max(args.steps, 10))
    name = args.workload
    headline = "cfg2" if name in ("default", "all") else name
    pairs = args.pairs or WORKLOADS[headline][3]
    probe = None
    if e2e_steps:
        probe = h2d_probe(ctx)
    if headline == "cfg4":
        rec = bench_cfg4(ctx, pairs, args.steps, args.warmup, e2e_steps, not args.no_cpu_baseline, True)
    else:
        rec = bench_weak(ctx, headline, pairs, args.steps, args.warmup, e2e_steps, not args.no_cpu_baseline, True)
    if rec.get("e2e") and probe:
        rec["e2e"]["h2d_ceiling_gbs"] = probe["gbs_aggregate"]
        rec["e2e"]["h2d_probe"] = probe
        rec["e2e"]["frac_of_h2d_ceiling"] = rec["e2e"]["h2d_gbs_achieved_aggregate"] / probe["gbs_aggregate"]
    if name in ("default", "all") and not args.no_extras:
        # the other four configurations, a few steps each, so that one driver-run record covers all five
        xs, xe = min(args.steps, 3), (0 if args.no_e2e else 3)
        extra = {}
        for nm in ("cfg1", "cfg3", "cfg4", "cfg5"):
            try:
                if nm == "cfg4":
                    r = bench_cfg4(ctx, WORKLOADS[nm][3], xs, 3, xe, False, False)
                else:
                    r = bench_weak(ctx, nm, WORKLOADS[nm][3], xs, 3, xe, False, False)
                if r.get("e2e") and probe:
                    r["e2e"]["h2d_ceiling_gbs"] = probe["gbs_aggregate"]
                    r["e2e"]["frac_of_h2d_ceiling"] = r["e2e"]["h2d_gbs_achieved_aggregate"] / probe["gbs_aggregate"]
                extra[nm] = brief(r)
            except Exception as e:      # a failing extra must not lose the headline
                extra[nm] = {"error": repr(e)}
        rec["all_configs"] = {"cfg2": "the headline fields of this line", **extra}
    if ctx.world == 1 and ctx.rank == 0 and not args.no_cpu_baseline:
        note("cpu baseline (reference CPU path on all host cores)")
        rec["cpu_baseline"] = cpu_baseline(headline)
    if ctx.rank == 0:
        emit(json.dumps(rec))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


class StdoutToStderr:
    """Everything libraries print on fd 1 while the benchmark runs (NCCL's version banner, warnings of forked
    workers) goes to stderr; stdout carries exactly one line: the JSON record."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None
_T0 = time.perf_counter()


def note(msg):
    """progress on stderr (stdout carries only the JSON record)"""
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.perf_counter() - _T0:7.1f} s] {msg}", file=sys.stderr, flush=True)


def emit(line):
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    global OUT
    args = parse_args()
    if args.gpus > 1 and "RANK" not in os.environ and args.impl == "ours":
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    with StdoutToStderr() as OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
