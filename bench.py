#!/usr/bin/env python
"""bench.py - the headline benchmark of BASELINE.json on B200:
IndexFlatIP exact top-100 over 10M x 768 fp32 (ViT-L/14 dim), synthetic unit vectors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one search of a batch of B queries (default B=1: the bandwidth-bound case the
reference serves, /root/reference/api/routes.py:1407) over the whole database:
  value    QPS with queries already in HBM (wb_search_dev), CUDA events on the launching stream
  e2e      QPS through the public API `index.search(numpy, k)` with pinned HOST buffers: H2D of
           the queries and D2H of (D, I) inside the timed region
  roofline achieved = N*d*4 bytes / scan-kernel duration (events inside the library, same stream)
           vs the measured copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the faiss-equivalent C restatement (oracle/cpu_flat.c) on this box's host cores,
           on a bounded row sample, scaled linearly (faiss itself is not installable: BASELINE.md 3)
N > 1: the 10M rows are split into N contiguous shards (strong scaling), one process per GPU, one
NCCL all-gather of the k candidates per step, K3 merge on every rank.
`--impl reference` times the CPU restatement with all host threads (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "QPS, IndexFlatIP top-100 (10Mx768 fp32)"
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--sweep", default="", help="comma list of extra batch sizes to report, e.g. 2,4,8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mixed", action="store_true",
                    help="config 5: half of the batch are 'combined' queries normalise(2.0*t + 1.0*i - 0.2*n) "
                         "(/root/reference/api/routes.py:759-850)")
    return ap.parse_args()


def workload_name(a):
    return f"IndexFlatIP top-{a.k} over {a.rows}x{a.dim} fp32, query batch {a.batch}" + (" (mixed text/combined)" if a.mixed else "")


# ------------------------------------------------------------------------------------------------
# synthetic "CLIP-like" data: rows = normalise(c_j + 0.6 g), generated on the device
# ------------------------------------------------------------------------------------------------
def fill_index_clustered(index, lo, hi, d, seed, device, chunk=500_000, ncentres=4096):
    import torch
    from wise_b200 import _capi
    L = _capi.lib()
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    centres = torch.nn.functional.normalize(torch.randn(ncentres, d, device=device, generator=gen), dim=1)
    st = torch.cuda.current_stream().cuda_stream
    index.reserve(hi - lo)
    sample = None
    for s in range(lo, hi, chunk):
        e = min(hi, s + chunk)
        g2 = torch.Generator(device=device)
        g2.manual_seed(seed * 1_000_003 + s)  # chunk-addressed: same rows whatever the sharding
        j = torch.randint(0, ncentres, (e - s,), device=device, generator=g2)
        x = centres[j] + 0.6 * torch.randn(e - s, d, device=device, generator=g2) / (d ** 0.5)
        x = torch.nn.functional.normalize(x, dim=1).contiguous()
        ids = torch.arange(s, e, dtype=torch.int64, device=device)
        _capi.check(L.wb_add_with_ids_dev(index._h, e - s, x.data_ptr(), ids.data_ptr(), st))
        torch.cuda.synchronize()
        if sample is None:
            sample = x
    return centres, sample


MIXED = False


def make_queries(centres, nq, d, seed, device):
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def draw(m):
        j = torch.randint(0, centres.shape[0], (m,), device=device, generator=gen)
        v = centres[j] + 0.6 * torch.randn(m, d, device=device, generator=gen) / (d ** 0.5)
        return torch.nn.functional.normalize(v, dim=1)

    q = draw(nq)
    if MIXED and nq > 1:  # second half: text x2.0 + image x1.0 - negative x0.2, renormalised
        h = nq // 2
        q[h:] = torch.nn.functional.normalize(2.0 * q[h:] + 1.0 * draw(nq - h) - 0.2 * draw(nq - h), dim=1)
    return q.contiguous()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU every 20 ms while the timed region runs (NVML through
    nvidia_ml_py: `nvidia-smi -lms` block-buffers its pipe output and delivered only 1-7 lines per run)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._halt.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                time.sleep(0.02)
        except Exception:
            pass

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "source": "nvml"}


def cpu_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"model": model, "cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}


def cpu_restatement(sample_rows, q, k, n_full, mt, reps=3):
    """Time oracle/cpu_flat.c (faiss-equivalent restatement) on `sample_rows`, scale to n_full rows."""
    from oracle import cpu as OC
    OC.flat_search(sample_rows[:1000], q, k, mt=mt)  # warm-up (thread pool, page-in)
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        OC.flat_search(sample_rows, q, k, mt=mt)
        ts.append(time.perf_counter() - t)
    t_med = float(np.median(ts))
    scale = n_full / sample_rows.shape[0]
    return q.shape[0] / (t_med * scale), t_med, (OC.max_threads() if mt else min(q.shape[0], OC.max_threads())), OC.simd_name()


# ------------------------------------------------------------------------------------------------
def run_reference(a):
    """--impl reference: the reference's CPU path (faiss-equivalent restatement; faiss is absent),
    all host threads, bounded sample.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as O
    n_s = min(a.cpu_sample_rows, a.rows)
    xb = O.clustered_unit(n_s, a.dim, 4096, 2024)
    q = O.clustered_unit(a.batch, a.dim, 4096, 2025)
    from oracle import cpu as OC
    OC.flat_search(xb[:1000], q, a.k, mt=True)
    for _ in range(a.warmup):
        OC.flat_search(xb, q, a.k, mt=True)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        OC.flat_search(xb, q, a.k, mt=True)
    dt = time.perf_counter() - t0
    scale = a.rows / n_s
    ms_step = dt / a.steps * 1e3 * scale
    qps = a.batch / (ms_step / 1e3)
    sample = (f"{n_s} of {a.rows} rows per step (time scaled x{scale:g}); oracle/cpu_flat.c orc_flat_search_mt, "
              f"{OC.simd_name()}, {OC.max_threads()} threads; faiss itself is not installable here")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a), "host": cpu_info()},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": OC.max_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(a):
    import torch
    import torch.distributed as dist
    from wise_b200 import _capi
    from wise_b200 import faiss_compat as faiss
    from wise_b200.sharded import ShardedIndex, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: wise_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    L = _capi.lib()

    lo, hi = shard_range(a.rows, rank, world)
    index = faiss.IndexIDMap(faiss.IndexFlatIP(a.dim, device=local_rank))
    centres, first_rows = fill_index_clustered(index, lo, hi, a.dim, 2024, device)
    assert index.ntotal == hi - lo
    sharded = ShardedIndex(index)
    L.wb_set_timing(index._h, 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(nq, steps, warmup):
        """Device-resident queries; returns (ms per step [max over ranks], scan-kernel ms per step, launches)."""
        qs = make_queries(centres, nq * (steps + warmup), a.dim, 2025, device).view(steps + warmup, nq, a.dim)
        for i in range(warmup):
            sharded.search_dev(qs[i], a.k)
        barrier()
        l0 = L.wb_launch_count(index._h)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            sharded.search_dev(qs[warmup + i], a.k)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1) / steps
        launches = L.wb_launch_count(index._h) - l0 + (steps if world > 1 else 0)  # + K3 merge of the gathered parts
        # per-launch scan durations OF THE TIMED REGION: event pairs recorded by the library around the
        # scan kernel on the launching stream (ring of 128), read back after the region has ended
        import ctypes
        buf = (ctypes.c_float * 128)()
        n = L.wb_scan_ms_history(index._h, buf, min(steps, 128))
        scan_ms = [buf[i] for i in range(n)]
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(np.mean(scan_ms)), int(launches)

    def timed_e2e(nq, steps, warmup):
        """Public API, pinned host buffers in, numpy out; wall clock around synchronous calls."""
        q_all = make_queries(centres, nq * (steps + warmup), a.dim, 2026, device).view(steps + warmup, nq, a.dim).cpu()
        pinned = torch.empty((nq, a.dim), dtype=torch.float32).pin_memory()
        api = index if world == 1 else sharded  # the call a user makes: index.search(numpy, k)
        for i in range(warmup):
            pinned.copy_(q_all[i])
            api.search(pinned.numpy(), a.k)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            pinned.copy_(q_all[warmup + i])
            D, I = api.search(pinned.numpy(), a.k)
        barrier()
        dt = (time.perf_counter() - t0) / steps
        t = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3, D, I

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # pre-roll (untimed, all ranks): keep the GPU under the same load for ~0.6 s so that nvidia-smi, which needs a
    # few hundred ms to start, has samples that bracket the timed region instead of one stray reading
    # (the number of pre-roll searches must be the same on every rank: the exchange is a collective)
    qpre = make_queries(centres, a.batch, a.dim, 2024, device)
    barrier()
    t_pre = time.perf_counter()
    for _ in range(3):
        sharded.search_dev(qpre, a.k)
    torch.cuda.synchronize()
    t3 = torch.tensor([(time.perf_counter() - t_pre) / 3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    for _ in range(int(min(2000, max(1, 0.6 / max(float(t3.item()), 1e-5))))):
        sharded.search_dev(qpre, a.k)
    torch.cuda.synchronize()
    ms_step, scan_ms, launches = timed_steps(a.batch, a.steps, a.warmup)
    clocks = sampler.finish() if sampler else None
    e2e_ms, D_last, I_last = timed_e2e(a.batch, a.steps, a.warmup)

    sweep = {}
    for b in [int(t) for t in a.sweep.split(",") if t.strip()]:
        m, s_ms, _ = timed_steps(b, max(3, a.steps // 3), 2)
        sweep[str(b)] = {"qps": b / (m / 1e3), "ms_per_step": m, "scan_ms": s_ms,
                         "scan_gbs": (hi - lo) * a.dim * 4 / (s_ms * 1e-3) / 1e9}

    if rank == 0:
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_file):
            peak, peak_src = float(json.load(open(peaks_file))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        local_bytes = (hi - lo) * a.dim * 4
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            t = json.load(open(tf))
            if t.get("rows") == hi - lo and t.get("dim") == a.dim and t.get("batch") == a.batch:
                traffic = t.get("dram_bytes_per_launch")
        if a.batch < 5:   # K1: CUDA-core scan, HBM-bound; one pass of the row store per <= 8 queries
            passes = (a.batch + 7) // 8
            achieved = local_bytes * passes / (scan_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "kernel": "scan_topk_kernel", "bytes_per_launch": local_bytes,
                        "launch_ms": scan_ms / passes, "peak_source": peak_src}
        elif a.batch <= 128:  # K2, one 128-query block at most: ONE pass over the rows per search -> HBM is the roofline
            achieved = local_bytes / (scan_ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "kernel": "gemm_topk_kernel (all epochs of a search, compactions included)",
                        "bytes_per_launch": local_bytes, "launch_ms": scan_ms, "peak_source": peak_src,
                        "note": "tcgen05 TF32 filter epochs + exact fp32 re-scoring of the candidates"}
        else:             # K2, several 128-query blocks per row tile: tensor pipe; all epochs of a search timed together
            flop = 2.0 * a.batch * (hi - lo) * a.dim
            achieved = flop / (scan_ms * 1e-3) / 1e12
            if os.path.exists(peaks_file):
                pk = json.load(open(peaks_file))
                tpeak = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])) / 2.0
                tsrc = "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32 issues at half the bf16 rate; no tf32 measurement exists)"
            else:
                tpeak, tsrc = 1400.0 / 2.0, "fallback 1.4 PF bf16 sustained / 2 (B200_PROFILING.md)"
            terms = 3 if os.environ.get("WB_GEMM_FILTER", "1") == "0" else 1
            roofline = {"bound": "tensor", "achieved": achieved, "peak": tpeak, "unit": "TFLOP/s", "frac": achieved / tpeak,
                        "traffic": traffic, "kernel": "gemm2_topk_kernel (all epochs of a search)", "flop_per_search": flop,
                        "search_ms": scan_ms, "issued_tflops": terms * achieved, "issued_frac": terms * achieved / tpeak,
                        "note": ("3xTF32 in every epoch: each algorithmic flop is issued 3 times" if terms == 3 else
                                 "one-term TF32 filter epochs (issued ~ algorithmic flops) + exact fp32 re-scoring; the "
                                 "kernel is bound by its per-chunk row pipeline, not by the tensor pipe (DESIGN.md 6)"),
                        "peak_source": tsrc}
        line = {
            "metric": METRIC, "value": a.batch / (ms_step / 1e3), "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "rows_per_gpu": hi - lo, "parallelism": f"row-shard x{world}",
                       "l2": "database shard (>= 3.8 GB) is far larger than the 126 MB L2; no flush needed",
                       "generator": "clustered unit vectors, 4096 centres, noise 0.6, seed 2024/2025"},
            "roofline": roofline,
            "e2e": {"value": a.batch / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": a.batch * a.dim * 4, "d2h_bytes_per_step": a.batch * a.k * 12},
            "gpu_launches": launches, "clocks": clocks,
        }
        if sweep:
            line["batch_sweep"] = sweep
        if world == 1 and not a.no_cpu_baseline:
            n_s = min(a.cpu_sample_rows, hi - lo)
            xs = np.empty((n_s, a.dim), np.float32)
            _capi.check(L.wb_export_rows(index._h, 0, n_s, _capi.ptr(xs), None, None))
            qh = make_queries(centres, a.batch, a.dim, 2027, device).cpu().numpy()
            qps1, t1, cores1, simd = cpu_restatement(xs, qh, a.k, a.rows, mt=False)
            qpsm, tm, coresm, _ = cpu_restatement(xs, qh, a.k, a.rows, mt=True)
            line["cpu_baseline"] = {
                "value": qps1, "unit": "queries/s", "cores": cores1, "kind": "port",
                "sample": f"first {n_s} of {a.rows} rows, {a.batch} queries, 3 reps, median {t1:.3f} s scaled x{a.rows / n_s:g}; "
                          f"oracle/cpu_flat.c orc_flat_search_seq ({simd}) = faiss's own threading (parallel over queries only)",
                "all_cores": {"value": qpsm, "cores": coresm, "median_s": tm, "fn": "orc_flat_search_mt"},
                "host": cpu_info()}
            # spot-check the GPU answer of the last e2e step against the restatement on the sample
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global MIXED
    a = parse()
    MIXED = a.mixed
    # Keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner) get stderr.
    real_stdout = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    real_stdout.flush()


if __name__ == "__main__":
    main()
