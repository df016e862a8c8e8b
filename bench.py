#!/usr/bin/env python
"""Headline benchmark: ReID eval queries/s (distance + ranking + CMC/mAP) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload large|rgbnt100|rgbnt201]
    python bench.py --impl reference ...        # the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[3], the config the 1/2/4/8-GPU metric is quoted on): 20 000
queries x 1 000 000 gallery, d = 1536, 50 000 ids, 8 cams, sigma 4, no re-ranking; the gallery is
sharded contiguously over the N ranks, queries are replicated (strong scaling).  One step = one
complete evaluation: label plan, L2-normalisation + fp16 split of both sets, same-identity
records (tcgen05 extract GEMM), thresholds, fused tcgen05 distance + rank-count GEMM, [N>1:
all-gather of records / all-reduce of counts], CMC/mAP finalisation and the D2H read of the metrics.

    value   device-resident features (CUDA events around K steps, max over ranks)
    e2e     ONE call of ShardedEvaluator.evaluate_host per step on PINNED HOST features and labels:
            the gallery is pulled over PCIe by the prepare kernel itself, queried rows first, and
            ranked slab by slab while the rest is still in flight; the metrics are read back.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import builtins
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Q, G, d, nid, ncam, sigma)
    "large": (20000, 1000000, 1536, 50000, 8, 4.0),
    "rgbnt100": (1715, 8575, 1536, 50, 8, 4.0),
    "rgbnt201": (836, 836, 1536, 30, 2, 4.0),
}
METRIC = "reid_eval_queries_per_sec"
UNIT = "queries/s"
CPU_CHUNK_LARGE = 128    # queries per CPU step on the 1 M gallery: ONE chunk size for both CPU numbers


def workload_desc(name):
    Q, G, d, nid, ncam, sigma = WORKLOADS[name]
    return ("%s: %d queries x %d gallery, d=%d fp32, %d ids, %d cams, sigma=%g, distance + ranking + CMC/mAP, "
            "no re-ranking (BASELINE.json configs[%d])" % (name, Q, G, d, nid, ncam, sigma,
                                                           {"large": 3, "rgbnt100": 2, "rgbnt201": 0}[name]))


def config_of(name):
    """Identical on both arms (the driver compares the dicts)."""
    return {"workload": workload_desc(name)}


def load_traffic(name, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/),
    valid for the workload / GPU count it was captured on; None otherwise."""
    p = os.path.join(ROOT, "profiles", "count_kernel_traffic.json")
    if os.path.exists(p):
        j = json.load(open(p))
        if j.get("workload") == name and j.get("n_gpus") == world:
            return j
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU side: the reference's algorithm (oracle port; the reference is pure Python and
# /root/reference does not exist on the GPU box) -- ONE code path for the reference arm and for
# the cpu_baseline leg of our arm
# ----------------------------------------------------------------------------------------------
def host_data(name, q_rows=None, seed=0):
    Q, G, d, nid, ncam, sigma = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((nid, d), dtype=np.float32)
    q_pid, g_pid = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    q_cam, g_cam = rng.integers(0, ncam, Q), rng.integers(0, ncam, G)
    nq = Q if q_rows is None else min(Q, q_rows)
    qf = centers[q_pid[:nq]] + np.float32(sigma) * rng.standard_normal((nq, d), dtype=np.float32)
    gf = np.empty((G, d), np.float32)
    for s in range(0, G, 65536):
        e = min(G, s + 65536)
        gf[s:e] = centers[g_pid[s:e]] + np.float32(sigma) * rng.standard_normal((e - s, d), dtype=np.float32)
    return qf, gf, q_pid[:nq], g_pid, q_cam[:nq], g_cam


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use the host's cores.
    Returns the BLAS thread count in effect."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=cores)
        blas = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(blas) if blas else torch.get_num_threads()
    except ImportError:
        return torch.get_num_threads()


def cpu_chunk_size(name):
    Q, G = WORKLOADS[name][:2]
    return CPU_CHUNK_LARGE if G > 100000 else Q


def cpu_reference_step(oracle, qf, gf_n, qp, gp, qc, gc, phases=None):
    """compute()-equivalent of the reference on a query chunk (eval_func treats queries
    independently, so chunking is exact): normalise -> euclidean_distance -> eval_func with the
    reference's default np.argsort (utils/metrics.py:341-369, 395-401, 110-169)."""
    t0 = time.perf_counter()
    qn = oracle.l2_normalize(qf)
    dist = oracle.euclidean_distance(qn, gf_n)
    t1 = time.perf_counter()
    timing = {}
    out = oracle.eval_func(dist, qp, gp, qc, gc, sort_kind=None, timing=timing)
    t2 = time.perf_counter()
    if phases is not None:
        phases.append({"normalise+sgemm_distance_s": t1 - t0, "argsort_s": timing.get("argsort_s"),
                       "per_query_loop_s": timing.get("loop_s"), "total_s": t2 - t0})
    return out


def cpu_measure(name, qf, gf, qp, gp, qc, gc, steps, warmup, budget_s):
    """Times `steps` CPU steps of the fixed chunk (after `warmup` untimed ones; at least one, so the
    sgemm / page cache are warm), cutting the number of timed steps so that the run stays within
    `budget_s`.  Returns (queries/s, ms per step, timed steps, sample text, phase split)."""
    from oracle import reid_oracle as oracle
    chunk = min(cpu_chunk_size(name), len(qf))
    G = gf.shape[0]
    t0 = time.perf_counter()
    gf_n = oracle.l2_normalize(gf)      # gallery normalisation amortised over the query chunks of one evaluation
    t_norm = time.perf_counter() - t0
    args = (oracle, qf[:chunk], gf_n, qp[:chunk], gp, qc[:chunk], gc)
    t0 = time.perf_counter()
    for _ in range(max(1, min(warmup, 1))):
        cpu_reference_step(*args)
    per = time.perf_counter() - t0
    done = max(1, min(steps, int(max(budget_s - per, per) / max(per, 1e-9))))
    phases, times = [], []
    for _ in range(done):
        t0 = time.perf_counter()
        cpu_reference_step(*args, phases=phases)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    split = {k: float(np.mean([p[k] for p in phases])) for k in phases[0] if phases[0][k] is not None}
    sample = ("%d-query chunk x full %d gallery per step, %d timed step(s) after 1 warm-up, %.2f s per step "
              "(normalise + fp32 sgemm distance + np.argsort + per-query CMC/AP loop: oracle port of "
              "utils/metrics.py:110-169,341-401; sgemm multi-threaded, argsort and the loop single-threaded as in the "
              "reference; gallery normalisation %.2f s once, outside the timed step)" % (chunk, G, done, ms * 1e-3, t_norm))
    return chunk / (ms * 1e-3), ms, done, sample, split


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    blas_threads = use_all_host_threads()
    name = args.workload
    cores = os.cpu_count() or 1
    qf, gf, qp, gp, qc, gc = host_data(name, q_rows=cpu_chunk_size(name))
    value, ms, done, sample, split = cpu_measure(name, qf, gf, qp, gp, qc, gc, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(name),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "blas_threads": blas_threads, "phase_split_s": split},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from demo2_b200 import metrics, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    Q, G, d, nid, ncam, sigma = WORKLOADS[name]
    peaks = load_peaks()

    # ---- synthetic data: labels from numpy (same on every rank), features drawn on the device ----
    rng = np.random.default_rng(0)
    q_pid, g_pid = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    q_cam, g_cam = rng.integers(0, ncam, Q), rng.integers(0, ncam, G)
    lo, hi = parallel.shard_range(G, world, rank)
    gen = torch.Generator(device=dev).manual_seed(1234)
    centers = torch.randn(nid, d, device=dev, generator=gen)
    qf = centers[torch.from_numpy(q_pid).to(dev)] + sigma * torch.randn(Q, d, device=dev, generator=gen)

    def draw_gallery(r):
        a, b = parallel.shard_range(G, world, r)
        gen_g = torch.Generator(device=dev).manual_seed(99 + r)
        out = torch.empty(b - a, d, device=dev)
        gp_dev = torch.from_numpy(g_pid[a:b]).to(dev)
        for s in range(0, b - a, 131072):
            e = min(b - a, s + 131072)
            out[s:e] = centers[gp_dev[s:e]] + sigma * torch.randn(e - s, d, device=dev, generator=gen_g)
        return out

    gl = hi - lo
    gf = draw_gallery(rank)
    labels = dict(q_pid=torch.from_numpy(q_pid).int().to(dev), g_pid=torch.from_numpy(g_pid[lo:hi]).int().to(dev),
                  q_cam=torch.from_numpy(q_cam).int().to(dev), g_cam=torch.from_numpy(g_cam[lo:hi]).int().to(dev))

    ev = parallel.ShardedEvaluator(world=world, rank=rank, group=None if world == 1 else dist.group.WORLD)

    def step(qx, gx, lab, timers=None):
        return ev.evaluate(qx, gx, lab["q_pid"], lab["g_pid"], lab["q_cam"], lab["g_cam"], g_index_base=lo,
                           normalize=True, max_rank=50, timers=timers)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        t = torch.tensor(values, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def stage_means(timers):
        return {k: float(np.mean([t[k][0].elapsed_time(t[k][1]) for t in timers]))
                for k in timers[0] if isinstance(timers[0][k], tuple)}

    # ---- device-resident timing ----
    for _ in range(max(args.warmup, 3)):
        res = step(qf, gf, labels)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.collect()
    gc.disable()           # no interpreter garbage collection (10 .. 100 ms pauses) inside a timed region
    barrier()
    e0.record()
    for _ in range(args.steps):
        t = {}
        res = step(qf, gf, labels, timers=t)
        timers.append(t)
    e1.record()
    barrier()
    gc.enable()
    clocks = sampler.stop() if rank == 0 else None
    count_ms_local = float(np.mean([t["count"][0].elapsed_time(t["count"][1]) for t in timers]))
    ms_total, count_ms = max_over_ranks([e0.elapsed_time(e1), count_ms_local])
    ms_step = ms_total / args.steps
    value = Q / (ms_step * 1e-3)
    launches = int(np.mean([t["launches"] for t in timers]))
    # dominant kernel: the fused count GEMM, CUDA-event time on its launch stream
    algo_tflop = 2.0 * Q * gl * d * 1e-12
    passes = 3
    peak = peaks["bf16_tflops_sustained"] / passes
    achieved = algo_tflop / (count_ms * 1e-3)
    stage_ms = stage_means(timers)
    non_gemm_ms = ms_step - count_ms

    # ---- multi-GPU checks (outside the timed region): the other sharded flows over NCCL ----
    mgc = multi_gpu_checks(dev, world, rank, res, qf, draw_gallery, q_pid, g_pid, q_cam, g_cam) if world > 1 else None

    # ---- end to end through the public API with HOST (pinned) buffers, one call per step ----
    q_host = qf.cpu().pin_memory()
    g_host = gf.cpu().pin_memory()
    lab_host = {k: v.cpu().pin_memory() for k, v in labels.items()}
    del qf, gf
    torch.cuda.empty_cache()

    def e2e_step(timers=None):
        return ev.evaluate_host(q_host, g_host, lab_host["q_pid"], lab_host["g_pid"], lab_host["q_cam"],
                                lab_host["g_cam"], g_index_base=lo, normalize=True, max_rank=50, timers=timers)

    for _ in range(2):
        res2 = e2e_step()
    barrier()
    n_e2e = max(3, min(args.steps, 6))
    e2e_timers = []
    gc.collect()
    gc.disable()
    e0.record()
    for _ in range(n_e2e):
        t = {}
        res2 = e2e_step(timers=t)
        e2e_timers.append(t)
    e1.record()
    barrier()
    gc.enable()
    e2e_total = e0.elapsed_time(e1)
    same = bool(res2.mAP == res.mAP and np.array_equal(res2.cmc, res.cmc))

    # the same without overlap, for comparison: copy everything to the device, then evaluate
    slot = (torch.empty_like(q_host, device=dev), torch.empty_like(g_host, device=dev),
            {k: torch.empty_like(v, device=dev) for k, v in lab_host.items()})

    def staged_step():
        slot[0].copy_(q_host, non_blocking=True)
        slot[1].copy_(g_host, non_blocking=True)
        for k, v in lab_host.items():
            slot[2][k].copy_(v, non_blocking=True)
        return step(*slot)

    staged_step()
    barrier()
    e0.record()
    for _ in range(2):
        staged_step()
    e1.record()
    barrier()
    e2e_ms, staged_ms = max_over_ranks([e2e_total / n_e2e, e0.elapsed_time(e1) / 2])
    del slot
    h2d = (q_host.numel() + g_host.numel()) * 4 + sum(v.numel() * 4 for v in lab_host.values())
    if world > 1 and isinstance(ev.coll, parallel.LibCollectives):
        h2d -= (q_host.numel() - (-(-Q // world)) * d) * 4     # each rank uploads 1/P of the queries; the rest comes over NVLink
    # metrics (mAP | num_valid | cmc) + plan sizes (+ the piece boundaries of the grouped single-GPU path)
    d2h = 32 + 50 * 4 + (world * (5 * 8) + 8 if world > 1 else 16 + 4 * max(e2e_timers[0].get("query_groups", 1) - 1, 0))

    traffic = load_traffic(name, world)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 hi/lo split on tcgen05, fp32 accumulate)",
                "data": "synthetic", "config": config_of(name),
                "run": {"sharding": "gallery rows split contiguously over ranks, queries replicated",
                        "l2": "inputs exceed L2 (gallery shard %.1f GB read per step)" % (gl * d * 4e-9),
                        "collectives": ev.coll.name if ev.coll is not None else None,
                        "result": {"mAP": float(res.mAP), "rank1": float(res.cmc[0]), "num_valid": int(res.num_valid)}},
                "clocks": clocks,
                "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "staged_ms_per_step": staged_ms, "staged_value": Q / (staged_ms * 1e-3),
                        "identical_to_device_resident_result": same,
                        "stage_ms": stage_means(e2e_timers), "slabs": e2e_timers[0].get("slabs"),
                        "query_groups": e2e_timers[0].get("query_groups", 1),
                        "queried_gallery_rows": e2e_timers[0].get("queried_rows"),
                        "api": "demo2_b200.parallel.ShardedEvaluator.evaluate_host: ONE call per step on pinned host "
                               "features + labels; the prepare kernel pulls the gallery over PCIe (no fp32 copy in HBM), "
                               "queried rows first (one GPU: per block of pid-sorted queries, so the count GEMM starts after "
                               "the first block's rows), and the count GEMM ranks every piece while the next ones are in "
                               "flight; cmc/mAP are read back.  staged_* = cudaMemcpy of everything, then the "
                               "device-resident call."},
                "gpu_launches": launches * args.steps,
                "roofline": {"bound": "tensor",
                             "kernel": "sqdist_gemm2_kernel<EpiCount> (CTA-pair tcgen05 GEMM, fused distance + rank-count)",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                             "traffic_note": (traffic["note"] if traffic else
                                              "no ncu capture for this workload / GPU count (see profiles/)"),
                             "ms_per_launch": count_ms,
                             "peak_note": "%s bf16 sustained %.1f TFLOP/s / %d fp16 split passes (hi*hi + hi*lo + lo*hi); "
                                          "achieved = 2*Q*G_local*d algorithmic flop / CUDA-event time"
                                          % (peaks["source"], peaks["bf16_tflops_sustained"], passes),
                             "executed_tflops": achieved * passes,
                             "frac_vs_burst": achieved / (peaks["bf16_tflops"] / passes),
                             "burst_note": "short launches (N >= 4) run above the sustained clock; frac_vs_burst uses the "
                                           "measured burst bf16 peak"},
                "stage_ms": stage_ms, "count_stage_ms": count_ms, "non_gemm_ms_per_step": non_gemm_ms}
        if mgc is not None:
            line["multi_gpu_checks"] = mgc
        # GPU legs first, CPU legs last: after a CPU leg the host's BLAS / OpenMP worker threads keep
        # spinning on every core for a while and the launch-bound sub-millisecond workloads measured
        # 3-15x slower (rgbnt100_eval 0.75 ms -> 2 .. 11 ms) -- a measurement artefact, not a kernel time
        if world == 1 and not args.no_other:
            try:
                line["other_workloads"] = other_workloads(dev, peaks)
            except Exception as exc:  # never lose the headline line
                line["other_workloads"] = {"error": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(name, q_host, g_host, q_pid, g_pid, q_cam, g_cam)
            if isinstance(line.get("other_workloads"), dict) and "error" not in line["other_workloads"]:
                try:
                    other_cpu_baselines(line["other_workloads"])
                except Exception as exc:
                    line["other_workloads"]["cpu_error"] = repr(exc)
        print(json.dumps(line))
    if world > 1:
        barrier()
        parallel.LibCollectives.destroy()
        dist.destroy_process_group()


def multi_gpu_checks(dev, world, rank, res, qf, draw_gallery, q_pid, g_pid, q_cam, g_cam):
    """Outside the timed region, over the real process group: (1) the gallery-sharded result of the
    headline run == the one-GPU evaluation of the whole problem (every rank re-draws all shards);
    (2) row-sharded re-ranking at RGBNT100 scale == one-GPU re_ranking, bit for bit; (3) evaluation
    under DDP (every rank keeps the features it extracted) == one-GPU evaluation."""
    import torch
    import torch.distributed as dist
    from demo2_b200 import metrics, parallel, reranking, synth
    out = {}
    # (1)
    g_all = torch.cat([draw_gallery(r) for r in range(world)])
    whole = metrics.evaluate_features(qf, g_all, q_pid, g_pid, q_cam, g_cam, normalize=True)
    del g_all
    ok1 = bool(whole.mAP == res.mAP and np.array_equal(whole.cmc, res.cmc) and torch.equal(whole.first, res.first))
    # (2)
    s = synth.make_named("rgbnt100", sigma=5.0, seed=0)
    q, g = s.qf.to(dev), s.gf.to(dev)
    rr = parallel.ShardedReranker(world=world, rank=rank, group=dist.group.WORLD)
    sharded = rr.re_ranking(q, g, 20, 6, 0.3, normalize=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    sharded = rr.re_ranking(q, g, 20, 6, 0.3, normalize=True)
    e1.record()
    torch.cuda.synchronize()
    whole_rr = reranking.re_ranking_device(q, g, 20, 6, 0.3, normalize=True)
    ok2 = bool(torch.equal(whole_rr, sharded))
    rr_ms = e0.elapsed_time(e1)
    # (3)
    feats = torch.cat([q, g])
    pids = np.concatenate([s.q_pids, s.g_pids])
    cams = np.concatenate([s.q_camids, s.g_camids])
    from torch.utils.data import DistributedSampler
    mine = np.asarray(list(DistributedSampler(range(len(pids)), num_replicas=world, rank=rank, shuffle=False)))
    ddp = parallel.DistributedR1mAP(s.num_query, world=world, rank=rank, group=dist.group.WORLD, feat_norm=True)
    for a in range(0, len(mine), 256):
        b = mine[a:a + 256]
        ddp.update((feats[torch.from_numpy(b).to(dev)], pids[b], torch.from_numpy(cams[b]), torch.from_numpy(b)))
    cmc, mAP = ddp.compute()
    one = metrics.evaluate_features(q, g, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True)
    ok3 = bool(mAP == one.mAP and np.array_equal(cmc, one.cmc) and torch.equal(ddp.last_result.first, one.first))
    flags = torch.tensor([ok1, ok2, ok3], device=dev, dtype=torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok1, ok2, ok3 = (bool(v) for v in flags.tolist())
    out.update({"sharded_eval_identical": ok1, "rerank_identical": ok2, "ddp_eval_identical": ok3,
                "rerank_ms": rr_ms, "rerank_shape": "RGBNT100 1715 + 8575, k1=20 k2=6, rows sharded over %d ranks" % world,
                "ddp_padded_samples": int(len(mine) * world - len(pids)),
                "note": "bit-for-bit comparisons with the one-GPU result, AND-reduced over all ranks"})
    return out


def other_cpu_baselines(res):
    """CPU port of the reference on the small configs (whole workload once), added to the entries
    other_workloads() produced."""
    from demo2_b200 import synth
    from oracle import reid_oracle as oracle
    for key in ("rgbnt201", "rgbnt100"):
        s = synth.make_named(key, sigma=4.0, seed=0)
        Q = s.qf.shape[0]
        qn, gn = oracle.l2_normalize(s.qf.numpy()), oracle.l2_normalize(s.gf.numpy())
        t0 = time.perf_counter()
        oracle.eval_func(oracle.euclidean_distance(qn, gn), s.q_pids, s.g_pids, s.q_camids, s.g_camids, sort_kind=None)
        t_plain = time.perf_counter() - t0
        t0 = time.perf_counter()
        final = oracle.re_ranking(qn, gn, 20, 6, 0.3)
        oracle.eval_func(final, s.q_pids, s.g_pids, s.q_camids, s.g_camids, sort_kind=None)
        t_rr = time.perf_counter() - t0
        res[key + "_eval"]["cpu_baseline"] = {"value": Q / t_plain, "unit": UNIT, "cores": os.cpu_count() or 1,
                                              "kind": "port", "sample": "whole workload once, %.2f s" % t_plain}
        res[key + "_rerank_k20_6"]["cpu_baseline"] = {
            "value": Q / t_rr, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "whole workload once (re_ranking(20, 6, 0.3) + eval_func), %.2f s" % t_rr}


def other_workloads(dev, peaks):
    """The remaining BASELINE.json configs on one GPU (device-resident inputs, CUDA-event time of
    the whole public-API call including the D2H of the metrics), each with the CPU port timed on
    the same inputs, the HBM-bound kernels against the measured copy bandwidth, and the triplet
    path next to the reference's own algorithm in eager torch on the same GPU."""
    import torch
    from demo2_b200 import metrics, reranking, synth, triplet_loss
    from oracle import reid_oracle as oracle

    spin_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    spin_b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)

    def timed(fn, iters=10, warm=3, busy_s=0.2):
        # Sub-millisecond workloads: ~40 ms of dense GEMM to raise the clocks after a host-side
        # pause, at least `warm` calls AND `busy_s` of the workload itself (with the allocation
        # pattern of the timed loop, see below), no interpreter GC inside the timed calls.  What
        # used to blow single entries up to 2 .. 11 ms was a cudaMalloc inside the timed loop.
        for _ in range(48):
            torch.mm(spin_a, spin_b)
        t0 = time.perf_counter()
        n = 0
        out = None
        while n < warm or time.perf_counter() - t0 < busy_s:
            out = fn()     # keep the previous result alive during the next call, exactly as the timed loop does:
            n += 1         # otherwise the first timed call needs a second buffer and pays a cudaMalloc (3 .. 40 ms)
            if n % 4 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()       # a full collection of the interpreter's heap takes 10 .. 100 ms: ten times one of these calls
        try:
            e0.record()
            for _ in range(iters):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
        finally:
            gc.enable()
        return e0.elapsed_time(e1) / iters, out

    hbm = peaks["hbm_gbs"]

    def roof(ms, nbytes, what):
        gbs = nbytes / ms * 1e-6
        return {"ms": ms, "algorithmic_bytes": int(nbytes), "achieved_gbs": gbs, "peak_gbs": hbm, "frac": gbs / hbm,
                "bytes": what}

    res = {}
    for key, cfg in (("rgbnt201", 1), ("rgbnt100", 2)):
        s = synth.make_named(key, sigma=4.0, seed=0)
        qf, gf = s.qf.to(dev), s.gf.to(dev)
        plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
        Q, G = qf.shape[0], gf.shape[0]
        N = Q + G
        ms, r = timed(lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))
        res[key + "_eval"] = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP),
                              "path": "label plan + distance GEMM + streaming rank count (as R1_mAP_eval.compute)"}
        ms, r = timed(lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True))
        res[key + "_eval_fused"] = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP),
                                    "max_positives_per_query": plan.max_cnt,
                                    "path": "fused rank-count GEMM epilogue (+ slab path for query blocks with more "
                                            "than 63 thresholds), plan reused"}

        def rr():
            dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
            return metrics.evaluate_matrix(dist, plan=plan)
        ms, r = timed(rr, iters=5)
        entry = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP),
                 "config": "BASELINE.json configs[%d]" % cfg}
        # the parameters R1_mAP_eval.compute hard-codes (utils/metrics.py:244: k1=50, k2=15)
        ms50, _ = timed(lambda: reranking.re_ranking_device(qf, gf, 50, 15, 0.3, normalize=True), iters=5)
        entry["re_ranking_k50_15_ms"] = ms50
        # re_ranking alone, eager and as ONE CUDA-graph replay (the ~14 launches of the call are
        # launch-latency-bound at this size; the library call itself has no host synchronisation)
        try:
            ms_e, eager = timed(lambda: reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True), iters=10)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
            graph.replay()
            torch.cuda.synchronize()
            ms_g, _ = timed(graph.replay, iters=20)
            entry["re_ranking_k20_6"] = {"eager_ms": ms_e, "graphed_ms": ms_g,
                                         "graphed_matches_eager": bool(torch.equal(static_out, eager))}
            del graph, static_out
        except Exception as exc:
            entry["re_ranking_k20_6"] = {"graphed_error": repr(exc)}
        # HBM-bound kernels of this config on their own (algorithmic bytes / CUDA-event time)
        allp = metrics.sqdist_device(torch.cat([qf, gf]), torch.cat([qf, gf]), normalize=True)
        ms_k, _ = timed(lambda: reranking.topk_rows(allp, 21), iters=20)
        entry["topk_rows"] = roof(ms_k, 4.0 * N * N, "4 B per entry of the %d x %d all-pairs matrix, read once" % (N, N))
        dist_m = allp[:Q, Q:].contiguous()
        ms_k, _ = timed(lambda: metrics.evaluate_matrix(dist_m, plan=plan), iters=20)
        entry["count_matrix(eval_func on the matrix)"] = roof(
            ms_k, 4.0 * Q * G, "4 B per (query, gallery) pair read once; whole call incl. records, thresholds, finalise, D2H")
        del allp, dist_m
        res[key + "_rerank_k20_6"] = entry

    # ---- a gallery with long positive lists: 1 M gallery, 5 850 ids (~171 images per id) ----
    try:
        res["large_r171"] = large_r171(dev, timed)
    except Exception as exc:
        res["large_r171"] = {"error": repr(exc)}

    # ---- config[4]: triplet hard mining, 3 modalities x [128, 768] ----
    xs, labels = synth.make_triplet_batch()
    xs = [x.to(dev).requires_grad_(True) for x in xs]
    labels = labels.to(dev)
    loss_fn = triplet_loss.TripletLoss()

    def fwd():
        return [loss_fn(x, labels)[0] for x in xs]

    def fwd_bwd():
        for x in xs:
            x.grad = None
        tot = sum(loss_fn(x, labels)[0] for x in xs)
        tot.backward()
        return tot
    ms_f, _ = timed(fwd, iters=50)
    ms_fb, _ = timed(fwd_bwd, iters=50)
    trip = {"fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "anchors_per_s_fwd": 384 / ms_f * 1e3,
            "config": "BASELINE.json configs[4]"}
    if hasattr(triplet_loss, "triplet_loss_multi"):
        def fwd_multi():
            return triplet_loss.triplet_loss_multi(xs, labels)[0]

        def fwd_bwd_multi():
            for x in xs:
                x.grad = None
            tot = triplet_loss.triplet_loss_multi(xs, labels)[0].sum()
            tot.backward()
            return tot
        trip["batched_fwd_ms"], _ = timed(fwd_multi, iters=50)
        trip["batched_fwd_bwd_ms"], _ = timed(fwd_bwd_multi, iters=50)
        # the same forward + backward captured in a CUDA graph (what a captured training step pays: the
        # two kernels, no Python / autograd dispatch); check=False: no status copy inside the capture
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())

            def fb_nocheck():
                tot = triplet_loss.triplet_loss_multi(xs, labels, check=False)[0].sum()
                tot.backward()
                return tot
            with torch.cuda.stream(side):
                for _ in range(3):
                    for x in xs:
                        x.grad = None
                    fb_nocheck()
            torch.cuda.current_stream().wait_stream(side)
            for x in xs:
                x.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = fb_nocheck()
            graph.replay()
            torch.cuda.synchronize()
            eager = triplet_loss.triplet_loss_multi(xs, labels, check=False)[0].sum()
            trip["graphed_fwd_bwd_ms"], _ = timed(graph.replay, iters=200)
            trip["graphed_matches_eager"] = bool(torch.equal(static_loss.detach(), eager.detach()))

            def f_nocheck():
                with torch.no_grad():
                    return triplet_loss.triplet_loss_multi(xs, labels, check=False)[0]
            gf = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                f_nocheck()
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(gf):
                f_nocheck()
            trip["graphed_fwd_ms"], _ = timed(gf.replay, iters=200)
        except Exception as exc:  # a capture problem must not cost the headline line
            trip["graphed_error"] = repr(exc)
    # the reference's own algorithm (layers/triplet_loss.py:16-31, 51-104, 121-135) in eager torch on this GPU
    ref_f, ref_fb = torch_eager_triplet(xs, labels, timed)
    trip["torch_eager_reference_algorithm"] = {"fwd_ms": ref_f, "fwd_bwd_ms": ref_fb,
                                               "note": "fp32, no autocast; the same three calls per step"}
    res["triplet_3x128x768"] = trip

    # context for the roofline: the library bf16 GEMM on the benchmark's own shape (K = 1536,
    # 20 000 rows against a 131 072-row slice of the gallery, bf16 output written), long enough to
    # run into the power cap like the count GEMM does
    M, N, K = 20000, 131072, 1536
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms, _ = timed(lambda: torch.matmul(a, b.t(), out=out), iters=150, warm=20)
    res["cublas_bf16_same_shape"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms * 1e-9, "shape": [M, N, K],
                                     "note": "one pass; the count GEMM executes three fp16 passes of this shape "
                                             "(compare with roofline.executed_tflops)"}
    return res


def large_r171(dev, timed):
    """The headline shape with 5 850 identities instead of 50 000: every query has ~150 valid
    positives, more than one threshold window of the count epilogue -> slab path."""
    import torch
    from demo2_b200 import metrics
    Q, G, d, nid, ncam, sigma = 20000, 1000000, 1536, 5850, 8, 4.0
    rng = np.random.default_rng(1)
    q_pid, g_pid = rng.integers(0, nid, Q), rng.integers(0, nid, G)
    q_cam, g_cam = rng.integers(0, ncam, Q), rng.integers(0, ncam, G)
    gen = torch.Generator(device=dev).manual_seed(5)
    centers = torch.randn(nid, d, device=dev, generator=gen)
    qf = centers[torch.from_numpy(q_pid).to(dev)] + sigma * torch.randn(Q, d, device=dev, generator=gen)
    gf = torch.empty(G, d, device=dev)
    gp_dev = torch.from_numpy(g_pid).to(dev)
    for s in range(0, G, 131072):
        e = min(G, s + 131072)
        gf[s:e] = centers[gp_dev[s:e]] + sigma * torch.randn(e - s, d, device=dev, generator=gen)
    plan = metrics.RankPlan(q_pid, g_pid, q_cam, g_cam)
    ms, r = timed(lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True), iters=3, warm=2)
    return {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP), "max_positives_per_query": plan.max_cnt,
            "shape": "20 000 x 1 000 000, d=1536, 5 850 ids",
            "path": "every 256-row query block is flagged: stored slab GEMM + streaming count (one GEMM pass "
                    "whatever the number of positives); compare with ms_per_step of the headline"}


def torch_eager_triplet(xs, labels, timed):
    """layers/triplet_loss.py:16-31 (euclidean_dist), :51-104 (hard_example_mining), :121-135
    (TripletLoss with margin=None -> SoftMarginLoss) restated with the same torch calls."""
    import torch

    def euclidean_dist(x, y):
        m, n = x.size(0), y.size(0)
        xx = torch.pow(x, 2).sum(1, keepdim=True).expand(m, n)
        yy = torch.pow(y, 2).sum(1, keepdim=True).expand(n, m).t()
        dist = xx + yy
        dist = dist - 2 * torch.matmul(x, y.t())
        return dist.clamp(min=1e-12).sqrt()

    def mining(dist_mat, lab):
        N = dist_mat.size(0)
        is_pos = lab.expand(N, N).eq(lab.expand(N, N).t())
        is_neg = lab.expand(N, N).ne(lab.expand(N, N).t())
        dist_ap, _ = torch.max(dist_mat[is_pos].contiguous().view(N, -1), 1, keepdim=True)
        dist_an, _ = torch.min(dist_mat[is_neg].contiguous().view(N, -1), 1, keepdim=True)
        return dist_ap.squeeze(1), dist_an.squeeze(1)

    soft = torch.nn.SoftMarginLoss()

    def loss_of(x):
        ap, an = mining(euclidean_dist(x, x), labels)
        y = an.new().resize_as_(an).fill_(1)
        return soft(an - ap, y)

    def fwd():
        return [loss_of(x) for x in xs]

    def fwd_bwd():
        for x in xs:
            x.grad = None
        tot = sum(loss_of(x) for x in xs)
        tot.backward()
        return tot
    f, _ = timed(fwd, iters=50)
    fb, _ = timed(fwd_bwd, iters=50)
    return f, fb


def cpu_baseline(name, q_host, g_host, q_pid, g_pid, q_cam, g_cam):
    """Oracle port of the reference's compute() on the fixed query chunk, host cores (same code
    path as `--impl reference`)."""
    blas_threads = use_all_host_threads()
    chunk = cpu_chunk_size(name)
    value, ms, done, sample, split = cpu_measure(name, q_host[:chunk].numpy(), g_host.numpy(), q_pid[:chunk], g_pid,
                                                 q_cam[:chunk], g_cam, steps=1, warmup=1, budget_s=30.0)
    return {"value": value, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "blas_threads": blas_threads,
            "sample": sample, "phase_split_s": split}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the small-workload timings")
    args = ap.parse_args()
    # stdout carries the ONE JSON line: libraries that write to file descriptor 1 meanwhile (NCCL
    # prints its version banner there) are pointed at stderr until the line is due
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    _print = builtins.print

    def print_line(*a, **k):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        _print(*a, **k)
        sys.stdout.flush()
        os.dup2(2, 1)

    global print
    print = print_line
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
