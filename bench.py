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
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import builtins
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


def workload_desc(name):
    Q, G, d, nid, ncam, sigma = WORKLOADS[name]
    return ("%s: %d queries x %d gallery, d=%d fp32, %d ids, %d cams, sigma=%g, distance + ranking + CMC/mAP, "
            "no re-ranking (BASELINE.json configs[%d])" % (name, Q, G, d, nid, ncam, sigma,
                                                           {"large": 3, "rgbnt100": 2, "rgbnt201": 0}[name]))


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
# reference arm: the reference's CPU algorithm (oracle port; the reference is pure Python and
# /root/reference does not exist on the GPU box) on a bounded query sample
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


def cpu_eval_chunk(oracle, qf, gf_n, qp, gp, qc, gc):
    """compute()-equivalent of the reference on a query chunk (eval_func treats queries
    independently, so chunking is exact): normalise -> euclidean_distance -> eval_func."""
    qn = oracle.l2_normalize(qf)
    dist = oracle.euclidean_distance(qn, gf_n)
    return oracle.eval_func(dist, qp, gp, qc, gc, sort_kind=None)  # the reference's default argsort


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reid_oracle as oracle
    import torch
    blas_threads = use_all_host_threads()
    name = args.workload
    Q, G, d, nid, ncam, sigma = WORKLOADS[name]
    cores = os.cpu_count() or 1
    qf, gf, qp, gp, qc, gc = host_data(name, q_rows=min(Q, 512))
    gf_n = oracle.l2_normalize(gf)  # gallery normalisation amortised over the query chunks of one evaluation
    # calibrate the chunk so that the whole run stays within a few minutes
    n_cal = min(len(qf), 8)
    t0 = time.perf_counter()
    cpu_eval_chunk(oracle, qf[:n_cal], gf_n, qp[:n_cal], gp, qc[:n_cal], gc)
    per_q = (time.perf_counter() - t0) / n_cal
    budget = 150.0 / max(1, args.steps + args.warmup)
    chunk = int(max(4, min(len(qf), min(256, budget / max(per_q, 1e-9)))))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cpu_eval_chunk(oracle, qf[:chunk], gf_n, qp[:chunk], gp, qc[:chunk], gc)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = chunk / (ms * 1e-3)
    sample = ("%d-query chunk x full %d gallery per step (normalise + fp32 sgemm distance + np.argsort + "
              "per-query CMC/AP loop, oracle port of utils/metrics.py:110-169,341-401; gallery normalised once "
              "outside the timed step)" % (chunk, G))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(name)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "blas_threads": blas_threads},
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
    gen_g = torch.Generator(device=dev).manual_seed(99 + rank)
    gl = hi - lo
    gf = torch.empty(gl, d, device=dev)
    gp_dev = torch.from_numpy(g_pid[lo:hi]).to(dev)
    for s in range(0, gl, 131072):
        e = min(gl, s + 131072)
        gf[s:e] = centers[gp_dev[s:e]] + sigma * torch.randn(e - s, d, device=dev, generator=gen_g)
    del centers
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

    # ---- device-resident timing ----
    for _ in range(max(args.warmup, 3)):
        res = step(qf, gf, labels)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    timers = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        t = {}
        res = step(qf, gf, labels, timers=t)
        timers.append(t)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    value = Q / (ms_step * 1e-3)
    # dominant kernel: the fused count GEMM, CUDA-event time on its launch stream
    count_ms = float(np.mean([t["count"][0].elapsed_time(t["count"][1]) for t in timers]))
    launches = int(np.mean([t["launches"] for t in timers]))
    cms = torch.tensor([count_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(cms, op=dist.ReduceOp.MAX)
    count_ms = float(cms.item())
    algo_tflop = 2.0 * Q * gl * d * 1e-12
    passes = 3
    peak = peaks["bf16_tflops_sustained"] / passes
    achieved = algo_tflop / (count_ms * 1e-3)
    stage_ms = {k: float(np.mean([t[k][0].elapsed_time(t[k][1]) for t in timers]))
                for k in timers[0] if isinstance(timers[0][k], tuple)}

    # ---- end to end through the public API with HOST (pinned) buffers ----
    q_host = qf.cpu().pin_memory()
    g_host = gf.cpu().pin_memory()
    lab_host = {k: v.cpu().pin_memory() for k, v in labels.items()}
    del qf, gf
    torch.cuda.empty_cache()

    # Every step uploads ITS OWN inputs from pinned host memory and reads its result back to the
    # host.  Two device input slots: the upload of step i+1 is issued on a copy stream before step
    # i's evaluation is launched, so PCIe traffic overlaps the GEMM of the previous step
    # (streaming evaluation); `serial_ms_per_step` is the same without that overlap.
    def alloc_slot():
        return (torch.empty_like(q_host, device=dev), torch.empty_like(g_host, device=dev),
                {k: torch.empty_like(v, device=dev) for k, v in lab_host.items()})

    def upload(slot):
        slot[0].copy_(q_host, non_blocking=True)
        slot[1].copy_(g_host, non_blocking=True)
        for k, v in lab_host.items():
            slot[2][k].copy_(v, non_blocking=True)

    slots = [alloc_slot(), alloc_slot()]
    copy_stream = torch.cuda.Stream()
    ev_up = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_serial_step():
        upload(slots[0])
        return step(*slots[0])

    def e2e_pipelined(n):
        main = torch.cuda.current_stream()

        def issue_upload(i):
            copy_stream.wait_event(ev_done[i % 2])          # slot free again (its last evaluation finished)
            with torch.cuda.stream(copy_stream):
                upload(slots[i % 2])
                ev_up[i % 2].record(copy_stream)
        for e in ev_done:
            e.record(main)
        issue_upload(0)
        r = None
        for i in range(n):
            if i + 1 < n:
                issue_upload(i + 1)
            main.wait_event(ev_up[i % 2])
            r = step(*slots[i % 2])                          # ends with the D2H read of cmc / mAP
            ev_done[i % 2].record(main)
        return r

    for _ in range(2):
        e2e_serial_step()
    barrier()
    n_e2e = max(3, min(args.steps, 6))
    e0.record()
    for _ in range(2):
        res2 = e2e_serial_step()
    e1.record()
    barrier()
    serial_ms = e0.elapsed_time(e1) / 2
    e2e_pipelined(2)
    barrier()
    e0.record()
    res2 = e2e_pipelined(n_e2e)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1), serial_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2[0].item()) / n_e2e
    serial_ms = float(t2[1].item())
    h2d = (q_host.numel() + g_host.numel()) * 4 + sum(v.numel() * 4 for v in lab_host.values())
    d2h = 4096 * 4 + 8 + 16 + 32  # metrics slab (cmc | mAP | num_valid) + plan info

    traffic = load_traffic(name, world)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 hi/lo split on tcgen05, fp32 accumulate)",
                "data": "synthetic",
                "config": {"workload": workload_desc(name),
                           "sharding": "gallery rows split contiguously over ranks, queries replicated",
                           "l2": "inputs exceed L2 (gallery shard %.1f GB read per step)" % (gl * d * 4e-9),
                           "result": {"mAP": float(res.mAP), "rank1": float(res.cmc[0]), "num_valid": int(res.num_valid)}},
                "clocks": clocks,
                "e2e": {"value": Q / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "serial_ms_per_step": serial_ms, "serial_value": Q / (serial_ms * 1e-3),
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "api": "demo2_b200.parallel.ShardedEvaluator.evaluate; every step copies its own pinned host "
                               "features + labels to the device and reads cmc/mAP back; the upload of step i+1 is "
                               "issued on a copy stream while step i computes (serial_* = no overlap)"},
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
                             "burst_note": "short launches between host-synchronised exchange stages (N >= 4) run "
                                           "above the sustained clock; frac_vs_burst uses the measured burst bf16 peak"},
                "stage_ms": stage_ms}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(name, q_host, g_host, q_pid, g_pid, q_cam, g_cam)
        if world == 1 and not args.no_other:
            del q_host, g_host
            try:
                line["other_workloads"] = other_workloads(dev)
            except Exception as exc:  # never lose the headline line
                line["other_workloads"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_workloads(dev):
    """The remaining BASELINE.json configs on one GPU (device-resident inputs, CUDA-event time of
    the whole public-API call including the D2H of the metrics): latency-bound sizes, reported
    as ms and queries/s next to the headline."""
    import torch
    from demo2_b200 import metrics, reranking, synth, triplet_loss

    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, out

    res = {}
    for key, shape, sigma in (("rgbnt201", "rgbnt201", 4.0), ("rgbnt100", "rgbnt100", 4.0)):
        s = synth.make_named(shape, sigma=sigma, seed=0)
        qf, gf = s.qf.to(dev), s.gf.to(dev)
        plan = metrics.RankPlan(s.q_pids, s.g_pids, s.q_camids, s.g_camids)
        Q = qf.shape[0]
        ms, r = timed(lambda: metrics.evaluate_auto(qf, gf, s.q_pids, s.g_pids, s.q_camids, s.g_camids, normalize=True))
        res[key + "_eval"] = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP),
                              "path": "label plan + distance GEMM + streaming rank count (as R1_mAP_eval.compute)"}
        ms, r = timed(lambda: metrics.evaluate_features(qf, gf, plan=plan, normalize=True))
        res[key + "_eval_fused"] = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP),
                                    "path": "fused rank-count GEMM epilogue, plan reused"}

        def rr():
            dist = reranking.re_ranking_device(qf, gf, 20, 6, 0.3, normalize=True)
            return metrics.evaluate_matrix(dist, plan=plan)
        ms, r = timed(rr, iters=5)
        res[key + "_rerank_k20_6"] = {"ms": ms, "queries_per_s": Q / ms * 1e3, "mAP": float(r.mAP)}
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
    ms_f, _ = timed(fwd, iters=20)
    ms_fb, _ = timed(fwd_bwd, iters=20)
    res["triplet_3x128x768"] = {"fwd_ms": ms_f, "fwd_bwd_ms": ms_fb, "anchors_per_s_fwd": 384 / ms_f * 1e3}
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


def cpu_baseline(name, q_host, g_host, q_pid, g_pid, q_cam, g_cam):
    """Oracle port of the reference's compute() on a bounded query sample, host cores."""
    from oracle import reid_oracle as oracle
    blas_threads = use_all_host_threads()
    G = g_host.shape[0]
    chunk = 128 if G > 100000 else min(q_host.shape[0], 1024)   # ~10-15 s of host work on the large gallery
    qf = q_host[:chunk].numpy()
    t0 = time.perf_counter()
    gf_n = oracle.l2_normalize(g_host.numpy())
    t_norm = time.perf_counter() - t0
    t0 = time.perf_counter()
    cpu_eval_chunk(oracle, qf, gf_n, q_pid[:chunk], g_pid, q_cam[:chunk], g_cam)
    dt = time.perf_counter() - t0
    return {"value": chunk / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "blas_threads": blas_threads,
            "sample": "%d-query chunk x full %d gallery, %.2f s (gallery normalisation %.2f s not included; sgemm "
                      "multi-threaded, argsort and the per-query loop single-threaded as in the reference)"
                      % (chunk, G, dt, t_norm)}


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
