#!/usr/bin/env python
"""bench.py — queries/sec of the posting-decode -> AND/OR -> BM25 -> top-k hot path on B200.

One "step" = one batch of synthetic queries through the whole hot path.  Default workload =
BASELINE.json configs[2], the configuration the metric is quoted on (100M-doc index):
5-term OR (Zipf-drawn ranks from {1..1000}), top-100, 100M docs in 8 segments, on ONE B200.
With --gpus N (torchrun, one rank per GPU) the SAME index is sharded by segment over the ranks
(strong scaling); every step ends with an NCCL all-gather of the ranks' top-k rows and a device
merge (merge_fruits).

value : whole-job queries/s with the batch descriptors already resident in HBM (kernels only).
e2e   : the same through tq_search_batch with HOST buffers (H2D descriptors + D2H results inside
        the timed region) — the number to compare with the reference arm.
--impl reference : the reference's CPU algorithm (oracle/ restatement of tantivy's Block-WAND path;
        the Rust crate itself cannot be built here) on all host cores, same workload/metric.
"""
import argparse
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
    # name: (op, n_segments, docs_per_segment, k, description)
    "or5_top100_100M_8seg": dict(op="or", n_segments=8, docs_per_segment=12_500_000, k=100, n_terms=5, max_rank=1000,
                                 desc="BASELINE.json configs[2]: 5-term OR, top-100, 100M docs, 8 segments"),
    "and2_top10_10M_1seg": dict(op="and", n_segments=1, docs_per_segment=10_000_000, k=10, n_terms=2, max_rank=0,
                                desc="BASELINE.json configs[1]: 2-term AND, top-10, 10M docs, 1 segment"),
    "term_top10_1M_1seg": dict(op="term", n_segments=1, docs_per_segment=1_000_000, k=10, n_terms=1, max_rank=0,
                               desc="BASELINE.json configs[0]: single-term top-10, 1M docs, 1 segment"),
    "mixed_top10_100M_8seg": dict(op="mixed", n_segments=8, docs_per_segment=12_500_000, k=10, n_terms=0, max_rank=1000,
                                  desc="BASELINE.json configs[3] shape: 40% 2-term AND, 40% 2-4-term OR, 20% term, top-10"),
}
AND_PAIRS = [(0.10, 0.10), (0.50, 0.02), (0.80, 0.005)]     # benches/intersection_bench.rs:107-113
TERM_LADDER = [0.0001, 0.01, 0.05, 0.15, 0.30]              # benches/and_or_queries.rs:134-140
ZIPF_C = 0.30                                               # rank 1 ~ 30 %


def zipf_density(rank):
    return min(0.5, ZIPF_C / rank)


def build_query_plan(wl, nq, n_batches, seed):
    """Returns (densities list, batches) where a batch is a list of (op, [term indices], k)."""
    rng = np.random.default_rng(seed)
    dens, index_of = [], {}

    def term(p):
        key = round(p, 12)
        if key not in index_of:
            index_of[key] = len(dens)
            dens.append(p)
        return index_of[key]

    if wl["max_rank"]:
        ranks = np.arange(1, wl["max_rank"] + 1)
        prob = (1.0 / ranks) / (1.0 / ranks).sum()
    batches = []
    for _ in range(n_batches):
        qs = []
        for _ in range(nq):
            op = wl["op"]
            if op == "mixed":
                u = rng.random()
                op = "and" if u < 0.4 else ("or" if u < 0.8 else "term")
                nt = 2 if op == "and" else (int(rng.integers(2, 5)) if op == "or" else 1)
            else:
                nt = wl["n_terms"]
            if wl["max_rank"]:
                rs = rng.choice(ranks, size=nt, replace=False, p=prob)
                terms = [term(zipf_density(int(r))) for r in rs]
            elif op == "and":
                a, b = AND_PAIRS[int(rng.integers(0, len(AND_PAIRS)))]
                terms = [term(a), term(b * (1 + 1e-9))]  # distinct lists even when a == b
            else:
                terms = [term(TERM_LADDER[int(rng.integers(0, len(TERM_LADDER)))])]
            qs.append((op, terms, wl["k"]))
        batches.append(qs)
    return dens, batches


OPS = {"term": 0, "and": 1, "or": 2}


def make_shard(wl, dens, rank, world, seed, dist=None, device=None):
    """This rank's segments of the synthetic index (segment s always has seed base+s) + global statistics."""
    import tantivy_b200 as T
    from tantivy_b200.sharding import ShardedIndex, assign_segments
    ords = assign_segments(wl["n_segments"], world, rank)
    t0 = time.time()
    ix = T.SynthIndex(len(ords), wl["docs_per_segment"], dens, seed=seed, segment_base=rank, segment_stride=world) if ords else None
    shard = ShardedIndex(ix, ords, len(dens), dist, device)
    shard.gen_s = time.time() - t0
    return shard


def marshal(shard, queries):
    return shard.marshal([(OPS[op], terms, k) for op, terms, k in queries])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(wl, shard, batches, steps, warmup, sample_queries, threads):
    """The reference CPU algorithm (oracle restatement, Block-WAND + TopNHeap + merge_top_k), all host
    cores, on a bounded sample of the same query stream. One step = `sample_queries` queries."""
    from oracle import tq_oracle as O
    oi = O.OracleIndex()
    shard.register(oi)
    flat = [q for b in batches for q in b]
    times = []
    for i in range(warmup + steps):
        qs = [flat[(i * sample_queries + j) % len(flat)] for j in range(sample_queries)]
        qb = marshal(shard, qs)
        t0 = time.perf_counter()
        oi.search_batch(qb, mode=1, n_threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return sample_queries * len(times) / total, 1000.0 * total / len(times)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="or5_top100_100M_8seg", choices=sorted(WORKLOADS))
    ap.add_argument("--nq", type=int, default=512, help="queries per step (batch)")
    ap.add_argument("--batches", type=int, default=4, help="distinct query batches cycled through the steps")
    ap.add_argument("--docs-per-segment", type=int, default=0, help="override (smoke runs); 0 = the workload's size")
    ap.add_argument("--seed", type=int, default=0x7A6E7469)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    wl = dict(WORKLOADS[args.workload])
    if args.docs_per_segment:
        wl["docs_per_segment"] = args.docs_per_segment
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    host_threads = os.cpu_count() or 1
    dens, batches = build_query_plan(wl, args.nq, args.batches, args.seed)
    config = {"workload": args.workload, "desc": wl["desc"], "queries_per_step": args.nq, "docs": wl["n_segments"] * wl["docs_per_segment"],
              "segments": wl["n_segments"], "k": wl["k"], "vocab_terms_materialised": len(dens),
              "sharding": f"segments round-robin over {world} rank(s)", "l2_policy": "inputs larger than L2 (see index_bytes/step_bytes)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        shard = make_shard(wl, dens, 0, 1, args.seed)
        # several queries per host thread, so that the threads stay busy while the heavy queries of the sample finish
        sample = args.cpu_sample or min(512, max(64, 4 * host_threads))
        qps, ms = cpu_reference_run(wl, shard, batches, args.steps, args.warmup, sample, host_threads)
        line = {"metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "impl": "reference", "config": config,
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": host_threads, "kind": "port",
                                 "sample": f"{sample} queries per step of the same query stream, all host threads, "
                                           "Block-WAND + TopNHeap + merge_top_k restatement (oracle/, mode=1)"},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    import tantivy_b200 as T
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier_sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    shard = make_shard(wl, dens, rank, world, args.seed, dist, dev)
    ctx = T.Context(local_rank)
    shard.register(ctx)
    qbs = [marshal(shard, b) for b in batches]
    k = wl["k"]
    nq = args.nq
    n_total = args.warmup + args.steps

    cross_gpu_merge = None
    if world > 1:
        from tantivy_b200.sharding import CrossGpuMerger
        cross_gpu_merge = CrossGpuMerger(ctx, dist, dev, nq, k)  # NCCL all-gather + device merge (K7)

    # ---- leg 1: `value` — descriptors resident, kernels only -----------------------------------------
    prepared = [ctx.prepare(qbs[i % len(qbs)]) for i in range(n_total)]  # also warms the block-table cache
    barrier_sync()
    for i in range(args.warmup):
        if world > 1:
            cross_gpu_merge.run(prepared[i])
            cross_gpu_merge(prepared[i])
        else:
            prepared[i].run()
    barrier_sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # per-launch device times of the timed steps: CUDA events recorded by the library on the stream each
    # kernel is launched on; steps are serialised (sync per step) so that the intervals do not overlap.
    kern = {"term_ms": [], "and_ms": [], "or_ms": [], "final_ms": [], "kernel_ms": []}
    launches = 0
    stats = None
    touched0 = ctx.stats()["or_windows"][5]  # cumulative bytes the pruned union kernel actually read
    t0 = time.perf_counter()
    for i in range(args.warmup, n_total):
        if world > 1:
            cross_gpu_merge.run(prepared[i])  # threshold exchange between the sample and the main pass
            cross_gpu_merge(prepared[i])
        else:
            prepared[i].run()
            prepared[i].results_dev()  # waits for the step
        stats = ctx.stats()
        for key in kern:
            kern[key].append(stats[key])
        launches += stats["kernel_launches"] + (7 if world > 1 else 0)  # + 3 x threshold export / import, cross-GPU merge
    barrier_sync()
    dt_value = time.perf_counter() - t0
    touched_per_step = (stats["or_windows"][5] - touched0) / max(args.steps, 1) if stats else 0
    clocks = sampler.stop()
    for b in prepared:
        b.close()

    # ---- leg 2: `e2e` — public API with host buffers ---------------------------------------------------
    outs = [qbs[i % len(qbs)].alloc_out() for i in range(2)]
    h2d = d2h = 0
    barrier_sync()
    for i in range(args.warmup):
        ctx.search_batch(qbs[i % len(qbs)], outs[i % 2])
    barrier_sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        if world == 1:
            ctx.search_batch(qbs[i % len(qbs)], outs[i % 2])
            st = ctx.stats()
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
        else:
            bt = ctx.prepare(qbs[i % len(qbs)])
            cross_gpu_merge.run(bt)
            o = cross_gpu_merge(bt)
            st = ctx.stats()
            h2d, d2h = st["h2d_bytes"], 0
            if rank == 0:
                res = [t.cpu() for t in o]
                d2h = sum(t.numel() * t.element_size() for t in res)
            bt.close()
    barrier_sync()
    dt_e2e = time.perf_counter() - t0

    # max over ranks
    if dist is not None:
        t = torch.tensor([dt_value, dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_value, dt_e2e = float(t[0]), float(t[1])

    if rank == 0:
        value = nq * args.steps / dt_value
        e2e = nq * args.steps / dt_e2e
        # roofline of the dominant kernel
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        op_ms = {name: float(np.mean(kern[name + "_ms"])) for name in ("term", "and", "or")}
        dominant = max(op_ms, key=op_ms.get)
        alg_bytes = stats["bytes_" + dominant]  # SURVEY.md §8(d), exhaustive form: every posting of every clause
        kernel_name = "k_" + dominant
        touched = None
        if dominant == "or" and stats.get("units_or_strip", 0) * 2 > stats["units_or"]:
            kernel_name = "k_or_strip"
            # pruned kernel (MaxScore): §8(d) asks for the formula restricted to the blocks actually decoded -- counted by
            # the kernel itself (packed block bytes + staged / gathered fieldnorm bytes) -- next to the exhaustive figure
            touched = float(touched_per_step) + 12.0 * wl["k"] * nq
        read_bytes = touched if touched else float(alg_bytes)
        basis = "bytes the kernel decoded (device counter)" if touched else ("exhaustive formula; k_and prunes leader docs, so this is an exhaustive-equivalent figure" if dominant == "and" else "exhaustive formula (the kernel reads every posting)")
        achieved = read_bytes / (op_ms[dominant] * 1e-3) / 1e9 if op_ms[dominant] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload, {}).get(kernel_name)
        roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": read_bytes, "achieved_basis": basis, "avg_launch_ms": op_ms[dominant],
                    "launches_per_step_of_this_kernel": 4 if kernel_name == "k_or_strip" else 1,  # 3 threshold rounds + the rest
                    "exhaustive_algorithmic_bytes_per_step": alg_bytes,
                    "exhaustive_equivalent_gbs": alg_bytes / (op_ms[dominant] * 1e-3) / 1e9 if op_ms[dominant] > 0 else 0.0,
                    "postings_per_launch": stats["postings"], "kernel_ms_per_step": {k2: float(np.mean(v)) for k2, v in kern.items()}}
        config.update({"index_bytes_this_rank": shard.index_bytes, "step_bytes_algorithmic": stats["algorithmic_bytes"],
                       "units_per_step": stats["units"], "index_generation_s": round(shard.gen_s, 2), "host_threads": host_threads})
        line = {"metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * dt_value / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "roofline": roofline, "clocks": clocks,
                "e2e": {"value": e2e, "unit": "queries/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": 1000.0 * dt_e2e / args.steps},
                "gpu_launches": int(launches)}
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or min(512, max(64, 4 * host_threads))
            # bounded: a few seconds per step; several queries per host thread keep the threads busy
            qps, ms = cpu_reference_run(wl, shard, batches, 2, 1, sample, host_threads)
            line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": host_threads, "kind": "port",
                                    "sample": f"3 x {sample} queries of the same stream (1 warm-up), all host threads, oracle/ "
                                              "restatement of tantivy's Block-WAND + TopNHeap + merge_top_k (mode=1)",
                                    "ms_per_sample": ms}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
