#!/usr/bin/env python
"""bench.py -- the interval-overlap join on B200, BASELINE.json's metric on BASELINE.json's config.

  python bench.py [--gpus N --steps K --warmup W]            our arm (libbinary_cuda through the C ABI)
  python bench.py --impl reference [...]                     the reference's own CPU IntervalTree
  torchrun --nproc-per-node N ... bench.py --gpus N ...      one rank per GPU, STRONG scaling, no collective

Workload (default): BASELINE.json's north-star config D -- 10 M targets x 100 M unsorted queries, hg38 law.
At N = 1 the step is the WHOLE 100 M-query batch on one GPU; at N ranks the batch is split into N contiguous
query ranges (binary_b200.sharding.shard_range; the reference's split is one task per chromosome,
sv2nl mapper.hpp:238-246), every rank holds the replicated index, nothing is exchanged on the data path.
A "step" is one pass of the hot path over the rank's batch: the fused count -> prefix-sum -> scatter join
against the prebuilt index (the index build is reported separately, with its own roofline, as the reference's
insert phase is). `value` = whole-job queries/s with inputs resident in HBM; `e2e` = the same metric through
the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region). At N = 1 the line also
carries the other two synthetic configs as keyed sub-results (`also.B`, `also.C`). Prints ONE JSON line on
rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "overlap_queries_per_sec", "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="D", choices=["B", "C", "D"],
                    help="BASELINE.json configs[3]=D (north star, default), [1]=B sparse, [2]=C dense")
    ap.add_argument("--queries", type=int, default=0, help="override the batch size (debug)")
    ap.add_argument("--targets", type=int, default=0, help="override target count (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the B and C sub-results")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU-baseline sample")
    return ap.parse_args()


def workload_for(name, args):
    from binary_b200 import synth
    w = synth.CONFIGS[name]
    main = name == args.workload
    n_t = (args.targets if main else 0) or w.n_targets
    n_q = (args.queries if main else 0) or w.n_queries
    return w, n_t, n_q


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_wait = time.perf_counter()
            while not self.rows and time.perf_counter() - t_wait < 3.0:
                time.sleep(0.05)  # nvidia-smi takes a moment to print its first sample
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples taken while the GPU idles between launches read low; "under load" = upper half
        sm_sorted = sorted(sm)
        med = sm_sorted[len(sm_sorted) * 3 // 4] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str):
    """dram read+write bytes per step from the committed ncu capture of this same command
    (profiles/traffic.json, written by tools/summarize_profile.py from a `ncu --set full` pass), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh)[workload]["bytes_per_step"]
    except Exception:
        return None


def pin_to_gpu_numa_node(local: int):
    """N > 1: run this rank on the CPUs next to its GPU (NVML's affinity mask), so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's NUMA node instead of wherever the launcher put the
    process. Returns the number of CPUs in the mask, or None when NVML does not answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def cpu_sample_size(workload: str, n_q: int) -> int:
    """Queries the CPU reference is timed on: ~10-30 core-seconds of work. B
    (0.65 hits/query, ~1 us per query and core) runs the WHOLE batch; C (83 hits/query, ~500 us per query and
    core) its first 100 k and D its first 2 M queries."""
    return min(n_q, {"B": 10_000_000, "C": 100_000, "D": 2_000_000}[workload])


def pair_hash_torch(hq, ht):
    """Order-independent 64-bit hash of (query_id, target_id) pairs on the device: sum of mix64(q<<32|t)
    mod 2^64 (the same mix the oracle's pair hash uses; int64 arithmetic wraps). Chunked to bound memory."""
    import torch
    acc = 0
    step = 1 << 26
    for a in range(0, hq.numel(), step):
        q = hq[a:a + step].to(torch.int64) & 0xFFFFFFFF
        t = ht[a:a + step].to(torch.int64) & 0xFFFFFFFF
        z = (q << 32) | t
        z = (z ^ ((z >> 30) & 0x3FFFFFFFF)) * (-4658895280553007687)      # 0xBF58476D1CE4E5B9
        z = (z ^ ((z >> 27) & 0x1FFFFFFFFF)) * (-7723592293110705685)     # 0x94D049BB133111EB
        z = z ^ ((z >> 31) & 0x1FFFFFFFF)
        acc = (acc + int(z.sum().item())) & 0xFFFFFFFFFFFFFFFF
    return acc


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(tg, tl, th, qg, ql, qh, threads=0):
    """Time the reference's CPU path on this box's host cores: oracle/_ref (the unmodified reference
    headers) when the prebuilt library is present, else the C port. Returns a cpu_baseline dict."""
    import oracle
    kind = "reference" if oracle.have_reference() else "port"
    orc = oracle.Oracle(kind)
    threads = threads or orc.hardware_threads()
    t0 = time.perf_counter()
    forest = orc.build(tl, th, tg)
    build_s = time.perf_counter() - t0
    off, _, query_s = forest.query(ql, qh, qg, threads=threads, want_targets=True)
    return {"value": ql.size / query_s, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"first {ql.size} queries of the workload's batch vs all {tl.size} targets; "
                      f"find_overlaps from {threads} threads on one shared read-only forest",
            "build_s": round(build_s, 3), "query_s": round(query_s, 3),
            "hit_pairs_per_sec": float(off[-1]) / query_s}, forest, off


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, n_t, n_q = workload_for(args.workload, args)
    tg, tl, th = w.targets(n_t)
    sample = args.cpu_sample or cpu_sample_size(args.workload, n_q)
    qg, ql, qh = w.queries(0, sample)
    base, forest, _ = cpu_reference_run(tg, tl, th, qg, ql, qh)
    threads = base["cores"]
    times, pairs = [], 0
    for i in range(args.warmup + args.steps):
        off, _, s = forest.query(ql, qh, qg, threads=threads, want_targets=True)
        pairs = int(off[-1])
        if i >= args.warmup:
            times.append(s)
    t = sum(times) / len(times)
    value = sample / t
    base.update(value=value, query_s=round(t, 4), hit_pairs_per_sec=pairs / t)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w.name, "n_targets": n_t, "queries_total": n_q, "queries_per_step": sample,
                       "note": "CPU reference: each step is a bounded sample (the first queries) of the workload"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class DeviceJoin:
    """One workload resident on this rank's GPU: targets, this rank's query range, the index, outputs."""

    def __init__(self, name, args, rank, world, local, dev, stream):
        import torch
        from binary_b200 import DeviceIndex, synth
        from binary_b200.sharding import shard_range
        self.torch, self.synth = torch, synth
        self.w, self.n_t, self.n_total = workload_for(name, args)
        self.key = name
        self.q_start, q_stop = shard_range(self.n_total, rank, world)
        self.n_q = q_stop - self.q_start
        self.tg, self.tl, self.th = self.w.targets(self.n_t)
        self.qg, self.ql, self.qh = self.w.queries(self.q_start, self.n_q)
        to_dev = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
        self.d_tg, self.d_tl, self.d_th = map(to_dev, (self.tg, self.tl, self.th))
        self.d_qg, self.d_ql, self.d_qh = map(to_dev, (self.qg, self.ql, self.qh))
        self.dev, self.local, self.stream = dev, local, stream

        def build():
            return DeviceIndex.build_dev(self.n_t, self.d_tl.data_ptr(), self.d_th.data_ptr(),
                                         self.d_tg.data_ptr(), device=local, stream=stream)
        # ---- index build (reported separately; the reference's insert phase) ----
        torch.cuda.synchronize()
        tb = time.perf_counter()
        self.ix = build()
        torch.cuda.synchronize()
        self.build_ms_first = (time.perf_counter() - tb) * 1e3
        times = []
        for _ in range(3):  # steady-state rebuilds reuse the pool memory of the first
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix2 = build()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            ix2.close()
        self.build_ms = min(times)
        self.info = self.ix.info()
        # ---- output buffers: size the pair buffer with one count pass ----
        self.d_off = torch.empty(self.n_q + 1, dtype=torch.int64, device=dev)
        self.ix.count_dev(self.n_q, self.d_ql.data_ptr(), self.d_qh.data_ptr(), self.d_off.data_ptr(),
                          self.d_qg.data_ptr(), stream)
        torch.cuda.synchronize()
        self.n_hits = int(self.d_off[-1].item())
        self.cap = self.n_hits + 1024
        self.d_hq = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.d_ht = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self.d_total = torch.zeros(1, dtype=torch.int64, device=dev)

    def step(self):
        self.ix.join_dev(self.n_q, self.d_ql.data_ptr(), self.d_qh.data_ptr(), self.d_off.data_ptr(), self.cap,
                         self.d_hq.data_ptr(), self.d_ht.data_ptr(), self.d_total.data_ptr(), self.d_qg.data_ptr(),
                         self.q_start & 0xFFFFFFFF, self.stream)

    def timed(self, steps, warmup, flush, barrier):
        """W warm-up steps, then K steps each bracketed by CUDA events on the launching stream (L2 flushed
        before every step, outside the events). Returns the per-step milliseconds."""
        torch = self.torch
        for _ in range(warmup):
            flush.zero_()
            self.step()
        barrier()
        evs = []
        from binary_b200 import _lib
        launches0 = _lib.load().bcu_launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.step()
            e1.record()
            evs.append((e0, e1))
        barrier()
        self.wall_s = time.perf_counter() - t0
        self.launches = int(_lib.load().bcu_launch_count() - launches0)  # our kernels inside the timed region
        assert int(self.d_total.item()) == self.n_hits
        return [a.elapsed_time(b) for a, b in evs]

    def roofline(self, step_ms, peak, peak_src):
        alg = self.synth.algorithmic_bytes(self.n_q, self.n_t, self.n_hits)
        kern_ms = float(np.mean(step_ms))
        achieved = alg / (kern_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "one join step = all launches of bcu_join_dev (probe + emit kernels)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(self.key), "algorithmic_bytes_per_launch": alg, "kernel_ms": kern_ms,
                "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0}

    def build_roofline(self, peak):
        b = self.synth.algorithmic_build_bytes(self.n_t, self.info["n_groups"])
        return {"bytes": b, "ms": self.build_ms, "achieved": b / (self.build_ms * 1e-3) / 1e9,
                "frac": b / (self.build_ms * 1e-3) / 1e9 / peak, "ms_first_call": self.build_ms_first,
                "formula": "P*2*16*n_t + 12*n_t, P = ceil((32 + ceil(log2 groups)) / 8) radix passes (SURVEY 8d)",
                "timing": "CUDA events around bcu_index_build_dev (K1 sort + K2 index), best of 3 rebuilds"}

    def free(self):
        self.ix.close()
        for k in [k for k in vars(self) if k.startswith("d_")]:
            delattr(self, k)
        self.torch.cuda.empty_cache()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from binary_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the overlap join has no CPU fallback")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    warmup = max(args.warmup, 3)
    peak, peak_src = measured_peak()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    j = DeviceJoin(args.workload, args, rank, world, local, dev, stream)
    sampler = ClockSampler(local).start() if rank == 0 else None
    step_ms = j.timed(args.steps, warmup, flush, barrier)
    launches = j.launches
    # the timed region is short: keep the same load running (untimed) so nvidia-smi sees it
    t_tail = time.perf_counter()
    while time.perf_counter() - t_tail < 0.6:
        for _ in range(3):
            j.step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms_total = float(sum(step_ms))

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        def pinned(a, dtype):
            t = torch.empty(a if isinstance(a, int) else a.size, dtype=dtype).pin_memory()
            if not isinstance(a, int):
                t.numpy()[:] = a.view(np.int32)
            return t
        n_q, cap = j.n_q, j.cap
        h_qg, h_ql, h_qh = pinned(j.qg, torch.int32), pinned(j.ql, torch.int32), pinned(j.qh, torch.int32)
        h_cnt = pinned(n_q, torch.int32)
        h_ht = pinned(cap, torch.int32)
        total = C.c_uint64()
        one_index = (C.c_void_p * 1)(j.ix._h)

        def e2e_step():
            # the host-buffer join of this rank's query range: per-query hit counts as u32 (half the bytes of the u64
            # offsets over PCIe) and the target column; hit_query = NULL (redundant with the counts, binary_cuda.h).
            # bcu_join_multi over ONE index is bcu_join's chunk pipeline with the counts option.
            _lib.check(lib.bcu_join_multi(one_index, 1, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(), None,
                                          h_cnt.data_ptr(), cap, None, h_ht.data_ptr(), C.byref(total)))
        for _ in range(2):
            e2e_step()
        barrier()
        k_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / k_e2e
        assert total.value == j.n_hits
        # parity of the e2e path against the device-resident run, once, after timing: same per-query counts, same
        # pair multiset (order-independent hash; both computed on the device)
        counts = (j.d_off[1:] - j.d_off[:-1])
        off_ok = bool(torch.equal(h_cnt.to(dev).to(torch.int64), counts))
        d_back = h_ht[:j.n_hits].to(dev)
        qid = torch.repeat_interleave(torch.arange(n_q, device=dev, dtype=torch.int64), counts) + j.q_start
        h_e2e = pair_hash_torch(qid, d_back)
        h_dev = pair_hash_torch(j.d_hq[:j.n_hits], j.d_ht[:j.n_hits])
        del d_back, qid, counts
        assert off_ok and h_e2e == h_dev, "the host-buffer join disagrees with bcu_join_dev"
        # what this host can move at best for the same bytes: the raw copies alone, both directions at once, all ranks
        # at the same time (no kernels, no dependencies) -- the floor of any e2e number on this box
        s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        d_in = torch.empty(3 * n_q, dtype=torch.int32, device=dev)
        d_res = torch.empty(n_q + j.n_hits, dtype=torch.int32, device=dev)
        floor_s = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                for k, h in enumerate((h_qg, h_ql, h_qh)):
                    d_in[k * n_q:(k + 1) * n_q].copy_(h, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_cnt.copy_(d_res[:n_q], non_blocking=True)
                h_ht[:j.n_hits].copy_(d_res[n_q:], non_blocking=True)
            torch.cuda.synchronize()
            floor_s.append(time.perf_counter() - t0)
        del d_in, d_res
        e2e_t = torch.tensor([e2e_s, min(floor_s[1:])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": j.n_total / float(e2e_t[0].item()), "unit": UNIT,
               "h2d_bytes_per_step": 12 * n_q, "d2h_bytes_per_step": 4 * n_q + 4 * j.n_hits,
               "ms_per_step": float(e2e_t[0].item()) * 1e3,
               "pcie_floor_ms": float(e2e_t[1].item()) * 1e3,
               "pcie_floor": "the same H2D and D2H bytes as plain concurrent copies from/to the same pinned buffers, "
                             "all ranks at once, max over ranks: e2e cannot be faster than this on this host",
               "api": "bcu_join_multi over this rank's index (host buffers, pinned; u32 counts, hit_query = NULL)",
               "checked": "per-query counts equal and pair hash equal to the device-resident result"}
        del h_qg, h_ql, h_qh, h_cnt, h_ht

    # ---- max over ranks, whole-job aggregate ----
    t = torch.tensor([dev_ms_total, float(j.n_hits)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms_total, hits_all = float(tmax[0].item()), int(tsum[1].item())
    else:
        hits_all = j.n_hits
    ms_per_step = dev_ms_total / args.steps
    value = j.n_total / (ms_per_step * 1e-3)
    if e2e is not None:  # whole-job bytes (all ranks), like `value`
        e2e["h2d_bytes_per_step"] = 12 * j.n_total
        e2e["d2h_bytes_per_step"] = 4 * j.n_total + 4 * hits_all

    line = None
    if rank == 0:
        roof = j.roofline(step_ms, peak, peak_src)
        if world > 1:  # the job's algorithmic bytes over the slowest rank's time, against N x peak
            alg_all = j.synth.algorithmic_bytes(j.n_total, j.n_t * world, hits_all)
            roof.update(achieved=alg_all / (ms_per_step * 1e-3) / 1e9, peak=peak * world,
                        frac=alg_all / (ms_per_step * 1e-3) / 1e9 / (peak * world),
                        algorithmic_bytes_per_launch=alg_all, kernel_ms=ms_per_step,
                        note="whole job: all ranks' bytes / max-over-ranks step time, peak = N x one GPU")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": j.w.name, "n_targets": j.n_t, "queries_total": j.n_total,
                       "queries_per_gpu": j.n_q, "hits_per_query": j.n_hits / max(j.n_q, 1),
                       "sharding": "replicated index, contiguous query range per GPU (shard_range), no collective",
                       "numa_cpus_per_rank": numa,
                       "l2": "flushed between timed steps (512 MiB write, not timed); inputs + outputs >> 126 MB L2",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks",
                       "index": j.info},
            "hit_pairs_per_sec": hits_all / (ms_per_step * 1e-3),
            "index_build_ms": j.build_ms, "build": j.build_roofline(peak),
            "wall_ms_per_step_incl_flush": j.wall_s / args.steps * 1e3,
            "roofline": roof, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        }
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample or cpu_sample_size(args.workload, j.n_q)
            base, forest, off = cpu_reference_run(j.tg, j.tl, j.th, j.qg[:sample], j.ql[:sample], j.qh[:sample])
            base["parity_on_sample"] = bool(np.array_equal(
                off.astype(np.int64), j.d_off[:sample + 1].cpu().numpy()))  # the reference's CSR offsets
            line["cpu_baseline"] = base
            del forest
        else:
            line["cpu_baseline"] = None
    j.free()

    # ---- the other synthetic configs as keyed sub-results (N = 1 only) ----
    if world == 1 and not args.no_also and rank == 0:
        line["also"] = {}
        for name in [k for k in ("B", "C") if k != args.workload]:
            s = DeviceJoin(name, args, 0, 1, local, dev, stream)
            smp = ClockSampler(local).start()
            ms = s.timed(max(5, args.steps), warmup, flush, barrier)
            clk = smp.stop()
            r = s.roofline(ms, peak, peak_src)
            line["also"][name] = {"workload": s.w.name, "value": s.n_q / (float(np.mean(ms)) * 1e-3), "unit": UNIT,
                                  "ms_per_step": float(np.mean(ms)), "hits_per_query": s.n_hits / s.n_q,
                                  "hit_pairs_per_sec": s.n_hits / (float(np.mean(ms)) * 1e-3),
                                  "roofline": r, "build": s.build_roofline(peak), "clocks": clk, "index": s.info}
            s.free()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
