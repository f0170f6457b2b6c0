#!/usr/bin/env python
"""bench.py -- the interval-overlap join on B200, BASELINE.json's metric on BASELINE.json's config.

  python bench.py [--gpus N --steps K --warmup W]            our arm (libbinary_cuda through the C ABI)
  python bench.py --impl reference [...]                     the reference's own CPU IntervalTree
  torchrun --nproc-per-node N ... bench.py --gpus N ...      one rank per GPU, weak scaling, no collective

A "step" is one pass of the hot path over one batch: the fused count -> prefix-sum -> scatter join of
the workload's queries against the prebuilt index (the index build is reported separately, as the
reference's insert phase is). `value` = whole-job queries/s with inputs resident in HBM; `e2e` = the
same metric through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="B", choices=["B", "C", "D"],
                    help="BASELINE.json configs[1]=B (headline), [2]=C dense, [3]=D scale-out")
    ap.add_argument("--queries", type=int, default=0, help="override queries per GPU (debug)")
    ap.add_argument("--targets", type=int, default=0, help="override target count (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU-baseline sample")
    return ap.parse_args()


def workload_for(args):
    from binary_b200 import synth
    w = synth.CONFIGS[args.workload]
    n_t = args.targets or w.n_targets
    # weak scaling: every GPU gets its own batch of the configuration's size (D: 100M split over 8)
    per_gpu = args.queries or (w.n_queries if args.workload != "D" else w.n_queries // 8)
    return w, n_t, per_gpu


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
        except OSError:
            self.proc = None

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
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture, or None."""
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
    """Queries the CPU reference is timed on: ~10-30 core-seconds of work. B (0.65 hits/query, ~1 us per
    query and core) runs the WHOLE batch; C (83 hits/query, ~500 us per query and core) and D are sampled."""
    return min(n_q, {"B": 10_000_000, "C": 100_000, "D": 2_000_000}[workload])


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
            "hit_pairs_per_sec": float(off[-1]) / query_s}, forest, orc


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, n_t, per_gpu = workload_for(args)
    tg, tl, th = w.targets(n_t)
    sample = args.cpu_sample or cpu_sample_size(args.workload, per_gpu)
    qg, ql, qh = w.queries(0, sample)
    base, forest, orc = cpu_reference_run(tg, tl, th, qg, ql, qh)
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
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w.name, "n_targets": n_t, "queries_per_step": sample,
                       "note": "CPU reference: each step is a bounded sample of the workload"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from binary_b200 import DeviceIndex, synth, _lib

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

    w, n_t, n_q = workload_for(args)
    tg, tl, th = w.targets(n_t)
    # shard = this rank's contiguous query range of the (counter-based) stream; no data-path collective
    # (binary_b200.sharding.shard_range over world*n_q queries: rank r gets [r*n_q, (r+1)*n_q))
    q_start = rank * n_q
    qg, ql, qh = w.queries(q_start, n_q)

    to_dev = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
    d_tg, d_tl, d_th = map(to_dev, (tg, tl, th))
    d_qg, d_ql, d_qh = map(to_dev, (qg, ql, qh))
    stream = torch.cuda.current_stream().cuda_stream

    # ---- index build (reported separately; the reference's insert phase) ----
    torch.cuda.synchronize()
    tb = time.perf_counter()
    ix = DeviceIndex.build_dev(n_t, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr(), device=local,
                               stream=stream)
    torch.cuda.synchronize()
    build_ms_first = (time.perf_counter() - tb) * 1e3
    build_ms = None
    for _ in range(2):  # steady-state rebuild: the second one reuses the pool memory of the first
        tb = time.perf_counter()
        ix2 = DeviceIndex.build_dev(n_t, d_tl.data_ptr(), d_th.data_ptr(), d_tg.data_ptr(), device=local,
                                    stream=stream)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - tb) * 1e3
        ix2.close()
    info = ix.info()

    # ---- output buffers: size the pair buffer with one count pass ----
    d_off = torch.empty(n_q + 1, dtype=torch.int64, device=dev)
    ix.count_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), d_qg.data_ptr(), stream)
    torch.cuda.synchronize()
    n_hits = int(d_off[-1].item())
    cap = n_hits + 1024
    d_hq = torch.empty(cap, dtype=torch.int32, device=dev)
    d_ht = torch.empty(cap, dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        ix.join_dev(n_q, d_ql.data_ptr(), d_qh.data_ptr(), d_off.data_ptr(), cap, d_hq.data_ptr(),
                    d_ht.data_ptr(), d_total.data_ptr(), d_qg.data_ptr(), q_start & 0xFFFFFFFF, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # runs through warm-up + timed region + an untimed tail of identical steps
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            time.sleep(0.05)  # nvidia-smi takes a moment to print its first sample
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    barrier()
    launches0 = lib.bcu_launch_count()
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall_s = time.perf_counter() - t_wall0
    launches = lib.bcu_launch_count() - launches0
    # the timed region is only ~10 ms: keep the same load running (untimed) so nvidia-smi sees it
    t_tail = time.perf_counter()
    while time.perf_counter() - t_tail < 0.6:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms_total = float(sum(step_ms))
    assert int(d_total.item()) == n_hits

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        def pinned(a, dtype):
            t = torch.empty(a if isinstance(a, int) else a.size, dtype=dtype).pin_memory()
            if not isinstance(a, int):
                t.numpy()[:] = a.view(np.int32)
            return t
        h_qg, h_ql, h_qh = pinned(qg, torch.int32), pinned(ql, torch.int32), pinned(qh, torch.int32)
        h_off = pinned(n_q + 1, torch.int64)
        h_hq, h_ht = pinned(cap, torch.int32), pinned(cap, torch.int32)
        total = C.c_uint64()

        def e2e_step():
            _lib.check(lib.bcu_join(ix._h, n_q, h_qg.data_ptr(), h_ql.data_ptr(), h_qh.data_ptr(),
                                    h_off.data_ptr(), cap, h_hq.data_ptr(), h_ht.data_ptr(), C.byref(total)))
        for _ in range(2):
            e2e_step()
        barrier()
        k_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / k_e2e
        assert total.value == n_hits and int(h_off[n_q].item()) == n_hits
        e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_q / float(e2e_t.item()), "unit": UNIT,
               "h2d_bytes_per_step": 12 * n_q, "d2h_bytes_per_step": 8 * (n_q + 1) + 8 * n_hits,
               "ms_per_step": float(e2e_t.item()) * 1e3, "api": "bcu_join (host buffers, pinned)"}

    # ---- max over ranks, whole-job aggregate ----
    t = torch.tensor([dev_ms_total, float(n_hits)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms_total, hits_all = float(tmax[0].item()), int(tsum[1].item())
    else:
        hits_all = n_hits
    ms_per_step = dev_ms_total / args.steps
    value = world * n_q / (ms_per_step * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = synth.algorithmic_bytes(n_q, n_t, n_hits)
        kern_ms = float(np.mean(step_ms))
        achieved = alg / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": w.name, "n_targets": n_t, "queries_per_gpu": n_q,
                       "hits_per_query": n_hits / n_q, "sharding": "replicated index, contiguous query range per GPU, no collective", "numa_cpus_per_rank": numa,
                       "l2": "flushed between timed steps (512 MiB write, not timed); step working set ~280 MB > 126 MB L2",
                       "timing": "CUDA events per step on the launching stream, summed; max over ranks",
                       "index": info},
            "hit_pairs_per_sec": hits_all / (ms_per_step * 1e-3),
            "index_build_ms": build_ms, "index_build_ms_first_call": build_ms_first,
            "wall_ms_per_step_incl_flush": wall_s / args.steps * 1e3,
            "roofline": {"bound": "hbm", "kernel": "bcu::probe_kernel + bcu::emit_kernel<true> + bcu::emit_long_kernel (one join step = 3 launches)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": kern_ms, "peak_source": peak_src,
                         "frac_of_nominal_8000": achieved / 8000.0},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        }
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample or cpu_sample_size(args.workload, n_q)
            base, forest, orc = cpu_reference_run(tg, tl, th, qg[:sample], ql[:sample], qh[:sample])
            line["cpu_baseline"] = base
        else:
            line["cpu_baseline"] = None
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
