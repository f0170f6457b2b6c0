"""PCIe floor of the end-to-end step on N GPUs AT ONCE: every GPU moves its share of config D's bytes (H2D query
columns 12 B/query, D2H u64 offsets + u32 targets) between pinned host memory and the device, both directions
overlapped, all GPUs concurrently -- what no host-buffer join can beat on this box.
usage: python tools/pcie_floor_multi.py [n_gpus]"""
import sys, time, threading, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
NQ, HITS = 100_000_000, 654_775_326
h2d_total, d2h_total = 12 * NQ, 8 * NQ + 4 * HITS
bufs = []
for d in range(n):
    with torch.cuda.device(d):
        a, b = h2d_total // n, d2h_total // n
        bufs.append((torch.empty(a, dtype=torch.uint8).pin_memory(), torch.empty(b, dtype=torch.uint8).pin_memory(),
                     torch.empty(a, dtype=torch.uint8, device=f"cuda:{d}"), torch.empty(b, dtype=torch.uint8, device=f"cuda:{d}"),
                     torch.cuda.Stream(d), torch.cuda.Stream(d)))
def step():
    for d, (hi, ho, di, do, s1, s2) in enumerate(bufs):
        with torch.cuda.stream(s1): di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
def sync():
    for d in range(n): torch.cuda.synchronize(d)
for _ in range(2): step()
sync(); t0 = time.perf_counter()
reps = 5
for _ in range(reps): step()
sync(); t = (time.perf_counter() - t0) / reps
print(f"{n} GPUs, config D bytes split {n} ways: H2D {h2d_total/1e9:.2f} GB + D2H {d2h_total/1e9:.2f} GB per step, both directions "
      f"and all GPUs concurrently: {t*1e3:.1f} ms per step = {(h2d_total+d2h_total)/t/1e9:.1f} GB/s aggregate "
      f"-> floor of the e2e metric {NQ/t:.3g} queries/s")
