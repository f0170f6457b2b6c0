"""PCIe floor of the e2e step: H2D of the query columns and D2H of offsets + pairs, alone and overlapped."""
import time, torch
n_h2d, n_d2h = 120_000_000, 132_000_000
h_in = torch.empty(n_h2d, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_d2h, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_h2d, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_d2h, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 120 MB: {a*1e3:.3f} ms ({n_h2d/a/1e9:.1f} GB/s)  D2H 132 MB: {b*1e3:.3f} ms ({n_d2h/b/1e9:.1f} GB/s)  both overlapped: {c*1e3:.3f} ms")
