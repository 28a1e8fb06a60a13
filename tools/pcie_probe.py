import torch, time
n = 16 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for _ in range(2):
    a = run(1, 0); b = run(0, 1); c = run(1, 1)
    print(f"16 MiB: H2D {n/a/1e9:.1f} GB/s  D2H {n/b/1e9:.1f} GB/s  both: {c*1e3:.3f} ms = {2*n/c/1e9:.1f} GB/s total (serial would be {(a+b)*1e3:.3f} ms)")
# small pieces
for sz in (256 << 10, 1 << 20, 4 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        with torch.cuda.stream(s1): d_in[:sz].copy_(h_in[:sz], non_blocking=True)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
    print(f"H2D {sz>>10} KiB: {t*1e6:.1f} us = {sz/t/1e9:.1f} GB/s")
