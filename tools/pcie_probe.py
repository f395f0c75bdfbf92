import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, label, nbytes):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{label}: {nbytes / dt / 1e9:.1f} GB/s")
t(lambda: d_a.copy_(h_in, non_blocking=True), "H2D 1 GiB", n)
t(lambda: h_out.copy_(d_b, non_blocking=True), "D2H 1 GiB", n)
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
t(both, "H2D + D2H concurrent (sum)", 2 * n)
