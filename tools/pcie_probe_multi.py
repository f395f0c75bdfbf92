"""How much pinned-host <-> device traffic the box sustains when several GPUs copy at once (one process per GPU)."""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    dev = int(sys.argv[2]); torch.cuda.set_device(dev)
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def both():
        with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    both(); torch.cuda.synchronize()
    # crude start alignment: wait for a wall-clock tick shared by all workers
    t_go = float(sys.argv[3])
    while time.time() < t_go: pass
    t0 = time.perf_counter()
    for _ in range(8): both()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"gpu {dev}: {8 * 2 * n / dt / 1e9:.1f} GB/s (H2D + D2H)", flush=True)
else:
    import torch
    ng = torch.cuda.device_count()
    print("numa nodes:", os.listdir("/sys/devices/system/node") if os.path.isdir("/sys/devices/system/node") else "n/a", "cpus:", os.cpu_count())
    for k in sorted({1, min(2, ng), ng}):
        t_go = time.time() + 25
        ps = [subprocess.Popen([sys.executable, __file__, "worker", str(d), str(t_go)], stdout=subprocess.PIPE, text=True) for d in range(k)]
        outs = [p.communicate()[0].strip() for p in ps]
        tot = sum(float(o.split(":")[1].split()[0]) for o in outs if o)
        print(f"{k} GPU(s) at once: total {tot:.1f} GB/s | " + " | ".join(outs), flush=True)
