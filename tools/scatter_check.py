"""Fused MFCC + gather check for `torchrun --nproc-per-node N` (N >= 2 GPUs of one box): every rank runs the scatter form of the
MFCC kernel on its shard, writing into every rank's copy of the matrix through CUDA-IPC peer mappings; afterwards every rank must
hold exactly what kernel + NCCL all-gather give.  Prints per-rank times of both."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from jeicyboodsp_b200 import synth  # noqa: E402
from jeicyboodsp_b200.binding import Context, Library  # noqa: E402
from jeicyboodsp_b200.sharding import PeerMatrix, allgather_features, shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
U_total, n = int(sys.argv[1]) if len(sys.argv) > 1 else 4001, 160_000       # an odd total: shards differ by one utterance
L = Library()
ctx = Context(L, local, stream=torch.cuda.current_stream().cuda_stream)
lo, hi = shard_range(U_total, rank, world)
U = hi - lo
plan = ctx.mfcc_plan(L.mfcc_params("bench"))
nf = plan.n_frames(n)
x = synth.denoise_streams_torch(U, n, dev, stream0=lo, sigma=25.0, seed=4)
feat = torch.empty((U, nf, 13), dtype=torch.float32, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_kernel = timed(lambda: plan.run(x, n, U, n, feat, nf * 13))
ref = [None]


def nccl():
    plan.run(x, n, U, n, feat, nf * 13)
    ref[0] = allgather_features(feat, U_total)


ms_nccl = timed(nccl)
pm = PeerMatrix(ctx, U_total, nf * 13)
dests = pm.dests(lo)
tiny = torch.zeros(1, device=dev)


def fused():
    plan.run_scatter(x, n, U, n, dests, nf * 13)
    dist.all_reduce(tiny)


ms_fused = timed(fused)
torch.cuda.synchronize(); dist.barrier()
M = pm.tensor().view(U_total, nf, 13)
same = bool(torch.equal(M, ref[0]))
ok = torch.tensor([0.0 if same else 1.0], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MAX)
# ---- the same through ONE NVLink multicast address (torch symmetric memory binds every rank's copy to an NVSwitch multicast object) ----
from jeicyboodsp_b200.sharding import MulticastMatrix  # noqa: E402
mc_line = "multicast: not available on this platform"
try:
    mm = MulticastMatrix(U_total, nf * 13, dev)
    if mm.available():
        mcd = mm.dest(lo)

        def fused_mc():
            plan.run_multicast(x, n, U, n, mcd, nf * 13)
            dist.all_reduce(tiny)

        ms_mc = timed(fused_mc)
        torch.cuda.synchronize(); dist.barrier()
        same_mc = bool(torch.equal(mm.tensor().view(U_total, nf, 13), ref[0]))
        okm = torch.tensor([0.0 if same_mc else 1.0], device=dev)
        dist.all_reduce(okm, op=dist.ReduceOp.MAX)
        ok = torch.maximum(ok, okm)
        mc_line = f"fused multicast scatter (multimem.st, + 1-element all-reduce) {ms_mc:.3f} ms, identical on every rank: {okm.item() == 0.0}"
except Exception as e:  # noqa: BLE001 - platforms without symmetric memory / multicast
    mc_line = f"multicast: {type(e).__name__}: {e}"
if rank == 0:
    print(mc_line, flush=True)
if rank == 0:
    print(f"world {world}: {U_total} utterances x 10 s; kernel only {ms_kernel:.3f} ms, kernel + NCCL all-gather {ms_nccl:.3f} ms, "
          f"fused scatter (+ 1-element all-reduce) {ms_fused:.3f} ms; every rank's matrix identical to the NCCL result: {ok.item() == 0.0}", flush=True)
del M
pm.close()
plan.close()
dist.destroy_process_group()
sys.exit(0 if ok.item() == 0.0 else 1)
