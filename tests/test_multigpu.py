"""The fused MFCC scatter (kernel writes into every rank's copy of the feature matrix; no all-gather).
* two shards in ONE process on both backends: each "rank" scatters its shard into both copies -> both copies == the unsharded result;
* two processes on two GPUs (torchrun, CUDA-IPC peer mappings over NVLink) against kernel + NCCL all-gather: skipped below 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from jeicyboodsp_b200 import synth
from jeicyboodsp_b200.sharding import shard_range
from backends import EmulBackend, GpuBackend  # noqa: E402  (tests/ is on sys.path via conftest)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}


@pytest.fixture(params=["emul", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    if request.param not in _CACHE:
        _CACHE[request.param] = EmulBackend() if request.param == "emul" else GpuBackend()
    return _CACHE[request.param]


def test_two_shards_scatter_into_both_copies(be):
    p = be.L.mfcc_params("bench")
    n, total = (6000, 5) if be.name == "emul" else (40000, 9)         # uneven shards: 3 + 2 / 5 + 4
    x = np.stack([synth.mfcc_utterance(70 + u, n) for u in range(total)])
    plan = be.ctx.mfcc_plan(p)
    nf = plan.n_frames(n)
    row = nf * 13
    whole = be.zeros((total, row), np.float32)
    assert plan.run(be.to_dev(x), n, total, n, whole, row) == nf
    copies = [be.zeros((total, row), np.float32) for _ in range(2)]     # rank 0's and rank 1's copy of the matrix
    for rank in range(2):
        lo, hi = shard_range(total, rank, 2)
        own, peer = copies[rank], copies[1 - rank]
        assert plan.run_scatter(be.to_dev(x[lo:hi]), n, hi - lo, n, [own[lo:], peer[lo:]], row) == nf
    want = be.to_host(whole)
    for c in copies:
        assert np.array_equal(be.to_host(c), want)
    plan.close()


def test_peer_matrix_single_rank_on_the_emulator():
    """sharding.PeerMatrix with no process group (world 1): allocation through the C ABI, the destination address of a unit, the scatter
    form writing into it, teardown.  (The emulator's device memory is host memory, so the result is read back through ctypes.)"""
    import ctypes as C

    from jeicyboodsp_b200.sharding import PeerMatrix
    be = _CACHE.setdefault("emul", EmulBackend())
    p = be.L.mfcc_params("bench")
    n, total, u0 = 6000, 4, 1
    plan = be.ctx.mfcc_plan(p)
    nf = plan.n_frames(n)
    row = nf * 13
    x = np.stack([synth.mfcc_utterance(90 + u, n) for u in range(2)])
    plain = be.zeros((2, row), np.float32)
    plan.run(be.to_dev(x), n, 2, n, plain, row)
    pm = PeerMatrix(be.ctx, total, row)
    assert pm.world == 1 and len(pm.addrs) == 1 and pm.dests(u0) == [pm.local + u0 * row * 4]
    C.memset(pm.local, 0, pm.nbytes)
    assert plan.run_scatter(be.to_dev(x), n, 2, n, pm.dests(u0), row) == nf
    got = np.ctypeslib.as_array((C.c_float * (total * row)).from_address(pm.local)).reshape(total, row).copy()
    assert np.array_equal(got[u0:u0 + 2], plain) and not got[:u0].any() and not got[u0 + 2:].any()
    pm.close()
    plan.close()


@pytest.mark.gpu
def test_fused_scatter_two_gpus_equals_nccl_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs of one box")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "scatter_check.py"), "301"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical to the NCCL result: True" in r.stdout
