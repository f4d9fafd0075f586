"""NVLink peer-memory all-reduce (csrc/bd_peer.cu): argument checks on one GPU, and — when the box has two GPUs —
the torchrun check against NCCL (`scripts/peer_allreduce_check.py`: bit-identical sums on 2 ranks, identical on all
ranks, flag protocol over 200 calls, CUDA-graph replay)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_handle_argument_checks():
    from marl_gym_pybullet_drones_b200 import _native
    lib = _native.load()
    h = C.c_void_p()
    assert lib.bd_peer_create(0, 0, 1, 1024, C.byref(h)) != 0            # a world of one has no peers
    assert b"world" in lib.bd_peer_last_error()
    assert lib.bd_peer_create(0, 2, 2, 1024, C.byref(h)) != 0            # rank outside the world
    assert lib.bd_peer_create(0, 0, 2, 1001, C.byref(h)) == 0
    assert lib.bd_peer_handle_size() == 64
    buf = (C.c_ubyte * 64)()
    assert lib.bd_peer_get_handle(h, buf) == 0 and any(buf)
    # not opened yet: the collective refuses to launch instead of dereferencing unmapped peers
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.bd_peer_allreduce(h, 1001, None, 0, st) != 0
    assert b"not opened" in lib.bd_peer_last_error()
    assert lib.bd_peer_open(h, bytes(64), 1) != 0                        # one handle per rank
    assert int(lib.bd_peer_data(h)) % 16 == 0
    lib.bd_peer_destroy(h)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_allreduce_equals_nccl_on_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "scripts", "peer_allreduce_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "identical on all ranks; graph ok" in res.stdout
