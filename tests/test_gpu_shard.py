"""one registration sharded by queries over 2 GPUs is bit-identical to the unsharded call and to the
oracle.  Needs two GPUs on the box (skipped otherwise); the processes are launched the way bench_shard.py is."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_query_sharded_registration_two_gpus_bit_equal():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench_shard.py"), "--config", "2", "--steps", "6", "--warmup", "2", "--oracle"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2
    assert d["parity"]["bit_equal_to_unsharded_on_every_rank"] is True
    assert d["parity"]["registrations_compared"] == 6
    assert d["parity"]["first_bit_equal_to_oracle"] is True


def test_shard_exchange_argument_checks_without_gpu():
    """The C ABI rejects nonsense before touching the device (no GPU needed: null context)."""
    import ctypes
    import laser_slam_b200 as ls
    L = ls.lib()
    h = (ctypes.c_ubyte * 64)()
    assert L.ls_shard_exchange_create(None, 0, 2, h) == ls.LS_ERR_ARG
    assert L.ls_shard_exchange_connect(None, h) == ls.LS_ERR_ARG
    assert L.ls_host_is_pinned(None) == ls.LS_ERR_ARG
