"""Host logic of the N>1 path on CPU: two processes, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from laser_slam_b200 import dist as lsd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = lsd.Exchange(rank, world, device=None)
    mine = lsd.shard(8, rank, world)
    results = []
    for step in range(3):
        T = np.eye(4)
        T[:3, 3] = [rank + 0.5 * step, -rank, 0.25]
        c, s = np.cos(0.01 * (rank + 1)), np.sin(0.01 * (rank + 1))
        T[:2, :2] = [[c, -s], [s, c]]
        recs = ex.allgather(lsd.pose_record(T, status=rank, key=step))
        results.append(recs)
    # split-phase form: post step s, collect it when step s+1 is about to be posted
    late = []
    assert ex.collect() is None
    for step in range(3):
        got = ex.collect()
        if got is not None:
            late.append(got)
        T = np.eye(4)
        T[:3, 3] = [10 * rank + step, 0, 0]
        ex.post(lsd.pose_record(T, status=0, key=100 + step))
    late.append(ex.collect())
    tmax = lsd.max_over_ranks([1.0 + rank, 5.0 - rank])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), recs=np.stack(results), late=np.stack(late), mine=np.array(mine),
             tmax=np.array(tmax))
    dist.destroy_process_group()


def test_two_rank_gloo_exchange(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    # every rank sees every rank's record, in rank order, identical on both sides
    assert r0["recs"].tobytes() == r1["recs"].tobytes()
    recs = r0["recs"]
    assert recs.shape == (3, 2)
    for step in range(3):
        for rank in range(2):
            assert recs[step, rank]["status"] == rank and recs[step, rank]["key"] == step
            assert np.allclose(recs[step, rank]["delta"][:3], [rank + 0.5 * step, -rank, 0.25])
            assert np.allclose(recs[step, rank]["delta"][3:], [0, 0, 0.01 * (rank + 1)], atol=1e-6)
    late = r0["late"]
    assert late.tobytes() == r1["late"].tobytes() and late.shape == (3, 2)
    for step in range(3):
        for rank in range(2):
            assert late[step, rank]["key"] == 100 + step and late[step, rank]["delta"][0] == 10 * rank + step
    # shards are disjoint and cover the sequences; time reduction is the max over ranks
    assert sorted(list(r0["mine"]) + list(r1["mine"])) == list(range(8)) and not set(r0["mine"]) & set(r1["mine"])
    assert list(r0["tmax"]) == [2.0, 5.0] and list(r1["tmax"]) == [2.0, 5.0]


def test_pose_record_layout():
    from laser_slam_b200 import dist as lsd
    assert lsd.RECORD_DTYPE.itemsize == 32
    rec = lsd.pose_record(np.eye(4), status=1, key=7)
    assert rec["status"] == 1 and rec["key"] == 7 and not rec["delta"].any()
    ex = lsd.Exchange(0, 1)
    assert ex.allgather(rec)[0] == rec
    ex.post(rec)
    assert ex.collect()[0] == rec and ex.collect() is None
    assert lsd.shard(10, 1, 4) == [1, 5, 9]
