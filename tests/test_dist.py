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


def _graph_worker(rank, world, port, out_dir):
    """Every rank feeds its replica of the shared estimator from the gathered records; the replicas must be identical."""
    import torch.distributed as dist
    from laser_slam_b200 import dist as lsd
    from oracle import posegraph_oracle as pg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = lsd.Exchange(rank, world, device=None)
    graph = lsd.ReplicatedGraph(world)
    rng = np.random.default_rng(100 + rank)          # every rank only knows ITS track's registrations
    for step in range(12):
        T = np.eye(4)
        T[:3, 3] = [0.8 + 0.01 * rng.normal(), 0.01 * rng.normal(), 0.0]
        a = 0.02 * (rank + 1) + 0.001 * rng.normal()
        T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        got = ex.collect()                            # split-phase: step s-1's records arrive while step s is posted
        if got is not None:
            graph.feed(got)
        ex.post(lsd.pose_record(T, status=0, key=step))
    graph.feed(ex.collect())
    keys = np.array(graph.keys, np.uint64)
    fac = [pg.make_factor(f["type"], f["key_a"], f["key_b"], f["meas"], f["sigma"], robust=f.get("robust", 0)) for f in graph.factors]
    est, _ = pg.optimize(fac, keys, np.stack(graph.poses), iters=3)
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), digest=np.array(graph.digest()), est=est, keys=keys, n_fac=len(fac))
    dist.destroy_process_group()


def test_replicated_estimator_graphs_are_identical_across_ranks(tmp_path):
    """BASELINE.json configs[2] / SURVEY.md §8e: the 32-byte records are all a rank needs to keep an exact replica of the
    shared pose graph -- same keys, same initial values, same factors, hence the same estimate -- on every rank."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_graph_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npz"), np.load(tmp_path / "g1.npz")
    assert str(g0["digest"]) == str(g1["digest"])
    assert np.array_equal(g0["keys"], g1["keys"]) and len(g0["keys"]) == 2 * 12 and int(g0["n_fac"]) == 2 * 12
    assert np.array_equal(g0["est"], g1["est"])                      # identical graphs -> bit-identical estimates
    track = g0["keys"] >> np.uint64(48)
    assert sorted(set(track.tolist())) == [0, 1]
    assert np.allclose(g0["est"][track == 1][0, 4:], [0, 100, 0], atol=1e-6)   # track 1 anchored at its prior
