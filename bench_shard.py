#!/usr/bin/env python
"""ONE registration sharded by queries over the GPUs of a node.

    python bench_shard.py --config 5 --steps 20 --warmup 3                         # 1 GPU: the unsharded call
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_shard.py --config 5 --steps 20 --warmup 3                             # N shards

One track; every step registers the next scan of the drive against the rolling sub-map, one registration per launch
(the latency case: a single dense sensor that has to be matched before the next scan arrives).  With N ranks every rank
holds the whole map and 1/N of the reading's queries; inside the persistent kernel every GPU stores its partial select
histograms and normal-equation sums into the peers' exchange buffers over NVLink (ls_icp_register_submap_sharded,
include/ls_b200.h).  The timed quantity is the latency of a
registration with the scans resident (max over ranks, CUDA events of the launch + host wall clock), and its inverse.

Parity, outside the clock: every timed registration's transform is compared BIT FOR BIT with the unsharded
ls_icp_register_submap of the same problem on the same rank; on rank 0 the first one is also checked against the
oracle when the workload is small enough (--config 2).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definitions and synthetic drives are shared with bench.py)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5, choices=(2, 5))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--oracle", action="store_true", help="also check the first registration against the CPU oracle")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison with the unsharded call (diagnostic runs)")
    args = ap.parse_args()
    wl = bench.select_workload(args.config)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    import laser_slam_b200 as ls
    from laser_slam_b200 import dist as lsd
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ls.Context(local)
    truth, odom, scans = bench.make_pools([0])[0]     # every rank generates the same drive
    POOL, K_MAP, N_SCAN, ITERS = bench.POOL, bench.K_MAP, bench.N_SCAN, bench.ITERS
    prm = ls.default_params(max_iterations=ITERS, use_differential=0)
    mp = ctx.create_map(POOL + 2, N_SCAN)
    sid = [mp.push_scan(s[0], s[1]) for s in scans]
    reg = lsd.ShardedRegistrar(ctx, rank, world) if world > 1 else None

    n_total = args.warmup + args.steps
    h = [bench.walk(s) for s in range(K_MAP + 1)]
    staged = []
    for s in range(n_total):
        idx = bench.walk(s + K_MAP + 1)
        h.append(idx)
        ref, ks, Ts = bench.submap_parts(truth, h)
        T0 = (np.linalg.inv(truth[ref]) @ odom[idx]).astype(np.float32) if abs(idx - ref) == 1 else np.eye(4, dtype=np.float32)
        staged.append((idx, ref, ks, Ts, T0))

    def one(s):
        idx, ref, ks, Ts, T0 = staged[s]
        if reg is None:
            return mp.register(sid[idx], [sid[k] for k in ks], Ts, T0, prm)
        return reg.register(mp, sid[idx], [sid[k] for k in ks], Ts, T0, prm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        one(s)
    barrier()
    t0 = time.perf_counter()
    outs = [one(s) for s in range(args.warmup, n_total)]
    barrier()
    wall = time.perf_counter() - t0
    icp_ms = float(np.mean([o["stats"].icp_ms for o in outs]))
    dev_ms = float(np.mean([o["stats"].device_ms for o in outs]))
    vals = lsd.max_over_ranks([wall, icp_ms, dev_ms], device=local if world > 1 else None)
    wall, icp_ms, dev_ms = vals

    # parity outside the clock: the unsharded call on this rank's own copy of the map
    same = True
    for s, o in zip(range(args.warmup, n_total), [] if args.no_parity else outs):
        idx, ref, ks, Ts, T0 = staged[s]
        u = mp.register(sid[idx], [sid[k] for k in ks], Ts, T0, prm)
        same = same and bool(np.array_equal(u["T"], o["T"])) and u["stats"].iterations == o["stats"].iterations \
            and u["stats"].last_kept == o["stats"].last_kept
    flags = lsd.max_over_ranks([0.0 if same else 1.0], device=local if world > 1 else None)
    same_all = flags[0] == 0.0
    oracle_equal = None
    if args.oracle and rank == 0:
        import oracle
        idx, ref, ks, Ts, T0 = staged[args.warmup]
        parts = [scans[k] if k == ref else oracle.transform_cloud(T, *scans[k]) for k, T in zip(ks, Ts)]
        r = oracle.icp(scans[idx][0], np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), T0,
                       oracle.default_params(max_iterations=ITERS, use_differential=0, num_threads=bench.usable_threads()))
        oracle_equal = bool(np.array_equal(outs[0]["T"], r["T"]))
    if rank == 0:
        print(json.dumps({
            "metric": wl["metric"] + ", ONE registration at a time, queries sharded over the GPUs",
            "value": args.steps / wall, "unit": "registrations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_registration_wall": 1e3 * wall / args.steps, "icp_kernel_ms": icp_ms, "device_ms": dev_ms,
            "higher_is_better": True, "scaling": "strong", "dtype": "f32 (exact integer reductions)", "data": "synthetic",
            "config": {"workload": wl["name"], "sharding": "queries of one registration" if world > 1 else "none (unsharded call)",
                       "shards": world, "iterations": ITERS},
            "parity": {"bit_equal_to_unsharded_on_every_rank": bool(same_all), "registrations_compared": len(outs),
                       "first_bit_equal_to_oracle": oracle_equal},
            "exchange": {"how": "each GPU stores its partial histograms / sums into a slot of every peer's buffer (CUDA IPC, NVLink) from inside the persistent kernel; readers sum local slots",
                         "bytes_per_iteration_per_peer_pair": "16 KiB histograms + 256 B sums (steady state), 2 arrival signals"},
        }), flush=True)
    if reg is not None:
        reg.close()
    if world > 1:
        dist.destroy_process_group()
    if not same_all:
        raise SystemExit("sharded registration differs from the unsharded one")


if __name__ == "__main__":
    main()
