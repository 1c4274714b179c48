"""bench.py --config 4: BASELINE.json configs[3], the IncrementalEstimator pose-graph solve.

5000 poses / 200 loop-closure BetweenFactors (1 prior + 4999 odometry + 4999 ICP (Cauchy) + 200
loop closures, 10 % gross outliers, dead-reckoned initial values).  A "step" = one IncrementalEstimator::estimate call =
3 Gauss-Newton iterations over the whole graph (reference laser_slam/src/incremental_estimator.cpp:151-163: three
isam2_.update() calls) through ls_pg_optimize(3) -- upload of the factor table and values, device solve, download of all
poses, i.e. exactly what the C++ host layer does per scan.  Also reported: the batch solve to convergence, the marginal
covariances (gtsam::Marginals), and the oracle (scipy sparse Cholesky Gauss-Newton) on the host next to them.
The solve does not shard at 5000 poses: with --gpus N every rank runs a replica."""
import json
import os
import time

import numpy as np


def main(args):
    import torch
    import laser_slam_b200 as ls
    from oracle import posegraph_oracle as pg
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)
    keys, init, factors, truth = pg.make_config4()
    g = ls.PoseGraph(local)
    g.add_poses(keys, init)
    g.add_factors(factors)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        g.set_poses(keys, init)
        g.optimize(3)
    sampler = bench.ClockSampler(local)
    sampler.start()
    dev = []
    l0 = g.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.set_poses(keys, init)
        st = g.optimize(3)
        dev.append(st.device_ms)
    barrier()
    t_e2e = time.perf_counter() - t0
    launches = g.launch_count - l0
    clocks = sampler.summary()
    # batch solve to convergence
    g.set_poses(keys, init)
    t0 = time.perf_counter()
    its, dmax = 0, 1.0
    while its < 40 and dmax > 1e-9:
        st = g.optimize(1)
        dmax, its = st.last_step_max, its + 1
    t_conv = time.perf_counter() - t0
    k2, est = g.poses()
    t0 = time.perf_counter()
    cov64 = g.marginals(keys[::len(keys) // 64][:64])
    t_m64 = time.perf_counter() - t0
    t0 = time.perf_counter()
    cov_all = g.marginals(keys)
    t_mall = time.perf_counter() - t0
    from laser_slam_b200 import dist as lsd
    t_e2e, t_dev = lsd.max_over_ranks([t_e2e, float(np.sum(dev)) * 1e-3], device=local)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # oracle on the host: 3 iterations, and to convergence (same stopping rule)
    t0 = time.perf_counter()
    ref3, _ = pg.optimize(factors, keys, init, iters=3)
    t_cpu3 = time.perf_counter() - t0
    t0 = time.perf_counter()
    refc, hist = pg.optimize(factors, keys, init, iters=40, tol=1e-9)
    t_cpuc = time.perf_counter() - t0
    err_t = float(np.abs(est[:, 4:] - refc[:, 4:]).max())
    err_r = float(np.abs(pg.so3_log(np.swapaxes(pg.quat_to_R(est[:, :4]), -1, -2) @ pg.quat_to_R(refc[:, :4]))).max())
    P, F = len(keys), len(factors)
    E = st.n_border
    alg_bytes = 3 * (P * 56 + F * 120 + (P + F - 1) * 288 + P * 48)   # poses + factors + block Hessian + gradient
    peak, peak_src = bench.load_peaks()
    out = {
        "metric": "pose-graph estimate() calls/s (5000 poses, 200 loop closures, 3 Gauss-Newton iterations per call)",
        "value": world * args.steps / t_dev, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[3]: IncrementalEstimator pose-graph, 5000 poses / 200 loop-closure BetweenFactors, "
                               "on-device Gauss-Newton (3 iterations per estimate() call)",
                   "poses": P, "factors": F, "border_factors": int(E), "replicas": world,
                   "sharding": "replicas only (the solve does not shard at this size, SURVEY.md §8e)",
                   "l2": "working set ~300 MB (dense Woodbury panel Z) > L2",
                   "to_convergence": {"gn_iterations": its, "wall_ms": 1e3 * t_conv, "last_step_max": dmax,
                                      "cpu_oracle_wall_ms": 1e3 * t_cpuc, "cpu_oracle_iterations": len(hist),
                                      "max_abs_diff_vs_oracle_m": err_t, "max_abs_diff_vs_oracle_rad": err_r},
                   "marginals": {"64_poses_wall_ms": 1e3 * t_m64, "all_poses_wall_ms": 1e3 * t_mall,
                                 "trace_of_last_pose_covariance": float(np.trace(cov_all[-1]))}},
        "e2e": {"value": world * args.steps / t_e2e, "unit": "solves/s",
                "h2d_bytes_per_step": int(P * 56 + F * 232 + P * 12), "d2h_bytes_per_step": int(P * 56),
                "note": "ls_pg_set_poses + ls_pg_optimize(3): host graph -> device tables, solve, all poses back"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "cr_fwd_keep_kernel / cr_bwd_kernel (block cyclic reduction of the chain Hessian over all right-hand sides)",
                     "achieved": alg_bytes / (t_dev / args.steps) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / (t_dev / args.steps) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "note": "launch- and level-latency bound (about 300 short launches per estimate), not HBM-bound (SURVEY.md §8d): the fraction is for completeness"},
        "cpu_baseline": {"value": 1.0 / t_cpu3, "unit": "solves/s", "cores": 1, "kind": "port",
                         "sample": "oracle/posegraph_oracle.py optimize(iters=3) on the same graph: scipy.sparse assembly + "
                                   "spsolve (SuperLU), one run"},
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
