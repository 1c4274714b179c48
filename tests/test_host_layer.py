"""End-to-end parity of the C++ host layer (laser_slam::LaserTrack + IncrementalEstimator over the C ABI) against a
Python restatement of the same per-scan flow built from the two oracles (ICP + pose graph).

The flow restated (reference laser_slam_ros/src/laser_slam_worker.cpp:124-173, laser_slam/src/laser_track.cpp:122-231,
466-519, laser_slam/src/incremental_estimator.cpp:151-163,268-291): odometry extends the trajectory; the new scan is
registered against the sub-map of the previous `nscan_in_sub_map` scans expressed in the frame of the previous scan,
starting from the trajectory's relative pose; a prior / odometry + ICP factors and the odometry pose as initial value go
to the estimator, which runs three Gauss-Newton passes; the trajectory is overwritten with the estimate."""
import numpy as np
import pytest

from oracle import posegraph_oracle as pg

SIG = [0.005] * 3 + [0.0015] * 3


def oracle_flow(oracle_mod, scans, odom7, K, icp_params):
    keys, factors, graph_poses, traj, icp_rel = [], [], np.zeros((0, 7)), [], []
    for k in range(len(scans)):
        key = 1000 + k
        if k == 0:
            traj.append(odom7[0].copy())
            factors.append(pg.make_factor(pg.PRIOR, key, key, odom7[0], [1e-7] * 6))
            icp_rel.append(np.array([1, 0, 0, 0, 0, 0, 0.0]))
        else:
            rel = pg.se3_compose(pg.se3_inverse(odom7[k - 1]), odom7[k])
            traj.append(pg.se3_compose(traj[k - 1], rel))
            part_idx = [k - 1] + [k - 2 - i for i in range(min(k - 1, K - 1))]
            ref, nrm = [], []
            for idx in part_idx:
                if idx == k - 1:
                    p, n = scans[idx]
                else:
                    T = pg.se3_to_matrix(pg.se3_compose(pg.se3_inverse(traj[k - 1]), traj[idx])).astype(np.float32)
                    if not oracle_mod.check_rigid(T):
                        T = oracle_mod.correct_rigid(T)
                    p, n = oracle_mod.transform_cloud(T, scans[idx][0], scans[idx][1])
                ref.append(p); nrm.append(n)
            T0 = pg.se3_to_matrix(pg.se3_compose(pg.se3_inverse(traj[k - 1]), traj[k])).astype(np.float32)
            r = oracle_mod.icp(scans[k][0], np.concatenate(ref), np.concatenate(nrm), T0, icp_params)
            T_icp = pg.se3_from_matrix(r["T"].astype(np.float64)) if r["rc"] == 0 else pg.se3_from_matrix(T0.astype(np.float64))
            icp_rel.append(T_icp)
            factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, rel, SIG, robust=0))
            factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, T_icp, SIG, robust=1))
        keys.append(key)
        graph_poses = np.concatenate([graph_poses, odom7[k][None]])  # newValues->insert(scan.key, pose.T_w)
        graph_poses, _ = pg.optimize(factors, np.array(keys, np.uint64), graph_poses, iters=3)
        traj = [p.copy() for p in graph_poses]                       # updateFromGTSAMValues
    return np.stack(traj), np.stack(icp_rel)


@pytest.mark.gpu
def test_laser_track_flow_matches_oracle_flow(oracle_mod, synth_mod):
    from laser_slam_b200 import host
    n_scans, K = 7, 4
    truth, odom = synth_mod.trajectory(2, n_scans)
    scans = [synth_mod.subsample(*synth_mod.scan(truth[k], 2, k), 16) for k in range(n_scans)]
    odom7 = pg.se3_from_matrix(odom)
    # LaserTrack falls back to libpointmatcher's setDefault() chain when no YAML is given
    po = oracle_mod.default_params(trim_ratio=0.85, min_diff_rot=0.001, min_diff_trans=0.001, smooth_length=3)
    ref_traj, ref_icp = oracle_flow(oracle_mod, scans, odom7, K, po)

    est = host.Estimator(n_workers=1, nscan_in_sub_map=K)
    got_icp = []
    for k in range(n_scans):
        icp7, st = est.step(0, k * 100_000_000, odom7[k], scans[k][0], scans[k][1])
        got_icp.append(icp7)
        if k > 0:
            assert st.iterations >= 1
    times, traj = est.trajectory(0)
    assert est.num_scans(0) == n_scans and list(times) == [k * 100_000_000 for k in range(n_scans)]
    got_icp = np.stack(got_icp)
    assert np.abs(got_icp[:, 4:] - ref_icp[:, 4:]).max() < 1e-6            # ICP relative poses
    assert np.abs(traj[:, 4:] - ref_traj[:, 4:]).max() < 1e-6              # estimated trajectory
    dR = np.swapaxes(pg.quat_to_R(traj[:, :4]), -1, -2) @ pg.quat_to_R(ref_traj[:, :4])
    assert np.abs(pg.so3_log(dR)).max() < 1e-6
    # and it is a sensible trajectory: closer to truth than raw odometry at the end
    truth7 = pg.se3_from_matrix(np.linalg.inv(truth[0]) @ truth)
    est_rel = pg.se3_compose(pg.se3_inverse(np.repeat(traj[:1], n_scans, 0)), traj)
    odo_rel = pg.se3_compose(pg.se3_inverse(np.repeat(odom7[:1], n_scans, 0)), odom7)
    assert np.linalg.norm(est_rel[-1, 4:] - truth7[-1, 4:]) < np.linalg.norm(odo_rel[-1, 4:] - truth7[-1, 4:])
    # buildSubMapAroundTime == oracle transform + concatenate around the centre scan
    sub, sub_n = est.build_submap(0, 3 * 100_000_000, 1, 3 * len(scans[0][0]))
    parts = [scans[3]]
    for idx in (2, 4):
        T = pg.se3_to_matrix(pg.se3_compose(pg.se3_inverse(traj[3]), traj[idx])).astype(np.float32)
        parts.append(oracle_mod.transform_cloud(T, *scans[idx]))
    assert sub.shape[0] == 3 * len(scans[0][0])
    assert np.abs(sub - np.concatenate([p[0] for p in parts])).max() < 1e-4
    est.close()


@pytest.mark.gpu
def test_two_workers_and_loop_closure(oracle_mod, synth_mod):
    """Multi-robot flow: two tracks with their own priors, then a loop closure links them: the prior of track 1 is
    removed and the first-association factor used (reference incremental_estimator.cpp:165-266)."""
    from laser_slam_b200 import host
    n_scans = 5
    truth, odom = synth_mod.trajectory(3, 2 * n_scans + 2)
    scans = [synth_mod.subsample(*synth_mod.scan(truth[k], 3, k), 32) for k in range(2 * n_scans + 2)]
    odom7 = pg.se3_from_matrix(odom)
    est = host.Estimator(n_workers=2, nscan_in_sub_map=3, do_icp_step_on_loop_closures=True, loop_closures_sub_maps_radius=1)
    # worker 0 drives scans 0..4, worker 1 drives scans 5..9 of the same street, with an offset world frame
    off = pg.se3_from_matrix(np.array([[1, 0, 0, 50.0], [0, 1, 0, 20.0], [0, 0, 1, 0], [0, 0, 0, 1.0]]))
    for k in range(n_scans):
        est.step(0, k * 10**8, odom7[k], *scans[k])
        est.step(1, k * 10**8, pg.se3_compose(off, odom7[n_scans + k]), *scans[n_scans + k])
    t0, traj0 = est.trajectory(0)
    t1, traj1_before = est.trajectory(1)
    # loop closure between node 4 of track 0 and node 0 of track 1 (consecutive scans of the street): world-frame
    # correction w_T_a_b such that T_w_a^-1 * w_T_a_b * T_w_b == true relative pose
    rel_true = pg.se3_from_matrix(np.linalg.inv(truth[4]) @ truth[5])
    w_T = pg.se3_compose(pg.se3_compose(traj0[4], rel_true), pg.se3_inverse(traj1_before[0]))
    est.loop_closure(0, 4 * 10**8, 1, 0, w_T)
    _, traj0_after = est.trajectory(0)
    _, traj1_after = est.trajectory(1)
    assert np.abs(traj0_after[:, 4:] - traj0[:, 4:]).max() < 0.05          # track 0 keeps its prior
    rel_after = pg.se3_compose(pg.se3_inverse(traj0_after[4]), traj1_after[0])
    assert np.abs(rel_after[4:] - rel_true[4:]).max() < 0.05               # track 1 was pulled into track 0's frame
    assert np.abs(traj1_after[0, 4:] - traj1_before[0, 4:]).max() > 10.0
    with pytest.raises(Exception):
        est.loop_closure(0, 123, 1, 0, w_T)                                # no node at that time: CHECK
    est.close()


def test_host_layer_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from laser_slam_b200 import host
    import laser_slam_b200 as ls
    with pytest.raises(ls.LsError, match="no usable CUDA device"):
        host.Estimator()


@pytest.mark.gpu
def test_batched_scan_callbacks_equal_one_callback_per_worker(synth_mod):
    """IncrementalEstimator::processPosesAndLaserScans (all workers' registrations of a step in ONE batched launch on the
    estimator's shared context and scan ring) gives every worker the same bits as calling
    LaserTrack::processPoseAndLaserScan worker by worker (reference laser_slam_ros/src/laser_slam_worker.cpp:133,158)."""
    from laser_slam_b200 import host
    W, n_scans = 3, 6
    data = []
    for w in range(W):
        truth, odom = synth_mod.trajectory(w, n_scans)
        sc = [synth_mod.subsample(*synth_mod.scan(truth[k], w, k), 8) for k in range(n_scans)]
        data.append((pg.se3_from_matrix(odom), sc))
    one = host.Estimator(n_workers=W, nscan_in_sub_map=3)
    bat = host.Estimator(n_workers=W, nscan_in_sub_map=3)
    for k in range(n_scans):
        ws = [w for w in range(W) if not (w == 2 and k == 0)]         # worker 2 starts one step late: mixed first / later scans
        feats = [np.ascontiguousarray(data[w][1][k][0]) for w in ws]
        nrms = [np.ascontiguousarray(data[w][1][k][1]) for w in ws]
        icp_b, st_b = bat.step_batch(ws, [k * 100 for _ in ws], [data[w][0][k] for w in ws], [f.ctypes.data for f in feats],
                                     [n.ctypes.data for n in nrms], [len(f) for f in feats])
        for j, w in enumerate(ws):
            icp_1, st_1 = one.step(w, k * 100, data[w][0][k], feats[j], nrms[j])
            assert np.array_equal(icp_1, icp_b[j]), (k, w)
            assert st_1.iterations == st_b[j].iterations and st_1.last_kept == st_b[j].last_kept
    for w in range(W):
        t1, p1 = one.trajectory(w)
        t2, p2 = bat.trajectory(w)
        assert np.array_equal(t1, t2) and np.abs(p1 - p2).max() < 1e-9       # same factors -> same estimate
    one.close()
    bat.close()


@pytest.mark.gpu
def test_split_step_with_prefetch_and_views_equals_step_batch(synth_mod):
    """beginPosesAndLaserScans / prefetchLaserScans(next scans) / endPosesAndLaserScans over DataPoints that borrow pinned
    caller memory (uploaded in place, asynchronously) gives the same bits as processPosesAndLaserScans over copies."""
    import torch
    from laser_slam_b200 import host
    W, n_scans = 3, 7
    data = []
    for w in range(W):
        truth, odom = synth_mod.trajectory(10 + w, n_scans)
        sc = [synth_mod.subsample(*synth_mod.scan(truth[k], 10 + w, k), 8) for k in range(n_scans)]
        data.append((pg.se3_from_matrix(odom), sc))
    # pinned staging the views borrow; kept alive for the life of the estimator
    pin_f = [[torch.from_numpy(np.ascontiguousarray(data[w][1][k][0])).pin_memory() for k in range(n_scans)] for w in range(W)]
    pin_n = [[torch.from_numpy(np.ascontiguousarray(data[w][1][k][1])).pin_memory() for k in range(n_scans)] for w in range(W)]
    ref = host.Estimator(n_workers=W, nscan_in_sub_map=3)
    spl = host.Estimator(n_workers=W, nscan_in_sub_map=3)
    ws = list(range(W))

    def ptrs(k, pinned):
        if pinned:
            return [pin_f[w][k].data_ptr() for w in ws], [pin_n[w][k].data_ptr() for w in ws]
        return [data[w][1][k][0].ctypes.data for w in ws], [data[w][1][k][1].ctypes.data for w in ws]

    for k in range(n_scans):
        ns = [len(data[w][1][k][0]) for w in ws]
        poses = [data[w][0][k] for w in ws]
        fp, npp = ptrs(k, False)
        icp_r, st_r = ref.step_batch(ws, [k * 100] * W, poses, fp, npp, ns)
        fp, npp = ptrs(k, True)
        spl.begin_batch(ws, [k * 100] * W, poses, fp, npp, ns, views=True)
        if k + 1 < n_scans and k != 3:          # one step goes without the hint: the upload then happens in its own begin
            fq, nq = ptrs(k + 1, True)
            spl.prefetch(ws, [(k + 1) * 100] * W, fq, nq, [len(data[w][1][k + 1][0]) for w in ws], views=True)
        icp_s, st_s = spl.end_batch()
        assert np.array_equal(icp_r, icp_s), k
        assert [s.iterations for s in st_r] == [s.iterations for s in st_s]
        assert [s.last_kept for s in st_r] == [s.last_kept for s in st_s]
    for w in ws:
        t1, p1 = ref.trajectory(w)
        t2, p2 = spl.trajectory(w)
        assert np.array_equal(t1, t2) and np.abs(p1 - p2).max() < 1e-9
    ref.close()
    spl.close()
