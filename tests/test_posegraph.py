"""Pose-graph path: oracle self-checks on the CPU, device solver vs oracle on the GPU.

Tolerance (BASELINE.json north_star): poses within 1e-4 m / 1e-5 rad of the CPU path.  Both sides run the same
float64 Gauss-Newton, so after the same number of iterations they agree to ~1e-9; the tests assert 1e-7."""
import numpy as np
import pytest

from oracle import posegraph_oracle as pg


def rand_pose(rng, scale=3.0, ang=0.4):
    T = np.eye(4)
    T[:3, :3] = pg.so3_exp(rng.normal(scale=ang, size=3))
    T[:3, 3] = rng.normal(size=3) * scale
    return pg.se3_from_matrix(T)


def test_oracle_so3_roundtrips():
    rng = np.random.default_rng(0)
    w = rng.normal(scale=0.8, size=(50, 3))
    assert np.allclose(pg.so3_log(pg.so3_exp(w)), w, atol=1e-12)
    q = pg.R_to_quat(pg.so3_exp(w))
    assert np.allclose(pg.quat_to_R(q), pg.so3_exp(w), atol=1e-12)
    P = np.stack([rand_pose(rng) for _ in range(20)])
    I = pg.se3_to_matrix(pg.se3_compose(P, pg.se3_inverse(P)))
    assert np.allclose(I, np.eye(4), atol=1e-12)


@pytest.mark.parametrize("ftype", [pg.PRIOR, pg.BETWEEN])
def test_oracle_jacobians_numeric(ftype):
    rng = np.random.default_rng(1)
    keys = np.array([5, 9], np.uint64)
    poses = np.stack([rand_pose(rng), rand_pose(rng)])
    meas = pg.se3_compose(pg.se3_inverse(poses[0]), poses[1]) if ftype == pg.BETWEEN else poses[0]
    meas = pg.se3_compose(meas, rand_pose(rng, 0.05, 0.03))
    fac = [pg.make_factor(ftype, 5, 9, meas, [0.05] * 3 + [0.015] * 3)]
    r0, Ja, Jb, ia, ib, _ = pg.linearize(fac, keys, poses)
    eps = 1e-6
    num = np.zeros((2, 6, 6))
    for node in range(2):
        for k in range(6):
            d = np.zeros((2, 6))
            d[node, k] = eps
            r1 = pg.linearize(fac, keys, pg.retract(poses, d))[0][0]
            r2 = pg.linearize(fac, keys, pg.retract(poses, -d))[0][0]
            num[node, :, k] = (r1 - r2) / (2 * eps)
    if ftype == pg.PRIOR:
        assert np.abs(num[0] - Jb[0]).max() < 1e-6 and np.abs(num[1]).max() == 0
    else:
        assert np.abs(num[0] - Ja[0]).max() < 1e-6 and np.abs(num[1] - Jb[0]).max() < 1e-6


def test_oracle_recovers_consistent_graph():
    """Noise-free measurements: Gauss-Newton from perturbed initial values returns the generating poses."""
    rng = np.random.default_rng(2)
    P = 40
    truth = [rand_pose(rng, 0.0, 0.0)]
    for k in range(1, P):
        truth.append(pg.se3_compose(truth[-1], rand_pose(rng, 0.5, 0.05)))
    truth = np.stack(truth)
    keys = np.arange(P, dtype=np.uint64) + 7
    sig = [0.005] * 3 + [0.0015] * 3
    fac = [pg.make_factor(pg.PRIOR, keys[0], 0, truth[0], [1e-7] * 6)]
    for k in range(1, P):
        fac.append(pg.make_factor(pg.BETWEEN, keys[k - 1], keys[k], pg.se3_compose(pg.se3_inverse(truth[k - 1]), truth[k]), sig))
    fac.append(pg.make_factor(pg.BETWEEN, keys[3], keys[30], pg.se3_compose(pg.se3_inverse(truth[3]), truth[30]), sig, robust=1))
    init = np.stack([pg.se3_compose(t, rand_pose(rng, 0.05, 0.01)) for t in truth])
    est, hist = pg.optimize(fac, keys, init, iters=8)
    assert np.abs(est[:, 4:] - truth[:, 4:]).max() < 1e-8 and hist[-1][0] < 1e-8


def build_gpu_graph(ls, keys, init, factors, tracks=None):
    g = ls.PoseGraph(0)
    g.add_poses(keys, init, tracks)
    idx = g.add_factors(factors)
    return g, idx


@pytest.mark.gpu
def test_gpu_matches_oracle_chain_with_loop_closures():
    import laser_slam_b200 as ls
    keys, init, factors, truth = pg.make_config4(n_poses=400, n_lc=15, lap=100, seed=3)
    g, _ = build_gpu_graph(ls, keys, init, factors)
    for iters in (1, 3):
        g.set_poses(keys, init)
        st = g.optimize(iters)
        ref, hist = pg.optimize(factors, keys, init, iters=iters)
        k2, est = g.poses()
        assert np.array_equal(k2, keys) and st.n_border == 15 and st.n_poses == 400
        assert np.abs(est[:, 4:] - ref[:, 4:]).max() < 1e-7
        assert np.abs(pg.so3_log(np.swapaxes(pg.quat_to_R(est[:, :4]), -1, -2) @ pg.quat_to_R(ref[:, :4]))).max() < 1e-7
        assert abs(st.cost_first - pg.linearize(factors, keys, init)[5]) < 1e-6 * max(1.0, st.cost_first)
    g.close()


@pytest.mark.gpu
def test_gpu_no_loop_closures_multi_track_and_fixed_node():
    import laser_slam_b200 as ls
    rng = np.random.default_rng(5)
    sig = [0.005] * 3 + [0.0015] * 3
    keys, init, tracks, factors = [], [], [], []
    for t in range(3):
        pose = rand_pose(rng, 20.0, 0.3)
        first = True
        for k in range(60):
            key = 1000 * (t + 1) + k
            keys.append(key); tracks.append(t)
            if first:
                factors.append(pg.make_factor(pg.PRIOR, key, 0, pose, [1e-7] * 6))
                first = False
            else:
                rel = rand_pose(rng, 0.5, 0.04)
                factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, pg.se3_compose(rel, rand_pose(rng, 0.01, 0.002)), sig))
                factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, pg.se3_compose(rel, rand_pose(rng, 0.004, 0.001)), sig, robust=1))
                pose = pg.se3_compose(pose, rel)
            init.append(pg.se3_compose(pose, rand_pose(rng, 0.03, 0.005)))
    keys = np.array(keys, np.uint64); init = np.stack(init)
    # inter-track loop closure + a fixed-first-node ICP factor (LaserTrack::appendICPFactors, laser_track.cpp:371-381)
    factors.append(pg.make_factor(pg.BETWEEN, 1010, 2030, pg.se3_compose(pg.se3_inverse(init[10]), init[60 + 30]), sig, robust=1))
    factors.append(pg.make_factor(pg.BETWEEN, 3005, 3020, pg.se3_compose(pg.se3_inverse(init[125]), init[140]), sig,
                                  fix_a=1, fixed_a7=init[125]))
    perm = rng.permutation(len(keys))  # insertion order interleaves the tracks (time order is kept inside a track)
    perm = np.array(sorted(perm, key=lambda i: (0, 0)))  # keep original order inside tracks; interleave below
    order = np.argsort(np.arange(len(keys)) % 60, kind="stable")
    g, _ = build_gpu_graph(ls, keys[order], init[order], factors, np.array(tracks)[order])
    st = g.optimize(3)
    ref, _ = pg.optimize(factors, keys, init, iters=3)
    k2, est = g.poses()
    back = {int(k): i for i, k in enumerate(k2)}
    est = est[[back[int(k)] for k in keys]]
    assert st.n_border == 1
    assert np.abs(est[:, 4:] - ref[:, 4:]).max() < 1e-7
    g.close()


@pytest.mark.gpu
def test_gpu_incremental_estimate_flow_and_factor_removal():
    """The per-scan flow of IncrementalEstimator::estimate: add one pose + two factors, 3 iterations, repeat;
    then remove a factor (estimateAndRemove's removeFactorIndices) and compare with the oracle."""
    import laser_slam_b200 as ls
    keys, init, factors, truth = pg.make_config4(n_poses=60, n_lc=0, lap=30, seed=9)
    g = ls.PoseGraph(0)
    g.add_poses(keys[:1], init[:1])
    idx_all = list(g.add_factors(factors[:1]))
    g.optimize(3)
    cur = {int(keys[0]): g.poses()[1][0]}
    o_poses = init[:1].copy()
    o_poses, _ = pg.optimize(factors[:1], keys[:1], o_poses, iters=3)
    for k in range(1, 60):
        new_f = factors[1 + 2 * (k - 1): 1 + 2 * k]
        g.add_poses(keys[k:k + 1], init[k:k + 1])
        idx_all += list(g.add_factors(new_f))
        g.optimize(3)
        o_poses = np.concatenate([o_poses, init[k:k + 1]])
        o_poses, _ = pg.optimize(factors[:1 + 2 * k], keys[:k + 1], o_poses, iters=3)
    est = g.poses()[1]
    assert np.abs(est[:, 4:] - o_poses[:, 4:]).max() < 1e-7
    # remove the robust ICP factor of the last edge
    g.remove_factors([idx_all[-1]])
    assert lib_num_factors(g) == len(factors) - 1
    g.optimize(2)
    o2, _ = pg.optimize(factors[:-1], keys, o_poses, iters=2)
    assert np.abs(g.poses()[1][:, 4:] - o2[:, 4:]).max() < 1e-7
    g.close()


def lib_num_factors(g):
    import laser_slam_b200 as ls
    return ls.lib().ls_pg_num_factors(g._h)


@pytest.mark.gpu
def test_gpu_prior_removed_after_linking_tracks():
    """estimateAndRemove: the second track loses its prior once a loop closure links it to track 0
    (reference incremental_estimator.cpp:212-237); the solver damps its gauge instead of refusing."""
    import laser_slam_b200 as ls
    rng = np.random.default_rng(11)
    sig = [0.005] * 3 + [0.0015] * 3
    loose = [0.05] * 3 + [0.015] * 3
    keys, init, tracks, factors, truth = [], [], [], [], []
    for t in range(2):
        pose = rand_pose(rng, 5.0, 0.2)
        for k in range(40):
            key = 1000 * (t + 1) + k
            keys.append(key); tracks.append(t)
            if k == 0:
                factors.append(pg.make_factor(pg.PRIOR, key, 0, pose, [1e-7] * 6))
            else:
                rel = rand_pose(rng, 0.5, 0.04)
                factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, pg.se3_compose(rel, rand_pose(rng, 0.004, 0.001)), sig, robust=1))
                pose = pg.se3_compose(pose, rel)
            truth.append(pose)
            init.append(pg.se3_compose(pose, rand_pose(rng, 0.02, 0.004)))
    keys = np.array(keys, np.uint64); init = np.stack(init); truth = np.stack(truth)
    g = ls.PoseGraph(0)
    g.add_poses(keys, init, np.array(tracks))
    idx = g.add_factors(factors)
    g.optimize(3)
    o, _ = pg.optimize(factors, keys, init, iters=3)
    # link: first-association factor (loose noise) replaces the prior of track 1
    lc = pg.make_factor(pg.BETWEEN, 1010, 2020, pg.se3_compose(pg.se3_inverse(truth[10]), truth[40 + 20]), loose)
    prior1 = [i for i, f in enumerate(factors) if f["type"] == pg.PRIOR and f["key_a"] == 2000][0]
    g.remove_factors([idx[prior1]])
    g.add_factors([lc])
    st = g.optimize(3)
    f2 = [f for i, f in enumerate(factors) if i != prior1] + [lc]
    o2, _ = pg.optimize(f2, keys, o, iters=3, damp_keys=[2000])
    est = g.poses()[1]
    assert st.n_border == 1
    assert np.abs(est[:, 4:] - o2[:, 4:]).max() < 1e-7
    assert np.abs(est[:, 4:] - truth[:, 4:]).max() < 0.1
    g.close()


@pytest.mark.gpu
def test_gpu_errors():
    import laser_slam_b200 as ls
    g = ls.PoseGraph(0)
    keys = np.array([1, 2], np.uint64)
    poses = np.array([[1, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 1, 0, 0]], np.float64)
    g.add_poses(keys, poses)
    with pytest.raises(ls.LsError):
        g.add_poses(keys[:1], poses[:1])                      # duplicate key
    g.add_factors([pg.make_factor(pg.BETWEEN, 1, 2, poses[1], [0.01] * 6)])
    with pytest.raises(ls.LsError):                            # no prior: gauge freedom -> refused loudly
        g.optimize(1)
    with pytest.raises(ls.LsError):
        g.add_factors([pg.make_factor(pg.BETWEEN, 1, 2, poses[1], [0.0] * 6)])
    g.close()


@pytest.mark.gpu
def test_gpu_config4_full_size():
    """BASELINE.json configs[3]: 5000 poses, 200 loop closures; 3 GN iterations vs the oracle."""
    import laser_slam_b200 as ls
    keys, init, factors, truth = pg.make_config4()
    g, _ = build_gpu_graph(ls, keys, init, factors)
    st = g.optimize(3)
    ref, hist = pg.optimize(factors, keys, init, iters=3)
    est = g.poses()[1]
    assert st.n_border == 200 and st.n_factors == len(factors)
    assert np.abs(est[:, 4:] - ref[:, 4:]).max() < 1e-6
    assert np.abs(pg.so3_log(np.swapaxes(pg.quat_to_R(est[:, :4]), -1, -2) @ pg.quat_to_R(ref[:, :4]))).max() < 1e-7
    g.close()


def test_oracle_marginals_match_sampling_free_identities():
    """Oracle self-check: the prior alone fixes the first pose's covariance; along a chain the covariance grows."""
    rng = np.random.default_rng(11)
    keys = np.arange(5, dtype=np.uint64) + 3
    poses = np.stack([rand_pose(rng) for _ in range(5)])
    sig0 = np.array([0.1, 0.2, 0.3, 0.01, 0.02, 0.03])
    fac = [pg.make_factor(pg.PRIOR, keys[0], 0, poses[0], sig0)]
    sig = [0.05] * 3 + [0.01] * 3
    for k in range(1, 5):
        fac.append(pg.make_factor(pg.BETWEEN, keys[k - 1], keys[k], pg.se3_compose(pg.se3_inverse(poses[k - 1]), poses[k]), sig))
    C = pg.marginals(fac, keys, poses, keys)
    R0 = pg.quat_to_R(poses[0, :4])        # the prior's translation residual lives in the measured pose's frame
    want0 = np.zeros((6, 6))
    want0[:3, :3] = R0 @ np.diag(sig0[:3] ** 2) @ R0.T
    want0[3:, 3:] = np.diag(sig0[3:] ** 2)
    assert np.allclose(C[0], want0, rtol=1e-9, atol=1e-15)
    # an independent route to the same blocks: eliminate every other pose (Schur complement), invert what is left
    H = pg.hessian(fac, keys, poses)
    for q in range(5):
        keep = np.arange(6 * q, 6 * q + 6)
        rest = np.setdiff1d(np.arange(30), keep)
        schur = H[np.ix_(keep, keep)] - H[np.ix_(keep, rest)] @ np.linalg.solve(H[np.ix_(rest, rest)], H[np.ix_(rest, keep)])
        assert np.allclose(np.linalg.inv(schur), C[q], rtol=1e-7, atol=1e-14)


@pytest.mark.gpu
def test_gpu_marginal_covariances_match_the_inverse_hessian():
    """ls_pg_marginals (LaserTrack::updateCovariancesFromGTSAMValues -> gtsam::Marginals::marginalCovariance, reference
    laser_slam/src/laser_track.cpp:421-429) against numpy.linalg.inv of the oracle's Hessian at the same estimate:
    chain + loop closures (border correction), several tracks, more keys than one chunk."""
    import laser_slam_b200 as ls
    keys, init, factors, truth = pg.make_config4(n_poses=300, n_lc=12, lap=100, seed=5)
    g, _ = build_gpu_graph(ls, keys, init, factors)
    g.optimize(4)
    k2, est = g.poses()
    want = pg.marginals(factors, keys, est, keys)
    got = g.marginals(keys)                              # 300 keys: five chunks of 64
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())
    assert (np.abs(got - want) / scale).max() < 1e-6
    assert np.allclose(got, np.swapaxes(got, 1, 2), atol=1e-12)          # symmetric
    assert (np.linalg.eigvalsh(got) > 0).all()                            # positive definite
    sub = keys[[250, 3, 77]]
    assert np.allclose(g.marginals(sub), want[[250, 3, 77]], rtol=1e-6, atol=1e-15)
    k3, est3 = g.poses()
    assert np.array_equal(est3, est)                                      # asking for marginals does not move the estimate
    g.close()
    # no loop closures, three tracks
    rng = np.random.default_rng(6)
    sig = [0.005] * 3 + [0.0015] * 3
    keys, init, tracks, factors = [], [], [], []
    for t in range(3):
        pose = rand_pose(rng, 20.0, 0.3)
        for k in range(40):
            key = 1000 * (t + 1) + k
            keys.append(key); tracks.append(t)
            if k == 0:
                factors.append(pg.make_factor(pg.PRIOR, key, 0, pose, [1e-3] * 6))
            else:
                rel = rand_pose(rng, 0.5, 0.04)
                factors.append(pg.make_factor(pg.BETWEEN, key - 1, key, rel, sig))
                pose = pg.se3_compose(pose, rel)
            init.append(pose)
    keys = np.array(keys, np.uint64); init = np.stack(init)
    g, _ = build_gpu_graph(ls, keys, init, factors, np.array(tracks))
    got = g.marginals(keys)
    want = pg.marginals(factors, keys, init, keys)
    assert (np.abs(got - want) / np.abs(want).max(axis=(1, 2), keepdims=True)).max() < 1e-6
    g.close()
