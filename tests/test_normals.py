"""Surface-normal estimation: oracle sanity on the CPU, device vs oracle on the GPU."""
import numpy as np
import pytest


def test_oracle_knn_matches_numpy(oracle_mod, small_pair):
    pts = small_pair["ref"][:3000]
    ids, d2 = oracle_mod.knn_self(pts, 6)
    mu = oracle_mod.mean(pts)
    c = (pts[:, :3] - mu).astype(np.float32)
    D = ((c[:, None, :] - c[None, :, :]) ** 2)
    D = (D[..., 0] + D[..., 1]) + D[..., 2]          # same float32 summation order as the oracle
    order = np.lexsort((np.arange(len(c))[None, :].repeat(len(c), 0), D), axis=1)[:, :6]
    assert np.array_equal(ids, order.astype(np.int32))
    assert np.array_equal(d2, np.take_along_axis(D, order, 1))
    assert (ids[:, 0] == np.arange(len(c))).all() and (d2[:, 0] == 0).all()   # self is the nearest


def test_oracle_normals_on_analytic_surfaces(oracle_mod):
    rng = np.random.default_rng(0)
    # a plane z = 0.3 x + 2, sensor at the origin below/above it
    xy = rng.uniform(-5, 5, (4000, 2))
    pts = np.ones((4000, 4), np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = xy[:, 0], xy[:, 1], 0.3 * xy[:, 0] + 2.0
    n = oracle_mod.knn_normals(pts, 10)
    ref = np.array([0.3, 0.0, -1.0]) / np.linalg.norm([0.3, 0.0, -1.0])   # pointing towards the origin (z decreasing)
    assert np.abs(n - ref).max() < 1e-4
    # a sphere of radius 5 around the sensor: normals point inwards
    v = rng.normal(size=(6000, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    sph = np.ones((6000, 4), np.float32); sph[:, :3] = 5 * v
    n = oracle_mod.knn_normals(sph, 12)
    assert ((n * v).sum(1) < -0.99).all()
    # fewer than 3 points -> zero normals
    assert not oracle_mod.knn_normals(pts[:2], 10).any()


def test_oracle_normals_agree_with_scene_normals(oracle_mod, scans):
    """On a full synthetic scan (2 cm range noise) the estimated normals match the analytic ones away from edges
    (a 16x azimuth-subsampled cloud would not: its 10-neighbourhoods degenerate into ring segments)."""
    n = oracle_mod.knn_normals(scans[0][0], 10, num_threads=8)
    agree = (n * scans[0][1]).sum(1)
    assert np.median(np.abs(agree)) > 0.97 and (np.abs(agree) > 0.9).mean() > 0.95
    # both are oriented towards the sensor; the sign is only well defined away from grazing incidence
    p = scans[0][0][:, :3]
    cosinc = -(scans[0][1] * p).sum(1) / np.linalg.norm(p, axis=1)
    front = cosinc > 0.5   # incidence within 60 deg of the normal: an estimate within ~25 deg keeps its sign
    assert front.mean() > 0.2 and (agree[front] > 0.9).mean() > 0.95


@pytest.mark.gpu
def test_gpu_normals_bit_exact(gpu_ctx, oracle_mod, scans, small_pair):
    for pts, k in [(small_pair["ref"], 10), (small_pair["reading"], 5), (scans[2][0], 10), (small_pair["ref"][:7], 10),
                   (small_pair["ref"][:2], 4)]:
        o = oracle_mod.knn_normals(pts, k, num_threads=8)
        g = gpu_ctx.estimate_normals(pts, k)
        assert np.array_equal(o, g), (len(pts), k, int((o != g).any(1).sum()))


@pytest.mark.gpu
def test_gpu_icp_with_estimated_normals(gpu_ctx, oracle_mod, scans, traj):
    """Scans pushed WITHOUT normals (estimated on the device) register as well as with the analytic normals."""
    import laser_slam_b200 as ls
    truth, odom = traj
    mp = gpu_ctx.create_map(8, 131072)
    sid = [mp.push_scan_estimate_normals(scans[k][0], 10) for k in range(5)]
    Tparts = [np.eye(4, dtype=np.float32) if k == 3 else (np.linalg.inv(truth[3]) @ truth[k]).astype(np.float32) for k in [3, 2, 1, 0]]
    T0 = (np.linalg.inv(truth[3]) @ odom[4]).astype(np.float32)
    g = mp.register(sid[4], [sid[3], sid[2], sid[1], sid[0]], Tparts, T0, ls.default_params(max_iterations=30, use_differential=0))
    rel = np.linalg.inv(truth[3]) @ truth[4]
    assert np.abs(g["T"][:3, 3] - rel[:3, 3]).max() < 0.01
    # the assembled sub-map carries the estimated normals, rotated with the scans
    pts, nrm = mp.assemble([sid[3]], [np.eye(4, dtype=np.float32)])
    assert np.array_equal(nrm, oracle_mod.knn_normals(scans[3][0], 10, num_threads=8))
    mp.close()


@pytest.mark.gpu
def test_default_chain_with_its_datapoints_filters_equals_oracle(gpu_ctx, oracle_mod, small_pair):
    """The whole chain of icp_default.yaml:1-27 -- reading RandomSampling(0.5), reference SurfaceNormal(knn 10), KDTreeMatcher,
    TrimmedDist(0.75), PointToPlane, Counter(40) + Differential -- in its deterministic form (ls_keep_point instead of
    rand(), exact k-NN normals): GPU path == oracle restatement, bit for bit; and it differs from the unfiltered chain."""
    import laser_slam_b200 as ls
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_abi import REF_YAML
    p = ls.params_from_yaml(REF_YAML)
    assert p.reading_sampling_prob == 0.5 and p.reference_normals_knn == 10
    rd, ref, nrm = ls.apply_chain_filters(gpu_ctx, small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], p)
    o = oracle_mod
    rd_o = small_pair["reading"][o.keep_mask(len(small_pair["reading"]), ls.READING_SALT, 0.5)]
    nrm_o = o.knn_normals(small_pair["ref"], 10)
    assert np.array_equal(rd, rd_o) and np.array_equal(nrm, nrm_o) and 0.45 < len(rd) / len(small_pair["reading"]) < 0.55
    po = o.default_params(max_iterations=p.max_iterations, trim_ratio=p.trim_ratio, use_differential=p.use_differential,
                          min_diff_rot=p.min_diff_rot, min_diff_trans=p.min_diff_trans, smooth_length=p.smooth_length)
    r = o.icp(rd_o, small_pair["ref"], nrm_o, small_pair["T0"], po, want_hist=True)
    g = gpu_ctx.icp_register(rd, ref, nrm, small_pair["T0"], p, want_ids=True, want_hist=True)
    assert r["rc"] == 0 and g["rc"] == 0 and g["stats"].iterations == r["stats"].iterations
    assert np.array_equal(g["T_iter_hist"], r["T_iter_hist"]) and np.array_equal(g["ids"], r["ids_hist"][-1])
    assert np.array_equal(g["T"], r["T"])
    plain = gpu_ctx.icp_register(small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], small_pair["T0"], p)
    assert not np.array_equal(plain["T"], g["T"])
    assert np.abs(plain["T"][:3, 3] - g["T"][:3, 3]).max() < 0.5           # both land near the same registration (512-pt rings)
