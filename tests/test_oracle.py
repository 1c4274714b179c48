"""CPU tests of the oracle itself (it is the normative definition of the path: parity unpinned by the
reference, so it is pinned here against independent implementations and committed golden vectors)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_kdtree_equals_brute_force(oracle_mod, small_pair):
    o = oracle_mod
    mu = o.mean(small_pair["ref"])
    refc = (small_pair["ref"][:, :3] - mu).astype(np.float32)
    q = o.transform_points(small_pair["T0"], small_pair["reading"])[:, :3] - mu
    ib, db = o.nn_brute(q, refc)
    ik, dk = o.nn_kdtree(q, refc, 2)
    assert np.array_equal(ib, ik) and np.array_equal(db, dk)


def test_nn_against_scipy(oracle_mod, small_pair):
    from scipy.spatial import cKDTree
    o = oracle_mod
    refc = small_pair["ref"][:, :3].copy()
    q = small_pair["reading"][:, :3].copy()
    ik, dk = o.nn_kdtree(q, refc)
    dist, idx = cKDTree(refc.astype(np.float64)).query(q.astype(np.float64), k=1)
    # scipy works in float64; equal indices wherever the float64 nearest is unique by a clear margin
    assert np.allclose(np.sqrt(dk), dist, rtol=1e-5, atol=1e-6)
    assert (ik == idx).mean() > 0.999


def test_tie_break_lowest_index(oracle_mod):
    o = oracle_mod
    ref = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [1, 0, 0]], np.float32)  # 0 and 4 coincide
    q = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    ib, db = o.nn_brute(q, ref)
    ik, dk = o.nn_kdtree(q, ref)
    assert list(ib) == [0, 0] and list(ik) == [0, 0]
    assert list(db) == [1.0, 0.0] and list(dk) == [1.0, 0.0]
    # many duplicates across kd-tree leaves
    rng = np.random.default_rng(1)
    base = rng.normal(size=(50, 3)).astype(np.float32)
    ref = np.concatenate([base] * 7)
    ib, _ = o.nn_brute(base, ref)
    ik, _ = o.nn_kdtree(base, ref)
    assert np.array_equal(ib, np.arange(50)) and np.array_equal(ik, ib)


def test_trim_limit_matches_numpy(oracle_mod):
    rng = np.random.default_rng(0)
    for n, ratio in [(1000, 0.75), (1001, 0.75), (17, 0.5), (5, 1.0), (1, 0.75), (4096, 0.1)]:
        d2 = rng.random(n).astype(np.float32)
        lim, nf = oracle_mod.trim_limit(d2, ratio)
        k = min(int(np.float32(n) * np.float32(ratio)), n - 1)
        assert nf == n and lim == np.sort(d2)[k]
    d2 = np.array([1, np.inf, 2, 3, np.inf], np.float32)
    lim, nf = oracle_mod.trim_limit(d2, 0.75)
    assert nf == 3 and lim == 3.0


def test_mean_is_correctly_rounded(oracle_mod):
    rng = np.random.default_rng(3)
    pts = np.ones((5000, 4), np.float32)
    pts[:, :3] = rng.normal(scale=30, size=(5000, 3)).astype(np.float32)
    mu = oracle_mod.mean(pts)
    assert np.allclose(mu, pts[:, :3].astype(np.float64).mean(0), rtol=0, atol=1e-6)
    perm = rng.permutation(5000)
    assert np.array_equal(mu, oracle_mod.mean(pts[perm]))  # order independent by construction


def test_normal_equations_and_solve(oracle_mod, small_pair):
    o = oracle_mod
    mu = o.mean(small_pair["ref"])
    refc = (small_pair["ref"][:, :3] - mu).astype(np.float32)
    Tpre = small_pair["T0"].copy()
    Tpre[:3, 3] -= mu
    step = o.transform_points(Tpre, small_pair["reading"])
    ids, d2 = o.nn_kdtree(step[:, :3].copy(), refc)
    lim, _ = o.trim_limit(d2, 0.75)
    A, b, kept, Ad, bd = o.normal_equations(step, refc, small_pair["ref_normals"], ids, d2, lim)
    keep = d2 <= lim
    assert kept == keep.sum()
    s = step[keep, :3].astype(np.float64)
    qn = refc[ids[keep]].astype(np.float64)
    nn = small_pair["ref_normals"][ids[keep]].astype(np.float64)
    F = np.concatenate([np.cross(s, nn), nn], 1)
    e = ((s - qn) * nn).sum(1)
    A64, b64 = F.T @ F, -(F.T @ e)
    assert np.allclose(A, A64, rtol=1e-5, atol=1e-3) and np.allclose(b, b64, rtol=1e-4, atol=1e-3)
    assert np.allclose(A, Ad, rtol=1e-6, atol=1e-2)  # fixed-point vs plain double accumulation of the same terms
    rc, T, x = o.solve_step(A, b)
    assert rc == 0 and np.allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)
    th = np.linalg.norm(x[:3])
    K = np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]]) / th
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    assert np.allclose(T[:3, :3], R, atol=1e-7) and np.allclose(T[:3, 3], x[3:], atol=1e-7)


def test_rank_deficient_solve_is_minimum_norm(oracle_mod):
    # all normals along z: only (rx, ry, tz) are observable; the minimum-norm solution leaves the rest at 0
    rng = np.random.default_rng(5)
    p = rng.uniform(-5, 5, (500, 3))
    n = np.tile([0, 0, 1.0], (500, 1))
    F = np.concatenate([np.cross(p, n), n], 1)
    e = rng.normal(scale=0.01, size=500)
    A, b = F.T @ F, -(F.T @ e)
    rc, T, x = oracle_mod.solve_step(A, b)
    assert rc == 0
    xp = np.linalg.pinv(A) @ b
    assert np.allclose(x, xp, atol=1e-9) and abs(x[2]) < 1e-12 and abs(x[3]) < 1e-12 and abs(x[4]) < 1e-12


def test_sincos(oracle_mod):
    for x in [0.0, 1e-9, 1e-3, 0.5, 0.78, 1.0, 2.0, 3.14, 6.0, 50.0, 1e3]:
        s, c = oracle_mod.sincos(x)
        assert abs(s - np.sin(x)) < 5e-15 * max(1.0, x) and abs(c - np.cos(x)) < 5e-15 * max(1.0, x)


def test_rigid_check_and_correct(oracle_mod):
    T = np.eye(4, dtype=np.float32)
    assert oracle_mod.check_rigid(T)
    T2 = T.copy()
    T2[:3, :3] *= 1.01
    assert not oracle_mod.check_rigid(T2)
    C = oracle_mod.correct_rigid(T2)
    assert np.allclose(C[:3, :3].T @ C[:3, :3], np.eye(3), atol=1e-6) and oracle_mod.check_rigid(C)


def test_icp_known_answer_rigid_copy(oracle_mod, small_pair):
    """A cloud against a rigidly moved copy of itself: ICP must recover the motion (analytic known answer)."""
    o = oracle_mod
    ref, nrm = small_pair["ref"], small_pair["ref_normals"]
    ang = np.deg2rad([0.4, -0.3, 0.9])
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [0.12, -0.07, 0.03]
    reading = o.transform_points(np.linalg.inv(T).astype(np.float32), ref)  # reading = T^-1 * ref  => T_ref<-reading = T
    r = o.icp(reading, ref, nrm, np.eye(4, dtype=np.float32), o.default_params(max_iterations=40, use_differential=0))
    assert r["rc"] == 0
    assert np.abs(r["T"][:3, 3] - T[:3, 3]).max() < 2e-4
    assert np.abs(r["T"][:3, :3] - T[:3, :3]).max() < 2e-5


def test_icp_identical_clouds_and_errors(oracle_mod, small_pair):
    o = oracle_mod
    ref, nrm = small_pair["ref"], small_pair["ref_normals"]
    r = o.icp(ref, ref, nrm, np.eye(4, dtype=np.float32))
    assert r["rc"] == 0 and np.allclose(r["T"], np.eye(4), atol=1e-6)  # theta == 0 -> rotation := I guard
    empty = np.zeros((0, 4), np.float32)
    assert o.icp(empty, ref, nrm, np.eye(4, dtype=np.float32))["rc"] == 1
    assert o.icp(ref, empty, np.zeros((0, 3), np.float32), np.eye(4, dtype=np.float32))["rc"] == 1


def test_icp_default_chain_converges_to_truth(oracle_mod, small_pair):
    r = oracle_mod.icp(small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], small_pair["T0"], want_hist=True)
    assert r["rc"] == 0 and r["stats"].converged == 1 and 4 <= r["stats"].iterations <= 40
    assert np.abs(r["T"][:3, 3] - small_pair["truth"][:3, 3]).max() < 0.02
    assert r["ids_hist"].shape[0] == r["stats"].iterations


def _independent_icp(reading4, ref4, nrm3, T0, iters, ratio=0.75):
    """The chain of SURVEY.md 8 'Oracle spec' restated a second time, in float64 numpy with scipy's kd-tree:
    shares no code with oracle/icp_oracle.cpp."""
    from scipy.spatial import cKDTree
    from scipy.spatial.transform import Rotation
    ref = ref4[:, :3].astype(np.float64)
    mu = ref.mean(0)
    Q = ref - mu
    tree = cKDTree(Q)
    Tpre = np.asarray(T0, np.float64).copy()
    Tpre[:3, 3] -= mu
    R = reading4[:, :3].astype(np.float64) @ Tpre[:3, :3].T + Tpre[:3, 3]
    T = np.eye(4)
    n = len(R)
    for _ in range(iters):
        S = R @ T[:3, :3].T + T[:3, 3]
        d, idx = tree.query(S)
        d2 = d * d
        k = int(n * ratio)
        limit = np.partition(d2, k)[k]
        keep = d2 <= limit
        s, q, nn = S[keep], Q[idx[keep]], nrm3[idx[keep]].astype(np.float64)
        F = np.hstack([np.cross(s, nn), nn])
        e = ((s - q) * nn).sum(1)
        x = np.linalg.solve(F.T @ F, -F.T @ e)
        step = np.eye(4)
        step[:3, :3] = Rotation.from_rotvec(x[:3]).as_matrix()
        step[:3, 3] = x[3:]
        T = step @ T
    Tm = np.eye(4)
    Tm[:3, 3] = mu
    return Tm @ T @ Tpre


def test_icp_against_an_independent_float64_restatement(oracle_mod, small_pair):
    """Pins the oracle's whole chain (not only its pieces) against a second implementation: after a fixed number of
    iterations both must give the same pose to the north-star tolerance (1e-4 m, 1e-5 rad)."""
    o = oracle_mod
    for iters in (1, 5, 25):
        r = o.icp(small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], small_pair["T0"],
                  o.default_params(max_iterations=iters, use_differential=0))
        T = _independent_icp(small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], small_pair["T0"], iters)
        assert r["rc"] == 0 and r["stats"].iterations == iters
        assert np.abs(r["T"][:3, 3] - T[:3, 3]).max() < 1e-4, iters
        dR = r["T"][:3, :3].astype(np.float64) @ T[:3, :3].T
        ang = np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2
        assert ang < 1e-5, (iters, ang)


def test_icp_against_the_independent_restatement_at_full_size(oracle_mod, config2):
    """Same cross-check on BASELINE.json's configuration (131072-point scan vs 524288-point 4-scan map)."""
    o = oracle_mod
    iters = 6
    r = o.icp(config2["reading"], config2["ref"], config2["ref_normals"], config2["T0"],
              o.default_params(max_iterations=iters, use_differential=0, num_threads=8))
    T = _independent_icp(config2["reading"], config2["ref"], config2["ref_normals"], config2["T0"], iters)
    assert r["rc"] == 0 and np.abs(r["T"][:3, 3] - T[:3, 3]).max() < 1e-4
    dR = r["T"][:3, :3].astype(np.float64) @ T[:3, :3].T
    assert np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / 2 < 1e-5


def test_golden_vectors(oracle_mod, synth_mod):
    """Committed golden vectors (tests/golden/make_golden.py): the oracle on the STORED seeded inputs must reproduce the
    stored outputs bit for bit (its arithmetic is operation-order fixed and compiled with -ffp-contract=off, so this
    holds across compilers and hosts).  The generator is only required to reproduce the stored inputs to float32
    rounding: it goes through libm's sin/cos, whose last bit differs between libm builds and CPU dispatch variants."""
    g = np.load(os.path.join(GOLDEN, "icp_small.npz"))
    r = oracle_mod.icp(g["reading"], g["ref"], g["ref_normals"], g["T0"],
                       oracle_mod.default_params(max_iterations=int(g["max_iterations"]),
                                                 use_differential=int(g["use_differential"])), want_hist=True)
    assert np.array_equal(r["T"], g["T"]) and np.array_equal(r["ids_hist"][-1], g["ids_last"])
    assert np.array_equal(r["T_iter_hist"], g["T_iter_hist"]) and np.array_equal(r["d2_last"], g["d2_last"])
    assert np.array_equal(np.array([zlib_crc(x) for x in r["ids_hist"]], np.uint32), g["ids_crc"])
    truth, odom = synth_mod.trajectory(int(g["seq"]), 3)
    a, an = synth_mod.subsample(*synth_mod.scan(truth[0], int(g["seq"]), 0), int(g["step"]))
    b, _ = synth_mod.subsample(*synth_mod.scan(truth[1], int(g["seq"]), 1), int(g["step"]))
    assert np.allclose(a, g["ref"], rtol=0, atol=2e-5) and np.allclose(b, g["reading"], rtol=0, atol=2e-5)
    assert np.allclose(an, g["ref_normals"], rtol=0, atol=1e-6)


def zlib_crc(a):
    import zlib
    return zlib.crc32(np.ascontiguousarray(a).tobytes())
