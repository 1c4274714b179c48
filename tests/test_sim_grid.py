"""CPU simulation of the CUDA spatial-hash query (tests/sim/grid_sim.cpp compiles the same
ls_grid.cuh the kernels use) against the oracle's brute force: exactness, tie-break, edge cases.
This checks the query LOGIC where there is no GPU; the kernels themselves are checked by -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SRC = os.path.join(HERE, "sim", "grid_sim.cpp")
SIM_LIB = os.path.join(HERE, "sim", "libgrid_sim.so")


@pytest.fixture(scope="module")
def sim():
    deps = [SIM_SRC] + [os.path.join(HERE, "..", "laser_slam_b200", "csrc", f) for f in ("ls_grid.cuh", "ls_math.cuh")]
    if not os.path.exists(SIM_LIB) or os.path.getmtime(SIM_LIB) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-o", SIM_LIB, SIM_SRC])
    L = ctypes.CDLL(SIM_LIB)
    vp = ctypes.c_void_p
    L.sim_nn.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, ctypes.c_float]

    def run(q, ref, cell=1.0, max_cells=1 << 22, split=32, warm=None, cap=np.inf):
        q = np.ascontiguousarray(q, np.float32)
        ref = np.ascontiguousarray(ref, np.float32)
        ids = np.empty(max(len(q), 1), np.int32)
        d2 = np.empty(max(len(q), 1), np.float32)
        w = np.ascontiguousarray(warm, np.int32) if warm is not None else None
        L.sim_nn(q.ctypes.data, len(q), ref.ctypes.data, len(ref), cell, max_cells, split,
                 w.ctypes.data if w is not None else None, ids.ctypes.data, d2.ctypes.data, None, None, None, cap)
        return ids[:len(q)], d2[:len(q)]
    return run


def _check(sim, oracle_mod, q, ref, **kw):
    ib, db = oracle_mod.nn_brute(q, ref)
    ig, dg = sim(q, ref, **kw)
    assert np.array_equal(ib, ig), f"{(ib != ig).sum()} index mismatches"
    assert np.array_equal(db, dg)


@pytest.mark.parametrize("cell,split", [(1.0, 32), (2.0, 16), (0.25, 16), (8.0, 64)])
def test_sim_matches_brute_force_on_lidar(sim, oracle_mod, small_pair, cell, split):
    mu = oracle_mod.mean(small_pair["ref"])
    refc = (small_pair["ref"][:, :3] - mu).astype(np.float32)
    q = (oracle_mod.transform_points(small_pair["T0"], small_pair["reading"])[:, :3] - mu).astype(np.float32)
    _check(sim, oracle_mod, q, refc, cell=cell, split=split)


def test_sim_warm_start_any_seed_is_exact(sim, oracle_mod, small_pair):
    rng = np.random.default_rng(0)
    refc = small_pair["ref"][:, :3].copy()
    q = small_pair["reading"][:, :3].copy()
    ib, db = oracle_mod.nn_brute(q, refc)
    for warm in (ib, rng.integers(0, len(refc), len(q)), np.zeros(len(q), np.int64)):
        ig, dg = sim(q, refc, warm=warm)
        assert np.array_equal(ib, ig) and np.array_equal(db, dg)


def test_sim_ties_duplicates_and_lattices(sim, oracle_mod):
    rng = np.random.default_rng(2)
    # integer lattice: massive exact ties, queries at cell centres and on lattice points
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ref = np.concatenate([g, g[rng.permutation(len(g))[:300]]])  # plus duplicates
    q = np.concatenate([g + 0.5, g, rng.uniform(-3, 15, (500, 3)).astype(np.float32)]).astype(np.float32)
    for cell, split in [(1.0, 16), (2.0, 32), (0.5, 16)]:
        _check(sim, oracle_mod, q, ref, cell=cell, split=split)


def test_sim_edge_cases(sim, oracle_mod):
    rng = np.random.default_rng(4)
    one = np.array([[1.5, -2.0, 0.25]], np.float32)
    q = rng.normal(scale=10, size=(64, 3)).astype(np.float32)
    _check(sim, oracle_mod, q, one)                                     # single map point
    same = np.repeat(one, 500, 0)
    _check(sim, oracle_mod, q, same)                                    # all points identical (one huge leaf)
    line = np.zeros((2000, 3), np.float32)
    line[:, 0] = np.linspace(-400, 400, 2000)
    _check(sim, oracle_mod, q, line, cell=1.0)                          # degenerate extent, far queries
    far = (rng.normal(size=(200, 3)) * [500, 500, 50]).astype(np.float32)
    cloud = rng.normal(scale=3, size=(5000, 3)).astype(np.float32)
    _check(sim, oracle_mod, far, cloud)                                 # queries far outside the bounding box
    ids, d2 = sim(q, np.zeros((0, 3), np.float32))                      # empty map: libnabo's "unfound"
    assert (ids == -1).all() and np.isinf(d2).all()
    dense = (rng.normal(scale=0.01, size=(20000, 3))).astype(np.float32)  # everything in one fine cell
    _check(sim, oracle_mod, rng.normal(scale=0.02, size=(300, 3)).astype(np.float32), dense)


def test_sim_small_cell_budget_grows_cells(sim, oracle_mod, small_pair):
    refc = small_pair["ref"][:, :3].copy()
    q = small_pair["reading"][:2000, :3].copy()
    _check(sim, oracle_mod, q, refc, cell=0.25, max_cells=4096)  # H0 doubles until the dense array fits


def test_sim_capped_search_is_exact_within_cap(sim, oracle_mod, small_pair):
    """With a finite cap the query returns the exact NN when d2 <= cap and 'not found' otherwise."""
    refc = small_pair["ref"][:, :3].copy()
    q = small_pair["reading"][:, :3].copy()
    ib, db = oracle_mod.nn_brute(q, refc)
    for cap in (0.0004, 0.01, 0.25, 4.0):
        for warm in (None, ib, np.zeros(len(q), np.int64)):
            ig, dg = sim(q, refc, warm=warm, cap=cap)
            inside = db <= np.float32(cap)
            assert np.array_equal(ig[inside], ib[inside]) and np.array_equal(dg[inside], db[inside])
            assert (ig[~inside] == -1).all() and np.isinf(dg[~inside]).all()


# ---- certified candidate lists (ls_grid.cuh vlist_query / vlist_build): answers must equal the plain search ------
@pytest.fixture(scope="module")
def sim_lists(sim):
    L = ctypes.CDLL(SIM_LIB)
    vp = ctypes.c_void_p
    L.sim_vlists.restype = ctypes.c_int
    L.sim_vlists.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int,
                             vp, vp, vp, vp, vp, vp, vp]

    def run(rd3, refc3, T_seq, caps, skin_w, skin_v, cell=1.0, split=32):
        rd3 = np.ascontiguousarray(rd3, np.float32)
        refc3 = np.ascontiguousarray(refc3, np.float32)
        Tcm = np.ascontiguousarray(np.stack([np.asarray(T, np.float32).T.ravel() for T in T_seq]))   # column-major
        K = len(T_seq)
        caps, skin_w, skin_v = (np.ascontiguousarray(a, np.float32) for a in (caps, skin_w, skin_v))
        hits, over = np.zeros(K, np.int32), np.zeros(K, np.int32)
        ids, d2 = np.empty(len(rd3), np.int32), np.empty(len(rd3), np.float32)
        bad = L.sim_vlists(rd3.ctypes.data, len(rd3), refc3.ctypes.data, len(refc3), cell, 1 << 22, split, Tcm.ctypes.data, K,
                           caps.ctypes.data, skin_w.ctypes.data, skin_v.ctypes.data, hits.ctypes.data, over.ctypes.data,
                           ids.ctypes.data, d2.ctypes.data)
        return bad, hits, over, ids, d2
    return run


def _pose_sequence(T_hist):
    """T_iter before every iteration (identity first) + the motion bound of the step that led to it."""
    Ts = [np.eye(4, dtype=np.float32)] + [np.asarray(T, np.float32) for T in T_hist[:-1]]
    w, v = [0.0], [0.0]
    for a, b in zip(Ts[:-1], Ts[1:]):
        S = b.astype(np.float64) @ np.linalg.inv(a.astype(np.float64))
        w.append(float(np.arccos(np.clip((np.trace(S[:3, :3]) - 1) / 2, -1, 1))))
        v.append(float(np.linalg.norm(S[:3, 3])))
    return Ts, w, v


def test_sim_lists_reproduce_the_search_over_an_icp_run(sim_lists, oracle_mod, small_pair):
    """Every query of every iteration of a real ICP run: the list path and the plain search give the same match."""
    o = oracle_mod
    r = o.icp(small_pair["reading"], small_pair["ref"], small_pair["ref_normals"], small_pair["T0"],
              o.default_params(max_iterations=20, use_differential=0), want_hist=True)
    mu = o.mean(small_pair["ref"])
    refc = (small_pair["ref"][:, :3] - mu).astype(np.float32)
    Tpre = small_pair["T0"].copy()
    Tpre[:3, 3] -= mu
    rd = o.transform_points(Tpre, small_pair["reading"])[:, :3].copy()
    Ts, w, v = _pose_sequence(r["T_iter_hist"])
    for caps in ([0.25] + [0.02] * 19, [0.25] + [0.004] * 19, [np.inf] * 20):
        bad, hits, over, ids, d2 = sim_lists(rd, refc, Ts, caps, w, v)
        assert bad == 0
        assert hits[0] == 0 and hits[-1] > 0.9 * len(rd), hits          # converged: almost every query is certified
    assert np.array_equal(ids, r["ids_hist"][-1])                       # last iteration == the oracle's correspondences


def test_sim_lists_ties_and_tiny_motions(sim_lists):
    """Lattice + duplicates (exact ties inside the lists), queries drifting by fractions of the spacing."""
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.05
    ref = np.concatenate([g, g[rng.permutation(len(g))[:100]]]).astype(np.float32)
    q = np.concatenate([g + 0.025, g, rng.uniform(-0.1, 0.6, (400, 3)).astype(np.float32)]).astype(np.float32)
    Ts = []
    for t in range(12):
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = [0.0004 * t, -0.0002 * t, 0.0001 * (t % 3)]
        Ts.append(T)
    for cap in (np.inf, 0.01, 0.0007):
        bad, hits, over, _, _ = sim_lists(q, ref, Ts, [cap] * 12, [0.0] * 12, [0.0005] * 12, cell=0.5, split=16)
        assert bad == 0 and hits[1:].sum() > 0
