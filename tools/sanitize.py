"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): the smoke registration plus one batched launch of four
small scan-to-sub-map problems, a normals estimate, a pose-graph solve with marginals and the input-side kernels -- every kernel family of the
library on inputs small enough for the tool's ~100x slow-down.  Results are still checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import laser_slam_b200 as ls
from laser_slam_b200 import synth
import oracle
from oracle import posegraph_oracle as pg

truth, odom = synth.trajectory(0, 6)
sc = [synth.subsample(*synth.scan(truth[k], 0, k), 16) for k in range(6)]
ctx = ls.Context(0)
T0 = (np.linalg.inv(truth[0]) @ odom[1]).astype(np.float32)
p = ls.default_params(max_iterations=8, use_differential=0)
g = ctx.icp_register(sc[1][0], sc[0][0], sc[0][1], T0, p, want_ids=True, want_hist=True)
r = oracle.icp(sc[1][0], sc[0][0], sc[0][1], T0, oracle.default_params(max_iterations=8, use_differential=0), want_hist=True)
assert np.array_equal(g["T"], r["T"]) and np.array_equal(g["ids"], r["ids_hist"][-1])
mp = ctx.create_map(8, 8192)
sid = [mp.push_scan(*sc[k]) for k in range(6)]
probs = []
for ref, rd, ks in [(3, 4, [3, 2, 1, 0]), (4, 5, [4, 3, 2]), (2, 3, [2, 1]), (1, 2, [1, 0])]:
    Ts = [np.eye(4, dtype=np.float32) if k == ref else (np.linalg.inv(truth[ref]) @ truth[k]).astype(np.float32) for k in ks]
    probs.append((sid[rd], [sid[k] for k in ks], Ts, (np.linalg.inv(truth[ref]) @ odom[rd]).astype(np.float32)))
batch = mp.register_batch(probs, p)
single = [mp.register(*pr, p) for pr in probs]
assert all(np.array_equal(b["T"], s["T"]) for b, s in zip(batch, single))
nr = ctx.estimate_normals(sc[0][0][:2048], knn=10)
assert np.array_equal(nr, oracle.knn_normals(sc[0][0][:2048], 10))
keys, init, factors, _ = pg.make_config4(n_poses=60, n_lc=4, lap=20, seed=2)
G = ls.PoseGraph(0)
G.add_poses(keys, init)
G.add_factors(factors)
G.optimize(3)
cov = G.marginals(keys[:10])
assert np.isfinite(cov).all()
# input side (ls_filters.cu): ingest, cylinder, voxel grid, de-skew
pts = sc[0][0][:4096]
rec = np.zeros((len(pts), 8), np.float32)
rec[:, :3] = pts[:, :3]
assert np.array_equal(ls.ingest_pointcloud2(rec.tobytes(), 32, 0, 4, 8, len(pts)), pts)
assert np.array_equal(ls.filter_cylinder(pts, [0, 0, 0], 15.0, 6.0, False), oracle.filter_cylinder(pts, [0, 0, 0], 15.0, 6.0, False))
assert np.array_equal(ls.voxel_grid(pts, 0.5), oracle.voxel_grid(pts, 0.5))
offs = [0, 100, 100, 1500, 4096]
Tp = [np.eye(4, dtype=np.float32)] + [(np.linalg.inv(truth[0]) @ truth[k]).astype(np.float32) for k in (1, 2, 3)]
Tf = oracle.rigid_inverse_f32(Tp[3])
assert np.array_equal(ls.deskew_revolution(pts, offs, Tp, Tf), oracle.deskew_revolution(pts, offs, Tp, Tf))
print("sanitize workload ok:", g["stats"].iterations, "iterations;", len(batch), "batched problems;", ctx.launch_count, "launches")
