"""First-contact GPU check: NN parity, ICP parity vs oracle, rough timing.  Run under gpurun."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import laser_slam_b200 as ls
from laser_slam_b200 import synth
import oracle

ctx = ls.Context(0)
truth, odom = synth.trajectory(0, 8)
scans = [synth.scan(truth[k], 0, k) for k in range(6)]

def submap(kref, ks):
    ref, nr = [], []
    for k in ks:
        T = np.linalg.inv(truth[kref]) @ truth[k]
        if k == kref:
            p, n = scans[k]
        else:
            p, n = oracle.transform_cloud(T.astype(np.float32), scans[k][0], scans[k][1])
        ref.append(p); nr.append(n)
    return np.concatenate(ref), np.concatenate(nr)

# ---- small NN parity
a, an = synth.subsample(*scans[0], 8)
b, bn = synth.subsample(*scans[1], 8)
T0 = (np.linalg.inv(truth[0]) @ odom[1]).astype(np.float32)
mu = oracle.mean(a)
refc = (a[:, :3] - mu).astype(np.float32)
Tpre = T0.copy(); Tpre[:3, 3] = T0[:3, 3] - mu
q = oracle.transform_points(Tpre, b)[:, :3].copy()
ik, dk = oracle.nn_kdtree(q, refc, 8)
ig, dg = ctx.nn_query(b, a, T0)
print("small nn: ids eq", (ig == ik).all(), "d2 eq", (dg == dk).all(), "mismatch", int((ig != ik).sum()))

# ---- small ICP parity
p = ls.default_params()
po = oracle.default_params(num_threads=8)
r = oracle.icp(b, a, an, T0, po, want_hist=True)
g = ctx.icp_register(b, a, an, T0, p, want_ids=True, want_hist=True, raise_on_convergence=False)
print("small icp: oracle iters", r['stats'].iterations, "gpu iters", g['stats'].iterations, "conv", r['stats'].converged, g['stats'].converged)
print("  T equal bitwise:", np.array_equal(r['T'], g['T']), "max abs diff", np.abs(r['T'] - g['T']).max())
nh = min(len(r['T_iter_hist']), len(g['T_iter_hist']))
for it in range(nh):
    if not np.array_equal(r['T_iter_hist'][it], g['T_iter_hist'][it]):
        print("  first T_iter mismatch at iter", it, np.abs(r['T_iter_hist'][it] - g['T_iter_hist'][it]).max()); break
else:
    print("  T_iter history bit-equal over", nh, "iterations")
print("  last ids equal:", np.array_equal(r['ids_hist'][-1], g['ids']), "kept", r['stats'].last_kept, g['stats'].last_kept, "limit", r['stats'].last_limit, g['stats'].last_limit)

# ---- full-size config 2
ref, nr = submap(3, [3, 2, 1, 0])
rd = scans[4][0]
T0 = (np.linalg.inv(truth[3]) @ odom[4]).astype(np.float32)
p2 = ls.default_params(max_iterations=30, use_differential=0)
po2 = oracle.default_params(max_iterations=30, use_differential=0, num_threads=8)
t = time.time(); r = oracle.icp(rd, ref, nr, T0, po2, want_hist=True); t_or = time.time() - t
t = time.time(); g = ctx.icp_register(rd, ref, nr, T0, p2, want_ids=True, want_hist=True); t_g = time.time() - t
print(f"full icp: oracle {t_or:.2f}s (8 thr), gpu call {t_g*1e3:.1f} ms, device {g['stats'].device_ms:.3f} ms build {g['stats'].build_ms:.3f} ms")
print("  grid cells", g['stats'].grid_cells, "tables", g['stats'].grid_tables, "overflow", g['stats'].grid_overflow)
print("  T equal bitwise:", np.array_equal(r['T'], g['T']), "max abs diff", np.abs(r['T'] - g['T']).max())
for it in range(30):
    if not np.array_equal(r['T_iter_hist'][it], g['T_iter_hist'][it]):
        print("  first T_iter mismatch at iter", it, np.abs(r['T_iter_hist'][it] - g['T_iter_hist'][it]).max()); break
else:
    print("  T_iter history bit-equal over 30 iterations")
print("  last ids equal:", np.array_equal(r['ids_hist'][-1], g['ids']), "mismatches", int((r['ids_hist'][-1] != g['ids']).sum()))
print("  pose err vs truth (m):", np.abs(g['T'][:3, 3] - (np.linalg.inv(truth[3]) @ truth[4])[:3, 3]).max())
for rep in range(5):
    g = ctx.icp_register(rd, ref, nr, T0, p2)
    print(f"  rep {rep}: device {g['stats'].device_ms:.3f} ms (build {g['stats'].build_ms:.3f})")

# ---- resident map path
mp = ctx.create_map(8, 131072)
sid = [mp.push_scan(*scans[k]) for k in range(5)]
Tparts = []
for k in [3, 2, 1, 0]:
    T = (np.linalg.inv(truth[3]) @ truth[k]).astype(np.float32)
    Tparts.append(np.eye(4, dtype=np.float32) if k == 3 else T)
g2 = mp.register(sid[4], [sid[3], sid[2], sid[1], sid[0]], Tparts, T0, p2, want_ids=True)
print("submap path: T equal to one-shot oracle bitwise:", np.array_equal(r['T'], g2['T']), "ids eq", np.array_equal(r['ids_hist'][-1], g2['ids']))
for rep in range(5):
    g2 = mp.register(sid[4], [sid[3], sid[2], sid[1], sid[0]], Tparts, T0, p2)
    print(f"  rep {rep}: device {g2['stats'].device_ms:.3f} ms (build {g2['stats'].build_ms:.3f})")
print("launches", ctx.launch_count)
