"""CPU-sim statistics of the grid query on config-2-shaped data (tuning aid; tests/sim, not product)."""
import ctypes, numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laser_slam_b200 import synth
import oracle
L = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'sim', 'libgrid_sim.so'))
vp = ctypes.c_void_p
L.sim_nn.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp, ctypes.c_float]

def sim_nn(q, refc, cell=1.0, max_cells=1 << 22, split=32, warm=None, cap=np.inf):
    q = np.ascontiguousarray(q, np.float32); refc = np.ascontiguousarray(refc, np.float32)
    ids = np.empty(len(q), np.int32); d2 = np.empty(len(q), np.float32); st = np.zeros(8)
    pc = np.empty(len(q), np.int32); pe = np.empty(len(q), np.int32)
    w = np.ascontiguousarray(warm, np.int32) if warm is not None else None
    L.sim_nn(q.ctypes.data, len(q), refc.ctypes.data, len(refc), cell, max_cells, split,
             w.ctypes.data if w is not None else None, ids.ctypes.data, d2.ctypes.data, st.ctypes.data, pc.ctypes.data, pe.ctypes.data, cap)
    global LAST_STEPS
    LAST_STEPS = pe >> 16
    pe = pe & 0xffff
    return ids, d2, st, pc, pe

if __name__ == "__main__":
    truth, odom = synth.trajectory(0, 8)
    scans = [synth.scan(truth[k], 0, k) for k in range(6)]
    ref = np.concatenate([oracle.transform_cloud(np.linalg.inv(truth[3]) @ truth[k], *scans[k])[0] for k in [3, 2, 1, 0]])
    mu = oracle.mean(ref); refc = (ref[:, :3] - mu).astype(np.float32)
    T0 = np.linalg.inv(truth[3]) @ odom[4]; T0[:3, 3] -= mu
    q = oracle.transform_points(T0, scans[4][0])[:, :3].copy()
    ik, dk = oracle.nn_kdtree(q, refc, 8)
    q2 = (q + np.float32(0.004)).astype(np.float32)
    ik2, dk2 = oracle.nn_kdtree(q2, refc, 8)
    for cell, split in [(1.0, 32), (2.0, 32), (1.0, 16)]:
        i1, d1, st, pc, pe = sim_nn(q, refc, cell, split=split)
        print(f"cell {cell} split {split} COLD eq {(i1 == ik).all()} {(d1 == dk).all()} cand mean {pc.mean():.1f} pct50/99/99.9/max {np.percentile(pc, [50, 99, 99.9, 100])} entries mean {pe.mean():.1f} {np.percentile(pe, [50, 99, 99.9, 100])}")
        i2, d2_, st2, pc, pe = sim_nn(q2, refc, cell, split=split, warm=ik)
        print(f"cell {cell} split {split} WARM eq {(i2 == ik2).all()} {(d2_ == dk2).all()} cand mean {pc.mean():.1f} pct50/99/99.9/max {np.percentile(pc, [50, 99, 99.9, 100])} entries mean {pe.mean():.1f} {np.percentile(pe, [50, 99, 99.9, 100])}")
        cost = (pc + pe).reshape(-1, 32)
        print("   per-warp max mean", cost.max(1).mean(), "sum mean", cost.sum(1).mean(), "sum max", cost.sum(1).max(), "max max", cost.max())
        lim = np.sort(dk2)[int(len(dk2) * 0.75)]
        for mult in (4.0, 16.0):
            i3, d3, st3, pc, pe = sim_nn(q2, refc, cell, split=split, warm=ik, cap=float(lim * mult))
            ok = d3 <= lim
            print(f"   CAP {mult}x limit ({lim*mult:.4f}): found {np.isfinite(d3).mean():.3f} kept-set eq {np.array_equal(i3[dk2<=lim], ik2[dk2<=lim])} cand mean {pc.mean():.1f} pct99/99.9/max {np.percentile(pc,[99,99.9,100])} entries mean {pe.mean():.1f} {np.percentile(pe,[99,99.9,100])}")
            cost = (pc + pe).reshape(-1, 32)
            print("      per-warp max mean", cost.max(1).mean(), "sum mean", cost.sum(1).mean(), "sum max", cost.sum(1).max(), "max max", cost.max())
            st_ = LAST_STEPS.reshape(-1, 32)
            print(f"      STEPS mean {LAST_STEPS.mean():.1f} pct50/99/max {np.percentile(LAST_STEPS,[50,99,100])} per-warp max: mean {st_.max(1).mean():.1f} max {st_.max()}")
            i4, d4, st4, pc, pe = sim_nn(q, refc, cell, split=split, warm=None, cap=float(lim * mult))
            print(f"      COLD capped: found {np.isfinite(d4).mean():.3f} cand mean {pc.mean():.1f} pct99/99.9/max {np.percentile(pc,[99,99.9,100])} entries mean {pe.mean():.1f} {np.percentile(pe,[99,99.9,100])}")
