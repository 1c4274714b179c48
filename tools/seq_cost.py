"""Single-registration device time of the bench workload per synthetic sequence (tuning aid)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import laser_slam_b200 as ls
from laser_slam_b200 import synth
import bench

ctx = ls.Context(0)
p = ls.default_params(max_iterations=30, use_differential=0)
lo, hi = int(sys.argv[1]), int(sys.argv[2])
for seq in range(lo, hi):
    truth, odom = synth.trajectory(seq, 8, y_start=-20.0)
    scans = [synth.scan(truth[k], seq, k) for k in range(6)]
    mp = ctx.create_map(8, 131072)
    sid = [mp.push_scan(*scans[k]) for k in range(6)]
    res = []
    for n in (4, 5):
        ref = n - 1
        ks = [ref - j for j in range(4) if ref - j >= 0]
        Ts = [np.eye(4, dtype=np.float32) if k == ref else (np.linalg.inv(truth[ref]) @ truth[k]).astype(np.float32) for k in ks]
        T0 = (np.linalg.inv(truth[ref]) @ odom[n]).astype(np.float32)
        for rep in range(2):
            g = mp.register(sid[n], [sid[k] for k in ks], Ts, T0, p)
        st = g["stats"]
        terr = np.abs(g["T"][:3, 3] - (np.linalg.inv(truth[ref]) @ truth[n])[:3, 3]).max()
        t0err = np.abs(T0[:3, 3] - (np.linalg.inv(truth[ref]) @ truth[n])[:3, 3]).max()
        res.append(f"n={n}: {st.device_ms:.2f} ms (build {st.build_ms:.2f}) kept {st.last_kept} limit {st.last_limit:.4f} tables {st.grid_tables} T0err {t0err:.3f} err {terr:.4f}")
    print(f"seq {seq}: " + " | ".join(res), flush=True)
    del mp
