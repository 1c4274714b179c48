"""Config-2 registration on resident scans, a few repetitions (profiling / phase-timing driver).

    python tools/prof_one.py [reps] [iterations]
    LS_BATCH=8            also time an 8-problem cooperative launch
    LS_PROF_Y=-6 LS_PROF_SCAN=14   where along the synthetic street (bench.py's pool starts at y = -20 m and its scans
                          14-23 pass the dense facades that make its slowest steps; the default here is the sparser
                          start of sequence 0)
"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import laser_slam_b200 as ls
from laser_slam_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = ls.Context(0)
k0 = int(os.environ.get("LS_PROF_SCAN", "0"))
SENSOR = int(os.environ.get("LS_PROF_SENSOR", "0"))   # 0 HDL-64 (131072 rays), 1 VLS-128 (262144 rays): config 5 with LS_PROF_K=8
K = int(os.environ.get("LS_PROF_K", "4"))
if "LS_PROF_Y" in os.environ or k0:
    truth, odom = synth.trajectory(0, k0 + K + 4, y_start=float(os.environ.get("LS_PROF_Y", "-20")))
    truth, odom = truth[k0:], odom[k0:]
    scans = [synth.scan(truth[k], 0, k0 + k, sensor=SENSOR) for k in range(K + 1)]
else:
    truth, odom = synth.trajectory(0, K + 4)
    scans = [synth.scan(truth[k], 0, k, sensor=SENSOR) for k in range(K + 1)]
mp = ctx.create_map(K + 4, len(scans[0][0]))
sid = [mp.push_scan(*scans[k]) for k in range(K + 1)]
ref_order = [K - 1 - j for j in range(K)]
Tparts = [np.eye(4, dtype=np.float32) if k == K - 1 else (np.linalg.inv(truth[K - 1]) @ truth[k]).astype(np.float32) for k in ref_order]
T0 = (np.linalg.inv(truth[K - 1]) @ odom[K]).astype(np.float32)
p = ls.default_params(max_iterations=iters, use_differential=0)
for k in ("cell_size", "leaf_split"):
    if os.environ.get("LS_" + k.upper()):
        setattr(p, k, type(getattr(p, k))(float(os.environ["LS_" + k.upper()])))
for rep in range(reps):
    g = mp.register(sid[K], [sid[k] for k in ref_order], Tparts, T0, p)
    print(f"rep {rep}: device {g['stats'].device_ms:.3f} ms build {g['stats'].build_ms:.3f} ms iters {g['stats'].iterations}")

B = int(os.environ.get("LS_BATCH", "0"))
if B:
    import time
    probs = [(sid[K], [sid[k] for k in ref_order], Tparts, T0)] * B
    call = mp.prepare_batch(probs, p)
    for rep in range(reps + 1):
        t0 = time.perf_counter(); rc, statuses, touts, stats = call(); dt = time.perf_counter() - t0
        print(f"batch {B} rep {rep}: wall {dt*1e3:.3f} ms -> {dt*1e3/B:.3f} ms/registration, device {stats[0].device_ms:.3f} ms, rc {rc} {list(statuses)}")
