"""Debug driver: B distinct config-2 problems in one launch (the shape of tests/test_gpu_icp.py's batch test).
    python tools/dbg_batch.py B iterations [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import laser_slam_b200 as ls
from laser_slam_b200 import synth
import oracle
import importlib.util
spec = importlib.util.spec_from_file_location("tg", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "test_gpu_icp.py"))
tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
B, iters = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = ls.Context(0)
probs = [tg._track_problem(oracle, synth, seq=b % 8, k0=3 * (b // 8) + (b % 3)) for b in range(B)]
mp = ctx.create_map(5 * B, 131072)
staged = []
for pr in probs:
    sid = {k: mp.push_scan(*pr["scans"][k]) for k in pr["scans"]}
    staged.append((sid[pr["reading"]], [sid[k] for k in pr["ks"]], pr["Ts"], pr["T0"]))
pg = ls.default_params(max_iterations=iters, use_differential=0)
for rep in range(reps):
    got = mp.register_batch(staged, pg)
    print("rep", rep, [g["rc"] for g in got], [g["stats"].last_kept for g in got][:4], flush=True)
if os.environ.get("CHECK"):
    po = oracle.default_params(max_iterations=iters, use_differential=0, num_threads=16)
    for b, pr in enumerate(probs):
        r = oracle.icp(pr["scans"][pr["reading"]][0], pr["refp"], pr["refn"], pr["T0"], po)
        print(b, np.array_equal(got[b]["T"], r["T"]), got[b]["stats"].last_kept == r["stats"].last_kept, flush=True)
