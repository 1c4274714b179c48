"""Generates tests/golden/icp_small.npz from the oracle on seeded synthetic inputs.

The reference holds no golden vectors for this path (reference laser_slam/test/test_empty.cpp:3-5) and
cannot be run here, so these vectors pin the ORACLE (and, through the -m gpu tests, the CUDA path)
against regressions; they are not outputs of the reference.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from laser_slam_b200 import synth  # noqa: E402

seq, step, max_iterations, use_differential = 0, 32, 12, 0
truth, odom = synth.trajectory(seq, 3)
a, an = synth.subsample(*synth.scan(truth[0], seq, 0), step)
b, bn = synth.subsample(*synth.scan(truth[1], seq, 1), step)
T0 = (np.linalg.inv(truth[0]) @ odom[1]).astype(np.float32)
r = oracle.icp(b, a, an, T0, oracle.default_params(max_iterations=max_iterations, use_differential=use_differential),
               want_hist=True)
assert r["rc"] == 0
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp_small.npz"),
                    seq=seq, step=step, max_iterations=max_iterations, use_differential=use_differential,
                    ref=a, ref_normals=an, reading=b, T0=T0, T=r["T"], ids_last=r["ids_hist"][-1],
                    T_iter_hist=r["T_iter_hist"], d2_last=r["d2_last"],
                    ids_crc=np.array([zlib.crc32(np.ascontiguousarray(x).tobytes()) for x in r["ids_hist"]], np.uint32))
print("wrote icp_small.npz:", a.shape, b.shape, "iterations", r["stats"].iterations)
