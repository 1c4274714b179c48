"""ctypes front-end of the synthetic Velodyne-shaped workload generator (laser_slam_b200/synth/synth.cpp).

CPU-only test/bench utility; it produces the libpointmatcher DataPoints layout the
reference hands to LaserTrack (reference laser_slam/include/laser_slam/common.hpp:113-120):
features 4xN column-major float32 (x,y,z,1) and a 3xN `normals` descriptor.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libls_synth.so")
_lib = None

HDL64, VLS128 = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "synth", "synth.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", _LIB_PATH, src])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.ls_synth_num_rays.restype = ctypes.c_int
        lib.ls_synth_num_rays.argtypes = [ctypes.c_int]
        lib.ls_synth_trajectory.restype = None
        lib.ls_synth_trajectory.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_void_p, ctypes.c_void_p]
        lib.ls_synth_scan.restype = ctypes.c_int
        lib.ls_synth_scan.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_double, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p,
                                      ctypes.c_void_p]
        _lib = lib
    return _lib


def trajectory(seq, n_poses, y_start=-100.0, speed=0.8):
    """(truth, odom): two (n_poses,4,4) float64 arrays of T_world_sensor."""
    lib = _load()
    truth = np.empty((n_poses, 16), np.float64)
    odom = np.empty((n_poses, 16), np.float64)
    lib.ls_synth_trajectory(seq, n_poses, y_start, speed, truth.ctypes.data, odom.ctypes.data)
    # stored column-major
    return truth.reshape(n_poses, 4, 4).transpose(0, 2, 1).copy(), odom.reshape(n_poses, 4, 4).transpose(0, 2, 1).copy()


def scan(T_ws, seq=0, scan_idx=0, sensor=HDL64, scene_seed=1, scene_length=300.0, range_sigma=0.02):
    """One scan from pose T_ws (4x4 float64).  Returns (features (N,4) float32, normals (N,3) float32);
    features.ravel() is exactly DataPoints::features.data() (column-major 4xN)."""
    lib = _load()
    n = lib.ls_synth_num_rays(sensor)
    feats = np.empty((n, 4), np.float32)
    nrm = np.empty((n, 3), np.float32)
    T = np.ascontiguousarray(np.asarray(T_ws, np.float64).T)  # column-major
    got = lib.ls_synth_scan(sensor, scene_seed, scene_length, seq, scan_idx, T.ctypes.data, range_sigma,
                            feats.ctypes.data, nrm.ctypes.data)
    assert got == n
    return feats, nrm


def subsample(feats, nrm, step):
    """Every `step`-th azimuth of every ring (keeps the ring-major order); for small test clouds."""
    n_az = 2048
    rings = feats.shape[0] // n_az
    idx = (np.arange(rings)[:, None] * n_az + np.arange(0, n_az, step)[None, :]).ravel()
    return np.ascontiguousarray(feats[idx]), np.ascontiguousarray(nrm[idx])
