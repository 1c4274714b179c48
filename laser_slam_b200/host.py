"""ctypes access to the C test hooks of the C++ host layer (laser_slam_b200/host/host_capi.cpp): drives
laser_slam::IncrementalEstimator / LaserTrack the way the ROS worker's scanCallback does
(reference laser_slam_ros/src/laser_slam_worker.cpp:124-173)."""
import ctypes
import os

import numpy as np

from . import IcpStats, LsError, ConvergenceError, LS_ERR_CONVERGENCE, build as _build_all

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libls_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            _build_all()
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.lsh_create.restype = vp
        L.lsh_create.argtypes = [ci, ci, ci, ci, ci, ci, ci, ci, ctypes.c_char_p, ctypes.c_char_p, ci]
        L.lsh_destroy.argtypes = [vp]
        L.lsh_destroy.restype = None
        L.lsh_last_error.argtypes = [vp]
        L.lsh_last_error.restype = ctypes.c_char_p
        L.lsh_step.argtypes = [vp, ci, i64, vp, vp, vp, ci, vp, ctypes.POINTER(IcpStats)]
        L.lsh_loop_closure.argtypes = [vp, ci, i64, ci, i64, vp]
        L.lsh_step_batch.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, ci, vp, vp]
        L.lsh_begin_batch.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, ci]
        L.lsh_prefetch.argtypes = [vp, ci, vp, vp, vp, vp, vp, ci]
        L.lsh_end_batch.argtypes = [vp, ci, vp, vp]
        L.lsh_trajectory.argtypes = [vp, ci, vp, vp, ci]
        L.lsh_num_scans.argtypes = [vp, ci]
        L.lsh_build_submap.argtypes = [vp, ci, i64, ci, vp, vp, ci]
        _lib = L
    return _lib


class Estimator:
    """laser_slam::IncrementalEstimator with n_workers LaserTracks."""

    def __init__(self, n_workers=1, nscan_in_sub_map=4, use_icp_factors=True, use_odom_factors=True, robust_icp=True,
                 device=0, do_icp_step_on_loop_closures=False, loop_closures_sub_maps_radius=2, icp_yaml_path=None):
        err = ctypes.create_string_buffer(512)
        self._h = lib().lsh_create(n_workers, nscan_in_sub_map, int(use_icp_factors), int(use_odom_factors), int(robust_icp),
                                   device, int(do_icp_step_on_loop_closures), loop_closures_sub_maps_radius,
                                   icp_yaml_path.encode() if icp_yaml_path else None, err, 512)
        if not self._h:
            raise LsError(err.value.decode() or "lsh_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().lsh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == LS_ERR_CONVERGENCE:
            raise ConvergenceError(lib().lsh_last_error(self._h).decode())
        if rc < 0:
            raise LsError(lib().lsh_last_error(self._h).decode())
        return rc

    def step(self, worker, time_ns, pose7, features4, normals3):
        """One scan callback; returns (icp T_a_b as 7 doubles, IcpStats)."""
        f = np.ascontiguousarray(features4, np.float32)
        nr = np.ascontiguousarray(normals3, np.float32)
        p = np.ascontiguousarray(pose7, np.float64)
        out = np.zeros(7, np.float64)
        st = IcpStats()
        self._check(lib().lsh_step(self._h, worker, int(time_ns), p.ctypes.data, f.ctypes.data, nr.ctypes.data, f.shape[0],
                                   out.ctypes.data, ctypes.byref(st)))
        return out, st

    def step_batch(self, workers, times_ns, poses7, feat_ptrs, nrm_ptrs, ns, with_estimator=True, views=False):
        """The scan callbacks of several workers at once: one batched launch registers them all
        (IncrementalEstimator::processPosesAndLaserScans).  feat_ptrs / nrm_ptrs: host addresses of each worker's
        4xN features and 3xN normals.  views=True hands the tracks DataPoints that BORROW those arrays (no copy; they
        must stay valid while the estimator lives; pinned arrays are then uploaded asynchronously, in place).
        Returns (icp T_a_b (len,7), list of IcpStats)."""
        k = len(workers)
        w = np.ascontiguousarray(workers, np.int32)
        t = np.ascontiguousarray(times_ns, np.int64)
        p = np.ascontiguousarray(poses7, np.float64).reshape(k, 7)
        fp = (ctypes.c_void_p * k)(*[int(a) for a in feat_ptrs])
        npp = (ctypes.c_void_p * k)(*[int(a) for a in nrm_ptrs])
        nn = np.ascontiguousarray(ns, np.int32)
        out = np.zeros((k, 7), np.float64)
        st = (IcpStats * k)()
        self._check(lib().lsh_step_batch(self._h, k, w.ctypes.data, t.ctypes.data, p.ctypes.data, ctypes.cast(fp, ctypes.c_void_p),
                                         ctypes.cast(npp, ctypes.c_void_p), nn.ctypes.data, int(bool(with_estimator)) | (2 if views else 0), out.ctypes.data,
                                         ctypes.cast(st, ctypes.c_void_p)))
        return out, list(st)

    def _ptr_arrays(self, workers, times_ns, feat_ptrs, nrm_ptrs, ns):
        k = len(workers)
        return (k, np.ascontiguousarray(workers, np.int32), np.ascontiguousarray(times_ns, np.int64),
                (ctypes.c_void_p * k)(*[int(a) for a in feat_ptrs]), (ctypes.c_void_p * k)(*[int(a) for a in nrm_ptrs]),
                np.ascontiguousarray(ns, np.int32))

    def begin_batch(self, workers, times_ns, poses7, feat_ptrs, nrm_ptrs, ns, views=False):
        """First half of step_batch (IncrementalEstimator::beginPosesAndLaserScans): stage and launch, return at once."""
        k, w, t, fp, npp, nn = self._ptr_arrays(workers, times_ns, feat_ptrs, nrm_ptrs, ns)
        p = np.ascontiguousarray(poses7, np.float64).reshape(k, 7)
        self._pending_k = k
        self._check(lib().lsh_begin_batch(self._h, k, w.ctypes.data, t.ctypes.data, p.ctypes.data, ctypes.cast(fp, ctypes.c_void_p),
                                          ctypes.cast(npp, ctypes.c_void_p), nn.ctypes.data, 2 if views else 0))

    def prefetch(self, workers, times_ns, feat_ptrs, nrm_ptrs, ns, views=False):
        """Hint (IncrementalEstimator::prefetchLaserScans): upload scans a later begin_batch will be handed."""
        k, w, t, fp, npp, nn = self._ptr_arrays(workers, times_ns, feat_ptrs, nrm_ptrs, ns)
        self._check(lib().lsh_prefetch(self._h, k, w.ctypes.data, t.ctypes.data, ctypes.cast(fp, ctypes.c_void_p),
                                       ctypes.cast(npp, ctypes.c_void_p), nn.ctypes.data, 2 if views else 0))

    def end_batch(self, with_estimator=True):
        """Second half of step_batch.  Returns (icp T_a_b (len,7), list of IcpStats)."""
        k = self._pending_k
        out = np.zeros((k, 7), np.float64)
        st = (IcpStats * k)()
        self._check(lib().lsh_end_batch(self._h, int(bool(with_estimator)), out.ctypes.data, ctypes.cast(st, ctypes.c_void_p)))
        return out, list(st)

    def loop_closure(self, track_a, time_a, track_b, time_b, w_T_a_b7):
        p = np.ascontiguousarray(w_T_a_b7, np.float64)
        self._check(lib().lsh_loop_closure(self._h, track_a, int(time_a), track_b, int(time_b), p.ctypes.data))

    def trajectory(self, worker=0):
        n = self._check(lib().lsh_trajectory(self._h, worker, None, None, 0))
        times = np.zeros(max(n, 1), np.int64)
        poses = np.zeros((max(n, 1), 7), np.float64)
        self._check(lib().lsh_trajectory(self._h, worker, times.ctypes.data, poses.ctypes.data, n))
        return times[:n], poses[:n]

    def num_scans(self, worker=0):
        return self._check(lib().lsh_num_scans(self._h, worker))

    def build_submap(self, worker, time_ns, radius, cap_points):
        f = np.zeros((cap_points, 4), np.float32)
        nr = np.zeros((cap_points, 3), np.float32)
        m = self._check(lib().lsh_build_submap(self._h, worker, int(time_ns), radius, f.ctypes.data, nr.ctypes.data, cap_points))
        return f[:m], nr[:m]


class Assembler:
    """laser_slam::VelodyneAssembler (include/laser_slam/velodyne_assembler.hpp): packets in, de-skewed revolutions out."""

    def __init__(self, T_sensor_base=None, naive=False, device=0):
        L = lib()
        L.lsh_assembler_create.restype = ctypes.c_void_p
        L.lsh_assembler_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.lsh_assembler_destroy.argtypes = [ctypes.c_void_p]
        L.lsh_assembler_destroy.restype = None
        L.lsh_assembler_add_packet.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                               ctypes.POINTER(ctypes.c_int64)]
        t = None if T_sensor_base is None else np.ascontiguousarray(np.asarray(T_sensor_base, np.float32).T).reshape(16)
        self._h = ctypes.c_void_p(L.lsh_assembler_create(None if t is None else t.ctypes.data, int(naive), device))
        self._cap = 1 << 20
        self._out = np.empty((self._cap, 4), np.float32)

    def add_packet(self, points4, T_fixed_base, stamp_ns):
        """None, or (revolution points (m,4), stamp of its last packet)."""
        p = np.ascontiguousarray(points4, np.float32)
        t = np.ascontiguousarray(np.asarray(T_fixed_base, np.float32).T).reshape(16)   # column-major
        m, st = ctypes.c_int(0), ctypes.c_int64(0)
        rc = lib().lsh_assembler_add_packet(self._h, p.ctypes.data, len(p), t.ctypes.data, int(stamp_ns), self._out.ctypes.data,
                                            self._cap, ctypes.byref(m), ctypes.byref(st))
        if rc < 0:
            raise RuntimeError(f"lsh_assembler_add_packet failed with {rc}")
        return (self._out[:m.value].copy(), int(st.value)) if rc == 1 else None

    def close(self):
        if self._h:
            lib().lsh_assembler_destroy(self._h)
            self._h = None
