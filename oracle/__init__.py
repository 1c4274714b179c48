"""ORACLE — test infrastructure only.  ctypes bindings of oracle/icp_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (laser_slam_b200/) never does.  PARITY UNPINNED — see icp_oracle.cpp.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libls_oracle.so")
_lib = None


class IcpParams(ctypes.Structure):
    _fields_ = [("max_iterations", ctypes.c_int), ("trim_ratio", ctypes.c_float),
                ("use_differential", ctypes.c_int), ("min_diff_rot", ctypes.c_float),
                ("min_diff_trans", ctypes.c_float), ("smooth_length", ctypes.c_int),
                ("num_threads", ctypes.c_int)]


class IcpStats(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int), ("converged", ctypes.c_int), ("max_iter_reached", ctypes.c_int),
                ("last_kept", ctypes.c_int), ("last_limit", ctypes.c_float), ("used_ratio", ctypes.c_float)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("icp_oracle.cpp", "icp_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.lso_default_params.argtypes = [ctypes.POINTER(IcpParams)]
        L.lso_mean.argtypes = [vp, ci, vp]
        L.lso_nn_brute.argtypes = [vp, ci, vp, ci, vp, vp]
        L.lso_nn_kdtree.argtypes = [vp, ci, vp, ci, vp, vp, ci]
        L.lso_trim_limit.argtypes = [vp, ci, cf, vp]
        L.lso_trim_limit.restype = cf
        L.lso_transform_points.argtypes = [vp, vp, ci, vp]
        L.lso_transform_cloud.argtypes = [vp, vp, vp, ci, ci, vp, vp]
        L.lso_check_rigid.argtypes = [vp]
        L.lso_check_rigid.restype = ci
        L.lso_correct_rigid.argtypes = [vp, vp]
        L.lso_normal_equations.argtypes = [vp, ci, vp, vp, ci, vp, vp, cf, vp, vp, vp, vp, vp]
        L.lso_solve_step.argtypes = [vp, vp, vp, vp]
        L.lso_solve_step.restype = ci
        L.lso_mat4_mul.argtypes = [vp, vp, vp]
        L.lso_sincos.argtypes = [ctypes.c_double, vp, vp]
        L.lso_knn_self.argtypes = [vp, ci, ci, vp, vp, ci]
        L.lso_knn_normals.argtypes = [vp, ci, ci, vp, ci]
        L.lso_icp.argtypes = [vp, ci, vp, vp, ci, ci, vp, ctypes.POINTER(IcpParams), vp,
                              ctypes.POINTER(IcpStats), vp, vp, vp]
        L.lso_icp.restype = ci
        _lib = L
    return _lib


def default_params(**kw):
    p = IcpParams()
    lib().lso_default_params(ctypes.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def colmajor(T):
    """4x4 (row, col) array -> 16 floats column-major (Eigen / TransformationParameters::data())."""
    return np.ascontiguousarray(np.asarray(T, np.float32).T).ravel()


def from_colmajor(t16):
    return np.asarray(t16).reshape(4, 4).T.copy()


def mean(ref4):
    ref4 = _f32(ref4)
    mu = np.empty(3, np.float32)
    lib().lso_mean(ref4.ctypes.data, ref4.shape[0], mu.ctypes.data)
    return mu


def nn_brute(q3, ref3):
    q3, ref3 = _f32(q3), _f32(ref3)
    ids = np.empty(q3.shape[0], np.int32)
    d2 = np.empty(q3.shape[0], np.float32)
    lib().lso_nn_brute(q3.ctypes.data, q3.shape[0], ref3.ctypes.data, ref3.shape[0], ids.ctypes.data, d2.ctypes.data)
    return ids, d2


def nn_kdtree(q3, ref3, num_threads=1):
    q3, ref3 = _f32(q3), _f32(ref3)
    ids = np.empty(q3.shape[0], np.int32)
    d2 = np.empty(q3.shape[0], np.float32)
    lib().lso_nn_kdtree(q3.ctypes.data, q3.shape[0], ref3.ctypes.data, ref3.shape[0], ids.ctypes.data,
                        d2.ctypes.data, num_threads)
    return ids, d2


def trim_limit(d2, ratio):
    d2 = _f32(d2)
    nf = ctypes.c_int(0)
    lim = lib().lso_trim_limit(d2.ctypes.data, d2.shape[0], ratio, ctypes.addressof(nf))
    return float(lim), nf.value


def transform_points(T, pts4):
    pts4 = _f32(pts4)
    out = np.empty_like(pts4)
    t = colmajor(T)
    lib().lso_transform_points(t.ctypes.data, pts4.ctypes.data, pts4.shape[0], out.ctypes.data)
    return out


def transform_cloud(T, pts4, nrm3):
    pts4, nrm3 = _f32(pts4), _f32(nrm3)
    out = np.empty_like(pts4)
    nout = np.empty_like(nrm3)
    t = colmajor(T)
    lib().lso_transform_cloud(t.ctypes.data, pts4.ctypes.data, nrm3.ctypes.data, 3, pts4.shape[0],
                              out.ctypes.data, nout.ctypes.data)
    return out, nout


def check_rigid(T):
    t = colmajor(T)
    return bool(lib().lso_check_rigid(t.ctypes.data))


def correct_rigid(T):
    t = colmajor(T)
    o = np.empty(16, np.float32)
    lib().lso_correct_rigid(t.ctypes.data, o.ctypes.data)
    return from_colmajor(o)


def normal_equations(step4, refc3, nrm3, ids, d2, limit):
    step4, refc3, nrm3, d2 = _f32(step4), _f32(refc3), _f32(nrm3), _f32(d2)
    ids = np.ascontiguousarray(ids, np.int32)
    A = np.empty((6, 6))
    b = np.empty(6)
    Ad = np.empty((6, 6))
    bd = np.empty(6)
    kept = ctypes.c_int(0)
    lib().lso_normal_equations(step4.ctypes.data, step4.shape[0], refc3.ctypes.data, nrm3.ctypes.data, 3,
                               ids.ctypes.data, d2.ctypes.data, limit, A.ctypes.data, b.ctypes.data,
                               ctypes.addressof(kept), Ad.ctypes.data, bd.ctypes.data)
    return A, b, kept.value, Ad, bd


def solve_step(A, b):
    A = np.ascontiguousarray(A, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    T = np.empty(16, np.float32)
    x = np.empty(6)
    rc = lib().lso_solve_step(A.ctypes.data, b.ctypes.data, T.ctypes.data, x.ctypes.data)
    return rc, from_colmajor(T), x


def mat4_mul(A, B):
    a, b = colmajor(A), colmajor(B)
    c = np.empty(16, np.float32)
    lib().lso_mat4_mul(a.ctypes.data, b.ctypes.data, c.ctypes.data)
    return from_colmajor(c)


def knn_self(pts4, k, num_threads=1):
    pts4 = _f32(pts4)
    n = pts4.shape[0]
    ids = np.empty((n, k), np.int32)
    d2 = np.empty((n, k), np.float32)
    lib().lso_knn_self(pts4.ctypes.data, n, k, ids.ctypes.data, d2.ctypes.data, num_threads)
    return ids, d2


def knn_normals(pts4, k=10, num_threads=1):
    pts4 = _f32(pts4)
    out = np.empty((pts4.shape[0], 3), np.float32)
    lib().lso_knn_normals(pts4.ctypes.data, pts4.shape[0], k, out.ctypes.data, num_threads)
    return out


def keep_mask(n, salt, prob):
    """Deterministic stand-in for RandomSamplingDataPointsFilter (reference laser_slam/configurations/icp_default.yaml:1-3;
    libpointmatcher draws `rand() / RAND_MAX < prob`, which no two processes reproduce): point i is kept iff
    hash32(i, salt) < prob * 2^32.  [DEFINED] -- the same counter-based rule as ls_keep_point (include/ls_b200.h)."""
    if not prob < 1.0:
        return np.ones(n, bool)
    if not prob > 0.0:
        return np.zeros(n, bool)
    i = np.arange(n, dtype=np.uint64)
    M = np.uint64(0xFFFFFFFF)
    h = (i * np.uint64(0x9E3779B1) + np.uint64(salt) * np.uint64(0x85EBCA77) + np.uint64(0x165667B1)) & M
    h ^= h >> np.uint64(15); h = (h * np.uint64(0x2C1B3C6D)) & M
    h ^= h >> np.uint64(12); h = (h * np.uint64(0x297A2D39)) & M
    h ^= h >> np.uint64(15)
    return h.astype(np.float64) < float(np.float32(prob)) * 4294967296.0


def filter_cylinder(pts4, center, radius_m, height_m, remove_points_inside=False):
    """applyCylindricalFilter (reference laser_slam_ros/include/laser_slam_ros/common.hpp:194-223): double arithmetic on
    float coordinates, <= / >= exactly as written there, input order kept."""
    p = np.asarray(pts4, np.float32)
    c = np.asarray(center, np.float64)
    d2 = (p[:, 0].astype(np.float64) - c[0]) ** 2 + (p[:, 1].astype(np.float64) - c[1]) ** 2
    dz = np.abs(p[:, 2].astype(np.float64) - c[2])
    r2, hh = float(radius_m) ** 2, float(height_m) / 2.0
    keep = ((d2 >= r2) | (dz >= hh)) if remove_points_inside else ((d2 <= r2) & (dz <= hh))
    return p[keep].copy()


def voxel_grid(pts4, leaf_size):
    """pcl::VoxelGrid as LaserSlamWorker::getFilteredMap runs it (reference laser_slam_ros/src/laser_slam_worker.cpp:434-441):
    cell = floor(p * (1 / leaf)) in float32, one output point per occupied cell, cells in ascending linear index (x fastest);
    [DEFINED] the centroid is the exact mean (fixed point 2^-24) rounded once -- PCL's float running sum depends on the order."""
    p = np.asarray(pts4, np.float32)
    leaf = np.broadcast_to(np.asarray(leaf_size, np.float32), (3,))
    inv = (np.float32(1.0) / leaf).astype(np.float32)
    ok = np.isfinite(p[:, :3]).all(1)
    q = p[ok]
    if len(q) == 0:
        return np.zeros((0, 4), np.float32)
    ijk = np.floor(q[:, :3] * inv[None, :]).astype(np.int64)
    mn = ijk.min(0)
    dim = ijk.max(0) - mn + 1
    key = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * dim[0] + (ijk[:, 2] - mn[2]) * dim[0] * dim[1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    heads = np.flatnonzero(np.concatenate([[True], ks[1:] != ks[:-1]]))
    fx = np.rint(q[order, :3].astype(np.float64) * 16777216.0).astype(np.int64)
    sums = np.add.reduceat(fx, heads, axis=0)
    cnt = np.diff(np.concatenate([heads, [len(ks)]]))
    out = np.ones((len(heads), 4), np.float32)
    out[:, :3] = (sums.astype(np.float64) / (cnt[:, None].astype(np.float64) * 16777216.0)).astype(np.float32)
    return out


def sincos(x):
    s = ctypes.c_double()
    c = ctypes.c_double()
    lib().lso_sincos(x, ctypes.addressof(s), ctypes.addressof(c))
    return s.value, c.value


def icp(reading4, ref4, ref_normals3, T0, params=None, want_hist=False):
    """Returns dict(rc, T (4x4 f32), stats, ids_hist (iters,n) or None, d2_last, T_iter_hist)."""
    reading4, ref4, ref_normals3 = _f32(reading4), _f32(ref4), _f32(ref_normals3)
    n, m = reading4.shape[0], ref4.shape[0]
    p = params or default_params()
    t0 = colmajor(T0)
    tout = np.empty(16, np.float32)
    st = IcpStats()
    ids_hist = np.full((p.max_iterations, max(n, 1)), -2, np.int32) if want_hist else None
    t_hist = np.zeros((p.max_iterations, 16), np.float32) if want_hist else None
    d2_last = np.empty(max(n, 1), np.float32)
    rc = lib().lso_icp(reading4.ctypes.data, n, ref4.ctypes.data, ref_normals3.ctypes.data, 3, m, t0.ctypes.data,
                       ctypes.byref(p), tout.ctypes.data, ctypes.byref(st),
                       ids_hist.ctypes.data if want_hist else None, d2_last.ctypes.data,
                       t_hist.ctypes.data if want_hist else None)
    return dict(rc=rc, T=from_colmajor(tout), stats=st,
                ids_hist=ids_hist[:st.iterations, :n] if want_hist else None, d2_last=d2_last[:n],
                T_iter_hist=(t_hist[:st.iterations].reshape(-1, 4, 4).transpose(0, 2, 1).copy()
                             if want_hist else None))


# ---------------------------------------------------------------------------------------------------------------------
# Velodyne assembler (reference sensor_drivers/velodyne_assembler/src/velodyne_assembler_ros.cpp)
def _is_identity(T):
    return bool(np.array_equal(np.asarray(T, np.float32), np.eye(4, dtype=np.float32)))


def _xform_points(T, pts4):
    """float32 point transform with the operation order of lso_transform_cloud; an exact identity copies verbatim."""
    p = np.ascontiguousarray(pts4, np.float32)
    if len(p) == 0 or _is_identity(T):
        return p.copy()
    out, _ = transform_cloud(T, p, np.zeros((len(p), 3), np.float32))
    return out


def deskew_revolution(points4, packet_offsets, T_packets, T_final):
    """out = T_final (x) (T_packets[k] (x) p) for the points of packet k: the two float32 transforms the reference applies
    to a packet (velodyne_assembler_ros.cpp:129-133 on arrival, :107-108 before publishing)."""
    p = np.ascontiguousarray(points4, np.float32)
    out = np.empty_like(p)
    for k in range(len(packet_offsets) - 1):
        a, b = int(packet_offsets[k]), int(packet_offsets[k + 1])
        out[a:b] = _xform_points(T_packets[k], p[a:b])
    return _xform_points(T_final, out)


def rigid_inverse_f32(T):
    """[DEFINED] inverse of a rigid 4x4 in float32: [R^T, -(R^T t)], every product and sum rounded to float32 in a fixed
    order (the reference calls Eigen's general 4x4 inverse, velodyne_assembler_ros.cpp:95,107)."""
    T = np.asarray(T, np.float32)
    out = np.eye(4, dtype=np.float32)
    for i in range(3):
        for j in range(3):
            out[i, j] = T[j, i]
    for i in range(3):
        a = np.float32(T[0, i] * T[0, 3])
        b = np.float32(T[1, i] * T[1, 3])
        c = np.float32(T[2, i] * T[2, 3])
        s = np.float32(np.float32(a + b) + c)
        out[i, 3] = -s
    return out


def matmul_f32(A, B):
    """4x4 float32 product, c_ij = ((a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j) + a_i3 b_3j, each operation rounded."""
    A, B = np.asarray(A, np.float32), np.asarray(B, np.float32)
    C = np.empty((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            s = np.float32(np.float32(A[i, 0] * B[0, j]) + np.float32(A[i, 1] * B[1, j]))
            s = np.float32(s + np.float32(A[i, 2] * B[2, j]))
            C[i, j] = np.float32(s + np.float32(A[i, 3] * B[3, j]))
    return C


class VelodyneAssembler:
    """Restatement of VelodyneAssemblerRos::pclCallback (velodyne_assembler_ros.cpp:57-143) without ROS: packets in
    (points in the sensor frame + the vehicle pose T_fixed_base at the packet's stamp), one de-skewed revolution out
    whenever the azimuth of a packet's first point wraps past +pi/2 (:99-103)."""
    START_ANGLE = np.pi / 2.0

    def __init__(self, T_sensor_base=None, naive=False):
        self.T_sensor_base = np.eye(4, dtype=np.float32) if T_sensor_base is None else np.asarray(T_sensor_base, np.float32)
        self.T_base_sensor = rigid_inverse_f32(self.T_sensor_base)
        self.naive = naive
        self.T_fixed_base_prev = np.eye(4, dtype=np.float32)
        self.T_start_cur = np.eye(4, dtype=np.float32)
        self.initialized = False
        self.last_az = 0.0
        self.last_stamp = 0
        self.parts = []      # points of the revolution being assembled, already in the start frame

    def add_packet(self, points4, T_fixed_base, stamp):
        """Returns None, or (revolution points, stamp of its last packet) when this packet starts a new revolution."""
        p = np.ascontiguousarray(points4, np.float32)
        if len(p) == 0:
            return None
        T_cur = np.eye(4, dtype=np.float32) if self.naive else np.asarray(T_fixed_base, np.float32)
        T_prev_cur = matmul_f32(rigid_inverse_f32(self.T_fixed_base_prev), T_cur)
        self.T_fixed_base_prev = T_cur
        az = float(np.arctan2(np.float64(p[0, 1]), np.float64(p[0, 0])))
        out = None
        if (self.last_az > self.START_ANGLE and az <= self.START_ANGLE) or not self.initialized:
            if self.initialized:
                cloud = np.concatenate(self.parts)
                out = (_xform_points(rigid_inverse_f32(self.T_start_cur), cloud), self.last_stamp)
            self.parts = [p.copy()]
            self.initialized = True
            self.T_start_cur = np.eye(4, dtype=np.float32)
        else:
            T_sp_sc = matmul_f32(matmul_f32(self.T_sensor_base, T_prev_cur), self.T_base_sensor)
            self.T_start_cur = matmul_f32(self.T_start_cur, T_sp_sc)
            self.parts.append(_xform_points(self.T_start_cur, p))
        self.last_az = az
        self.last_stamp = stamp
        return out
