"""Input / map-maintenance side: oracle sanity on the CPU, device vs oracle on the GPU."""
import numpy as np
import pytest


def test_oracle_cylinder_and_voxel_grid_on_known_answers(oracle_mod):
    o = oracle_mod
    pts = np.array([[0, 0, 0, 1], [3, 4, 0, 1], [3, 4, 20, 1], [3, 4.001, 0, 1], [10, 0, 0, 1]], np.float32)
    inside = o.filter_cylinder(pts, [0, 0, 0], 5.0, 40.0, False)
    outside = o.filter_cylinder(pts, [0, 0, 0], 5.0, 40.0, True)
    assert [tuple(p[:3]) for p in inside] == [(0, 0, 0), (3, 4, 0), (3, 4, 20)]        # on the radius / half height: kept (<=)
    assert np.array_equal(outside, pts[1:])                                            # >= keeps the boundary points too
    cloud = np.array([[0.1, 0.1, 0.1, 1], [0.3, 0.1, 0.1, 1], [0.9, 0.9, 0.9, 1], [-0.2, 0.1, 0.1, 1], [1.2, 0.2, 0.2, 1]], np.float32)
    v = o.voxel_grid(cloud, 1.0)
    assert len(v) == 3                                                                # cells x = -1, 0, 1
    assert np.allclose(v[0, :3], [-0.2, 0.1, 0.1]) and np.allclose(v[2, :3], [1.2, 0.2, 0.2])
    assert np.allclose(v[1, :3], np.float32([0.1 + 0.3 + 0.9, 0.1 + 0.1 + 0.9, 0.1 + 0.1 + 0.9]) / 3, atol=1e-7)
    assert len(o.voxel_grid(np.zeros((0, 4), np.float32), 0.5)) == 0


@pytest.mark.gpu
def test_gpu_input_side_matches_oracle(oracle_mod, scans):
    import laser_slam_b200 as ls
    o = oracle_mod
    pts = scans[0][0]
    # PointCloud2-shaped payload: 32-byte records {x, y, z, pad, intensity, ring, pad, pad}
    rec = np.zeros((len(pts), 8), np.float32)
    rec[:, 0:3] = pts[:, :3]
    rec[:, 4] = 17.0
    got = ls.ingest_pointcloud2(rec.tobytes(), 32, 0, 4, 8, len(pts))
    assert np.array_equal(got, pts)
    center = [1.5, -2.0, 0.3]
    for inside_removed in (False, True):
        a = ls.filter_cylinder(pts, center, 18.0, 6.0, inside_removed)
        b = o.filter_cylinder(pts, center, 18.0, 6.0, inside_removed)
        assert np.array_equal(a, b) and 0 < len(a) < len(pts)
    for leaf in (0.5, (0.25, 0.5, 1.0), 3.0):
        a = ls.voxel_grid(pts, leaf)
        b = o.voxel_grid(pts, leaf)
        assert a.shape == b.shape and np.array_equal(a, b)
    both = np.concatenate([pts[:5000], pts[:5000]])                                   # duplicates: same cells, same centroids
    assert np.array_equal(ls.voxel_grid(both, 0.5), ls.voxel_grid(pts[:5000], 0.5))
    assert len(ls.voxel_grid(np.zeros((0, 4), np.float32), 0.5)) == 0 and len(ls.filter_cylinder(np.zeros((0, 4), np.float32), center, 1, 1)) == 0


# ---- Velodyne assembler / de-skew (reference sensor_drivers/velodyne_assembler/src/velodyne_assembler_ros.cpp:57-143)
def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    T = np.eye(4)
    T[:2, :2] = [[c, -s], [s, c]]
    return T


def _drive(n_rev=2.5, packets_per_rev=36, pts_per_packet=40, seed=3, v=(8.0, 0.5, 0.0), yaw_rate=0.6, static=False):
    """A spinning sensor on a moving vehicle: returns the packets [(points in the sensor frame at the packet's time,
    T_fixed_base (float32), stamp, world points)], T_sensor_base and the pose T_fixed_sensor(t) as a function."""
    rng = np.random.default_rng(seed)
    T_base_sensor = np.eye(4)
    T_base_sensor[:3, 3] = [1.2, 0.0, 1.8]
    T_base_sensor = T_base_sensor @ _rot_z(0.05)
    T_sensor_base = np.linalg.inv(T_base_sensor)
    dt = 0.1 / packets_per_rev                       # 10 Hz revolutions

    def T_fixed_base(t):
        if static:
            return np.eye(4)
        T = _rot_z(yaw_rate * t)
        T[:3, 3] = np.asarray(v) * t
        return T

    packets = []
    for k in range(int(n_rev * packets_per_rev)):
        t = k * dt
        az0 = np.pi / 2 - 2 * np.pi * (k % packets_per_rev) / packets_per_rev      # the sensor spins clockwise
        az = az0 - rng.uniform(0, 2 * np.pi / packets_per_rev, pts_per_packet)
        az[0] = az0                                                               # the packet's first point decides the wrap
        r = rng.uniform(4.0, 40.0, pts_per_packet)
        z = rng.uniform(-2.0, 3.0, pts_per_packet)
        p = np.stack([r * np.cos(az), r * np.sin(az), z, np.ones_like(r)], 1)
        Tfb = T_fixed_base(t)
        world = (Tfb @ T_base_sensor @ p.T).T
        packets.append((p.astype(np.float32), Tfb.astype(np.float32), int(t * 1e9), world))
    return packets, T_sensor_base.astype(np.float32), lambda t: T_fixed_base(t) @ T_base_sensor


def test_oracle_assembler_deskews_a_moving_sensor(oracle_mod):
    """The restated assembler puts every point of a revolution where the static world is, seen from the sensor at the
    revolution's LAST packet; revolutions end when the first point's azimuth wraps past +pi/2."""
    o = oracle_mod
    packets, T_sensor_base, T_fixed_sensor = _drive()
    asm = o.VelodyneAssembler(T_sensor_base)
    revs, start = [], 0
    for k, (p, Tfb, stamp, _) in enumerate(packets):
        out = asm.add_packet(p, Tfb, stamp)
        if out is not None:
            revs.append((out, start, k))
            start = k
    assert [(a, b) for _, a, b in revs] == [(0, 36), (36, 72)]                  # 2.5 revolutions fed, 2 complete
    for (cloud, stamp), a, b in revs:
        assert stamp == packets[b - 1][2]
        world = np.concatenate([packets[k][3] for k in range(a, b)])
        want = (np.linalg.inv(T_fixed_sensor(packets[b - 1][2] * 1e-9)) @ world.T).T
        assert cloud.shape == want.shape
        assert np.abs(cloud[:, :3] - want[:, :3]).max() < 2e-3                   # float32 chain of ~36 composed transforms
        raw = np.concatenate([packets[k][0] for k in range(a, b)])
        assert np.abs(raw[:, :3] - want[:, :3]).max() > 0.3                      # ... and the motion was worth undoing
    # naive assembling ignores the motion: the packets are concatenated (through T_sensor_base * I * T_base_sensor, which
    # is the identity only up to float32 rounding -- as in the reference)
    naive = o.VelodyneAssembler(T_sensor_base, naive=True)
    outs = [naive.add_packet(p, Tfb, st) for p, Tfb, st, _ in packets]
    first = [x for x in outs if x is not None][0][0]
    assert np.allclose(first, np.concatenate([packets[k][0] for k in range(36)]), atol=1e-4)
    # empty packets are ignored (reference :78)
    assert o.VelodyneAssembler().add_packet(np.zeros((0, 4), np.float32), np.eye(4), 0) is None


def test_oracle_deskew_revolution_is_two_float32_transforms(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-30, 30, (50, 3)), np.ones((50, 1))], 1).astype(np.float32)
    offs = [0, 10, 10, 35, 50]                                                     # packet 1 is empty
    Ts = [np.eye(4, dtype=np.float32)] + [(_rot_z(0.01 * k) + np.pad(np.zeros((3, 3)), ((0, 1), (0, 1)))).astype(np.float32) for k in (1, 2, 3)]
    for k in (1, 2, 3):
        Ts[k][:3, 3] = [0.1 * k, -0.05 * k, 0.0]
    Tf = o.rigid_inverse_f32(Ts[3])
    got = o.deskew_revolution(pts, offs, Ts, Tf)
    assert np.array_equal(got[:10], o.transform_cloud(Tf, pts[:10], np.zeros((10, 3), np.float32))[0])   # identity packet: only T_final
    mid = o.transform_cloud(Ts[2], pts[10:35], np.zeros((25, 3), np.float32))[0]
    assert np.array_equal(got[10:35], o.transform_cloud(Tf, mid, np.zeros((25, 3), np.float32))[0])
    assert np.array_equal(o.deskew_revolution(pts, offs, [np.eye(4)] * 4, np.eye(4)), pts)               # identities copy verbatim
    assert np.allclose(o.matmul_f32(Tf, Ts[3]), np.eye(4), atol=1e-6)


@pytest.mark.gpu
def test_gpu_deskew_and_assembler_match_oracle(oracle_mod):
    import laser_slam_b200 as ls
    from laser_slam_b200 import host
    o = oracle_mod
    rng = np.random.default_rng(11)
    # the kernel alone: ragged packets, empty ones, identity and non-identity transforms
    sizes = [0, 7, 1, 0, 300, 64, 0, 1000, 33]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    pts = np.concatenate([rng.uniform(-50, 50, (offs[-1], 3)), np.ones((offs[-1], 1))], 1).astype(np.float32)
    Ts = []
    for k in range(len(sizes)):
        T = _rot_z(0.02 * k).astype(np.float32)
        T[:3, 3] = [0.3 * k, -0.1 * k, 0.01 * k]
        Ts.append(np.eye(4, dtype=np.float32) if k in (1, 5) else T)
    for Tf in (o.rigid_inverse_f32(Ts[-1]), np.eye(4, dtype=np.float32)):
        assert np.array_equal(ls.deskew_revolution(pts, offs, Ts, Tf), o.deskew_revolution(pts, offs, Ts, Tf))
    # the C++ assembler (host bookkeeping + one kernel per revolution) against the restated reference callback
    for static in (False, True):
        packets, T_sensor_base, _ = _drive(n_rev=3.2, packets_per_rev=24, pts_per_packet=97, seed=21, static=static)
        a, b = host.Assembler(T_sensor_base), o.VelodyneAssembler(T_sensor_base)
        n_rev = 0
        for p, Tfb, stamp, _ in packets:
            x, y = a.add_packet(p, Tfb, stamp), b.add_packet(p, Tfb, stamp)
            assert (x is None) == (y is None)
            if x is not None:
                n_rev += 1
                assert x[1] == y[1] and x[0].shape == y[0].shape and np.array_equal(x[0], y[0])
        assert n_rev == 3
        a.close()


def test_assembler_float32_helpers_equal_the_oracle_bits(oracle_mod, tmp_path):
    """VelodyneAssembler::multiply / rigidInverse (C++, host) and oracle.matmul_f32 / rigid_inverse_f32 round identically."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    build = os.path.join(root, "laser_slam_b200", "_build")
    exe = str(tmp_path / "assembler_math")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "compile", "assembler_math.cpp"), "-o", exe, "-L", build, "-lls_b200",
                        f"-Wl,-rpath,{build}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(2)
    for trial in range(5):
        A = _rot_z(rng.uniform(-3, 3)) @ np.array([[1, 0, 0, 0], [0, np.cos(0.3), -np.sin(0.3), 0], [0, np.sin(0.3), np.cos(0.3), 0], [0, 0, 0, 1.0]])
        A[:3, 3] = rng.uniform(-100, 100, 3)
        B = _rot_z(rng.uniform(-3, 3))
        B[:3, 3] = rng.uniform(-5, 5, 3)
        A, B = A.astype(np.float32), B.astype(np.float32)
        args = [f"{int(x):08x}" for x in np.concatenate([A.T.reshape(16), B.T.reshape(16)]).view(np.uint32)]   # column-major
        out = subprocess.run([exe] + args, capture_output=True, text=True)
        assert out.returncode == 0
        rows = [np.array([int(t, 16) for t in line.split()], np.uint32).view(np.float32).reshape(4, 4).T for line in out.stdout.splitlines()]
        assert np.array_equal(rows[0], oracle_mod.matmul_f32(A, B))
        assert np.array_equal(rows[1], oracle_mod.rigid_inverse_f32(A))
