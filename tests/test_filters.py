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
