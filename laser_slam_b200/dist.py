"""Multi-GPU driver logic of the path (one process per GPU, launched by torch.distributed.run).

The path shards by independent tracks: track/sequence s runs on rank s % world.  Per step every rank
contributes one 32-byte pose record {delta[6], status, key}; on GPUs the exchange is ls_comm_allgather_pose_records
(one ncclAllGather over NVLink, C ABI); with the gloo backend (CPU tests of this host logic) the same records travel
through torch.distributed.all_gather.  Timing reductions (max over ranks) always use torch.distributed."""
import ctypes

import numpy as np

RECORD_DTYPE = np.dtype([("delta", np.float32, 6), ("status", np.int32), ("key", np.int32)])
assert RECORD_DTYPE.itemsize == 32


def shard(n_items, rank, world):
    """Items owned by `rank` (round-robin): disjoint across ranks, union = range(n_items)."""
    return list(range(rank, n_items, world))


def pose_record(T_a_b, status=0, key=0):
    """4x4 relative pose -> the 32-byte record (translation, rotation vector)."""
    T = np.asarray(T_a_b, np.float64)
    R = T[:3, :3]
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s, c = np.linalg.norm(v), 0.5 * (np.trace(R) - 1.0)
    th = np.arctan2(s, c)
    w = v * (th / s if s > 1e-12 else 1.0)
    rec = np.zeros((), RECORD_DTYPE)
    rec["delta"][:3] = T[:3, 3]
    rec["delta"][3:] = w
    rec["status"] = status
    rec["key"] = key
    return rec


class Exchange:
    """All-gather of pose records across ranks."""

    def __init__(self, rank, world, device=None):
        import torch.distributed as dist
        self.rank, self.world, self.device = rank, world, device
        self._comm = None
        if world > 1 and device is not None and dist.get_backend() == "nccl":
            import torch
            from . import lib
            L = lib()
            L.ls_comm_unique_id.argtypes = [ctypes.c_void_p]
            L.ls_comm_init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
            L.ls_comm_destroy.argtypes = [ctypes.c_void_p]
            L.ls_comm_destroy.restype = None
            L.ls_comm_allgather_pose_records.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            L.ls_comm_allgather_pose_records_begin.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            L.ls_comm_allgather_pose_records_end.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            uid = np.zeros(128, np.uint8)
            if rank == 0 and L.ls_comm_unique_id(uid.ctypes.data) != 0:
                raise RuntimeError("ls_comm_unique_id failed (NCCL not loadable)")
            t = torch.from_numpy(uid).cuda(device)
            dist.broadcast(t, 0)  # the NCCL id travels over the already-initialised process group
            uid = t.cpu().numpy()
            h = ctypes.c_void_p()
            rc = L.ls_comm_init(device, rank, world, uid.ctypes.data, ctypes.byref(h))
            if rc != 0:
                raise RuntimeError(f"ls_comm_init failed with {rc}")
            self._comm, self._L = h, L

    def allgather(self, rec):
        """rec: RECORD_DTYPE scalar.  Returns an array of `world` records in rank order."""
        out = np.zeros(self.world, RECORD_DTYPE)
        if self.world == 1:
            out[0] = rec
            return out
        mine = np.array(rec, RECORD_DTYPE).reshape(1)
        if self._comm is not None:
            rc = self._L.ls_comm_allgather_pose_records(self._comm, mine.ctypes.data, out.ctypes.data)
            if rc != 0:
                raise RuntimeError(f"ls_comm_allgather_pose_records failed with {rc}")
            return out
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(mine.view(np.uint8).copy())
        bufs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(bufs, t)
        return np.concatenate([b.numpy() for b in bufs]).view(RECORD_DTYPE)

    def post(self, rec):
        """Split-phase all-gather: enqueue this rank's record and return; `collect()` returns the gathered records.
        At most one exchange in flight.  Falls back to the blocking path off NCCL (gloo tests, world 1)."""
        if self._comm is None:
            self._posted = self.allgather(rec)
            return
        mine = np.array(rec, RECORD_DTYPE).reshape(1)
        self._mine = mine  # keep alive until the call returns (copied to pinned memory inside)
        rc = self._L.ls_comm_allgather_pose_records_begin(self._comm, mine.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ls_comm_allgather_pose_records_begin failed with {rc}")
        self._posted = True

    def collect(self):
        """Records of the exchange posted last, or None if nothing is in flight."""
        p = getattr(self, "_posted", None)
        if p is None:
            return None
        self._posted = None
        if p is True:
            out = np.zeros(self.world, RECORD_DTYPE)
            rc = self._L.ls_comm_allgather_pose_records_end(self._comm, out.ctypes.data)
            if rc != 0:
                raise RuntimeError(f"ls_comm_allgather_pose_records_end failed with {rc}")
            return out
        return p

    def close(self):
        self.collect()
        if self._comm is not None:
            self._L.ls_comm_destroy(self._comm)
            self._comm = None


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (step times: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(values, dtype=torch.float64, device=("cuda" if device is not None else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


# ---------------------------------------------------------------------------------------------------------------------
# The shared estimator fed from the gathered records (the reference has ONE iSAM2 for all tracks,
# reference laser_slam/include/laser_slam/incremental_estimator.hpp:67; here every rank keeps a replica and feeds it the
# same records in the same order, so the replicas stay identical without any further exchange).
ICP_SIGMAS = [0.005] * 3 + [0.0015] * 3          # reference laser_slam_ros/config/config_example.yaml:4-6
PRIOR_SIGMAS = [1e-7] * 6                        # reference laser_slam/src/laser_track.cpp:56-64


def record_pose7(rec):
    """Record delta (translation, rotation vector) -> {qw,qx,qy,qz,tx,ty,tz}, the pose layout of the C ABI."""
    d = np.asarray(rec["delta"], np.float64)
    th = float(np.linalg.norm(d[3:]))
    if th < 1e-12:
        q = np.array([1.0, 0.5 * d[3], 0.5 * d[4], 0.5 * d[5]])
    else:
        q = np.concatenate([[np.cos(0.5 * th)], np.sin(0.5 * th) / th * d[3:]])
    return np.concatenate([q / np.linalg.norm(q), d[:3]])


def _compose7(a, b):
    qa, qb = a[:4], b[:4]
    q = np.array([qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2] - qa[3] * qb[3],
                  qa[0] * qb[1] + qa[1] * qb[0] + qa[2] * qb[3] - qa[3] * qb[2],
                  qa[0] * qb[2] - qa[1] * qb[3] + qa[2] * qb[0] + qa[3] * qb[1],
                  qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1] + qa[3] * qb[0]])
    w, x, y, z = qa
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return np.concatenate([q / np.linalg.norm(q), R @ b[4:] + a[4:]])


class ReplicatedGraph:
    """Host mirror of the replicated pose graph: `feed(records)` turns one step's gathered records into new nodes and
    factors -- track t's first record anchors it with a prior at (0, 100 t, 0) (reference laser_track.cpp:163-172,
    force_priors), every later one adds the node `previous * delta` and a Cauchy ICP BetweenFactor (reference
    laser_track.cpp:431-451) -- and hands them to `sink` (an ls PoseGraph on the GPU; None just keeps the lists, which
    is what the CPU tests compare across ranks).  Records with status != 0 still extend the track (the reference keeps
    the odometry guess, laser_track.cpp:495-502)."""

    def __init__(self, world, sink=None):
        self.world, self.sink = world, sink
        self.keys, self.poses, self.factors, self.tracks = [], [], [], []
        self.last = [None] * world        # (key, pose7) of every track's newest node
        self.count = [0] * world

    @staticmethod
    def key(track, index):
        return (int(track) << 48) | int(index)

    def feed(self, records):
        new_keys, new_poses, new_tracks, new_factors = [], [], [], []
        for t in range(self.world):
            rec = records[t]
            k = self.key(t, self.count[t])
            if self.last[t] is None:
                pose = np.array([1.0, 0, 0, 0, 0.0, 100.0 * t, 0.0])
                new_factors.append(dict(type=0, key_a=k, key_b=k, meas=pose, sigma=PRIOR_SIGMAS))
            else:
                rel = record_pose7(rec)
                pose = _compose7(self.last[t][1], rel)
                new_factors.append(dict(type=1, key_a=self.last[t][0], key_b=k, meas=rel, sigma=ICP_SIGMAS, robust=1))
            new_keys.append(k); new_poses.append(pose); new_tracks.append(t)
            self.last[t] = (k, pose)
            self.count[t] += 1
        self.keys += new_keys; self.poses += new_poses; self.tracks += new_tracks; self.factors += new_factors
        if self.sink is not None:
            self.sink.add_poses(np.array(new_keys, np.uint64), np.stack(new_poses), np.array(new_tracks, np.uint32))
            self.sink.add_factors(new_factors)

    def digest(self):
        """Bytes that are equal on two ranks iff their graphs are (keys, initial values, factor table)."""
        import hashlib
        h = hashlib.sha256()
        h.update(np.array(self.keys, np.uint64).tobytes())
        h.update(np.stack(self.poses).tobytes() if self.poses else b"")
        for f in self.factors:
            h.update(np.array([f["type"], f.get("robust", 0)], np.int64).tobytes())
            h.update(np.array([f["key_a"], f["key_b"]], np.uint64).tobytes())
            h.update(np.asarray(f["meas"], np.float64).tobytes())
        return h.hexdigest()


class ShardedRegistrar:
    """One registration sharded by queries over the ranks of a node (ls_icp_register_submap_sharded).

    Every rank owns a Context + Map on its GPU and pushes the same scans; `register` is a collective that returns the
    same bits on every rank.  The only thing that travels through torch.distributed is the 64-byte IPC handle of every
    rank's exchange buffer, once, at construction; per iteration the shards meet inside the persistent kernel (stores
    into each other's buffers over NVLink)."""

    def __init__(self, ctx, rank, world):
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, world
        mine = ctx.shard_exchange_create(rank, world)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles[0] = mine
        ctx.shard_exchange_connect(handles)

    def register(self, map_, reading_id, part_ids, T_parts, T0, params=None, raise_on_convergence=True):
        return map_.register_sharded(reading_id, part_ids, T_parts, T0, params, raise_on_convergence)

    def close(self):
        """Collective: nobody frees its buffer while a peer may still store into it."""
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        from . import lib
        lib().ls_shard_exchange_close(self.ctx._h)
