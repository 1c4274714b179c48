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
