"""Communicator set-up helpers for the Python layer (stands where ``MPI_Init`` / ``MPI_COMM_WORLD``
stand in the reference: /root/reference/pgsd/pgsd/pgsd.c:1487-1488, scripts/benchmark-write.cc:23).

* :func:`init_nccl` -- production: one process per GPU; the 128-byte NCCL id is shipped by any
  caller-supplied broadcast (torch.distributed, MPI, a file).
* :func:`init_torch_distributed` -- "host" transport over an initialised ``torch.distributed``
  process group (gloo on CPU-only hosts, used by the world_size-2 tests): the library calls back
  into an all-gather of uint64 vectors.
"""
import ctypes as C

from . import _lib

_keep = []  # callbacks must outlive the communicator


def init_nccl(rank, nprocs, broadcast_bytes, cuda_device):
    """``broadcast_bytes(b: bytes, nbytes) -> bytes`` broadcasts rank 0's buffer to every rank."""
    lib = _lib.load()
    uid = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.pgsd_b200_nccl_unique_id(uid), "pgsd_b200_nccl_unique_id")
    raw = broadcast_bytes(uid.raw, 128)
    _lib.check(lib.pgsd_b200_comm_init_nccl(rank, nprocs, raw, cuda_device), "pgsd_b200_comm_init_nccl")


def init_torch_distributed(group=None):
    """Use an initialised torch.distributed process group (any backend) as the rank communicator."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    rank, nprocs = dist.get_rank(group), dist.get_world_size(group)

    def to_i64(v):
        return v - (1 << 64) if v >= (1 << 63) else v

    def allgather_u64(ctx, send, recv, n):
        try:
            mine = torch.tensor([to_i64(send[i]) for i in range(n)], dtype=torch.int64)
            outs = [torch.empty(n, dtype=torch.int64) for _ in range(nprocs)]
            dist.all_gather(outs, mine, group=group)
            k = 0
            for t in outs:
                for v in t.tolist():
                    recv[k] = v & 0xFFFFFFFFFFFFFFFF
                    k += 1
            return 0
        except Exception:
            return -1

    cb = _lib.ALLGATHER_FN(allgather_u64)
    _keep.append(cb)
    _lib.check(lib.pgsd_b200_comm_init_host(rank, nprocs, cb, None), "pgsd_b200_comm_init_host")
    return rank, nprocs


def finalize():
    _lib.load().pgsd_b200_comm_finalize()
