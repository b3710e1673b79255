"""One rank of the distributed-reorder GPU test: P processes share cuda:0 and talk through the built-in
shared-memory communicator; each holds one partition of the frame (the split rule of benchmark-write.cc:33-37),
calls pgsd_b200_reorder_distributed and writes what it owns afterwards to <outdir>/rank<r>_<case>.npz.
Usage: dist_reorder_worker.py RANK NPROCS SEGMENT OUTDIR"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from pgsd_sph_b200 import _lib
from pgsd_sph_b200.devmem import DeviceArray


PYWRAP_N = 200003


def cases():
    """name, N, key generator(rng, N) -> uint32 ids of the WHOLE frame; the same on every rank."""
    yield "perm_300k", 300001, lambda rng, n: rng.permutation(n).astype(np.uint32)
    yield "perm_5000", 5000, lambda rng, n: rng.permutation(n).astype(np.uint32)          # fewer buckets than ranks can happen
    yield "perm_1M", 1 << 20, lambda rng, n: rng.permutation(n).astype(np.uint32)          # the shared copy grows: re-export
    yield "sorted_70k", 70000, lambda rng, n: np.arange(n, dtype=np.uint32)
    yield "one_dup", 90000, lambda rng, n: np.concatenate([rng.permutation(n - 1), [777]]).astype(np.uint32)   # -> 1
    yield "dup_in_partial_bucket", 90000, lambda rng, n: np.concatenate([rng.permutation(n - 1), [89500]]).astype(np.uint32)  # -> 1 (found by the placement)
    yield "out_of_range", 50000, lambda rng, n: (rng.permutation(n) + 4096).astype(np.uint32)                  # -> 1
    yield "perm_300k_again", 300001, lambda rng, n: rng.permutation(n).astype(np.uint32)   # after failures: state is clean


def main():
    rank, nprocs, segment, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    lib = _lib.load()
    _lib.check(lib.pgsd_b200_comm_init_shm(rank, nprocs, segment.encode()), "comm_init_shm")
    _lib.check(lib.pgsd_b200_device_init(0), "device_init")
    for ci, (name, n, gen) in enumerate(cases()):
        rng = np.random.default_rng(1000 + ci)
        ids = gen(rng, n)
        pos = rng.standard_normal((n, 3)).astype(np.float32)
        tag = (ids ^ np.uint32(0x9e3779b9)).astype(np.uint32)
        dens = rng.standard_normal(n)                      # float64: 2-word rows
        lo = sum(n // nprocs + (1 if r < n % nprocs else 0) for r in range(rank))
        cnt = n // nprocs + (1 if rank < n % nprocs else 0)
        sl = slice(lo, lo + cnt)
        cap = -(-(-(-n // 1024)) // nprocs) * 1024        # ceil(ceil(n / 1024) / P) * 1024
        d_ids = DeviceArray.from_numpy(ids[sl])
        ins = [DeviceArray.from_numpy(np.ascontiguousarray(a[sl])) for a in (pos, tag, dens)]
        outs = [DeviceArray((cap, 3), np.float32), DeviceArray((cap,), np.uint32), DeviceArray((cap,), np.float64)]
        d_sorted = DeviceArray((cap,), np.uint32)
        fields = (_lib.Field * 3)(_lib.Field(ins[0].ptr, outs[0].ptr, 12), _lib.Field(ins[1].ptr, outs[1].ptr, 4),
                                  _lib.Field(ins[2].ptr, outs[2].ptr, 8))
        n_out, id_first = C.c_uint64(0), C.c_uint64(0)
        rc = lib.pgsd_b200_reorder_distributed(cnt, d_ids.ptr, cap, C.byref(n_out), C.byref(id_first), d_sorted.ptr, 3, fields, None)
        if rc < 0:
            raise RuntimeError(f"{name}: rc={rc}: {lib.pgsd_b200_last_error().decode()}")
        k = int(n_out.value) if rc == 0 else 0
        np.savez(os.path.join(outdir, f"rank{rank}_{name}.npz"), rc=rc, n_out=k, id_first=int(id_first.value),
                 ids=d_sorted.to_numpy()[:k], pos=outs[0].to_numpy()[:k], tag=outs[1].to_numpy()[:k], dens=outs[2].to_numpy()[:k])
        for a in ins + outs + [d_ids, d_sorted]:
            a.free()
    # the Python-level wrapper (pgsd_sph_b200.hoomd.reorder_by_id_distributed) on one more frame
    from pgsd_sph_b200 import hoomd
    n = PYWRAP_N
    rng = np.random.default_rng(77)
    ids = rng.permutation(n).astype(np.uint32)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    lo = sum(n // nprocs + (1 if r < n % nprocs else 0) for r in range(rank))
    cnt = n // nprocs + (1 if rank < n % nprocs else 0)
    first, sid, out = hoomd.reorder_by_id_distributed(DeviceArray.from_numpy(ids[lo:lo + cnt]),
                                                      {"pos": DeviceArray.from_numpy(pos[lo:lo + cnt])})
    np.savez(os.path.join(outdir, f"rank{rank}_pywrap.npz"), id_first=first, ids=sid.to_numpy(), pos=out["pos"].to_numpy())
    # config 3 read-back end to end: every rank reads its row slice of a file (written by the parent test) straight
    # into device memory, then the distributed reorder
    path = os.path.join(outdir, "frame.gsd")
    if os.path.exists(path):
        from pgsd_sph_b200 import fl
        with fl.open(path, 'r', 'pgsd-b200', 'hoomd', [1, 4]) as f:
            first, sid, out = hoomd.read_frame_distributed(f, 0)
        np.savez(os.path.join(outdir, f"rank{rank}_file.npz"), id_first=first, ids=sid.to_numpy(),
                 **{k: v.to_numpy() for k, v in out.items()})
    lib.pgsd_b200_comm_finalize()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
