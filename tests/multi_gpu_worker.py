"""Worker of tests/test_gpu_multi.py: one process per GPU under torch.distributed.run (NCCL)."""
import hashlib
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import bench  # noqa: E402  (data generator + Dist helper shared with the benchmark)


def main():
    what, tmp = sys.argv[1], sys.argv[2]
    from pgsd_sph_b200 import _lib, comm, hoomd
    from pgsd_sph_b200.devmem import DeviceArray
    dist = bench.Dist(int(os.environ["WORLD_SIZE"]))
    lib = _lib.load()
    _lib.check(lib.pgsd_b200_device_init(dist.local), "device_init")
    comm.init_nccl(dist.rank, dist.world, dist.bcast_bytes, dist.local)
    assert lib.pgsd_b200_comm_kind() == b"nccl" and lib.pgsd_b200_comm_size() == dist.world
    n_total = 3 * 1024 * 1024 + 12345          # ragged split, > 1 piece per chunk and rank
    rows, start = bench.rank_rows(n_total, dist.world, dist.rank)
    n = rows[dist.rank]
    cols = bench.make_soa(n_total, start, n, 99)
    if what == "write":
        mine, ref, blob = (os.path.join(tmp, x) for x in ("mine.gsd", "ref.gsd", "blob.bin"))
        for frame_count in (1,):
            bench.write_one_frame(dist, mine, n_total, cols, rows, nlogs=3)
        if dist.rank == 0:
            with open(blob, "wb") as fh:
                fh.truncate(40 * n_total)
        dist.barrier()
        fd = os.open(blob, os.O_RDWR)
        for j, c in enumerate(cols):
            os.pwrite(fd, c.tobytes(), (j * n_total + start) * 4)
        os.close(fd)
        dist.barrier()
        if dist.rank == 0:
            t = bench.ref_driver_bench(ref, n_total, 1, blob, dist.world, nlogs=3)
            assert t is not None, "oracle/_ref/ref_driver missing"
            a, b = open(mine, "rb").read(), open(ref, "rb").read()
            assert len(a) == len(b), (len(a), len(b))
            assert hashlib.sha256(a).hexdigest() == hashlib.sha256(b).hexdigest(), "file differs from the reference's"
    else:
        from oracle import reorder_oracle
        ids = cols[9]
        pos = np.ascontiguousarray(np.stack(cols[0:3], axis=1))
        tag = (ids ^ np.uint32(0x5bd1e995)).astype(np.uint32)
        first, sid, out = hoomd.reorder_by_id_distributed(DeviceArray.from_numpy(ids),
                                                          {"pos": DeviceArray.from_numpy(pos), "tag": DeviceArray.from_numpy(tag),
                                                           "dens": DeviceArray.from_numpy(cols[6])})
        # the oracle needs the whole frame: every rank regenerates it (partition-independent generator)
        full = bench.make_soa(n_total, 0, n_total, 99)
        fpos = np.ascontiguousarray(np.stack(full[0:3], axis=1))
        shares = reorder_oracle.reorder_distributed(full[9], {"pos": fpos, "tag": (full[9] ^ np.uint32(0x5bd1e995)).astype(np.uint32),
                                                              "dens": full[6]}, dist.world)
        f0, want_ids, want = shares[dist.rank]
        assert first == f0, (first, f0)
        assert sid.to_numpy().tobytes() == want_ids.tobytes()
        for k in want:
            assert out[k].to_numpy().tobytes() == want[k].tobytes(), k
    dist.barrier()
    if dist.rank == 0:
        print(f"WORKER-OK {what} ranks={dist.world}", flush=True)
    lib.pgsd_b200_comm_finalize()
    dist.close()
    lib.pgsd_b200_shutdown()


if __name__ == "__main__":
    main()
