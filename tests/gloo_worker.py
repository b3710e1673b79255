"""One rank of the world_size-2 gloo test: writes the golden HOOMD frames through the Python drop-in
API with the 'host' communicator (torch.distributed all-gather callback), rows split over the ranks."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main(path, auto):
    import torch.distributed as dist
    dist.init_process_group("gloo")
    from golden.make_golden import GOLDEN_FRAMES, GOLDEN_N
    from pgsd_sph_b200 import comm, fl, synth
    rank, nprocs = comm.init_torch_distributed()
    rows = synth.split_rows(GOLDEN_N, nprocs)
    start = synth.row_starts(rows)[rank]
    with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
        for i in range(GOLDEN_FRAMES):
            fr = synth.make_frame(GOLDEN_N, i)
            for k, a in synth.frame_scalars(GOLDEN_N, i):
                f.write_chunk(k, a, write_all=False)
            for k, a in fr.items():
                mine = np.ascontiguousarray(a[start:start + rows[rank]])
                if auto:
                    f.write_chunk(k, mine, offset='auto')
                else:
                    f.write_chunk(k, mine, offset=rows, rank=rank)
            f.write_chunk("log/value/kinetic_energy", np.array([0.5 * i + 1.25], dtype=np.float32), write_all=False)
            f.write_chunk("log/value/potential_energy", np.array([-3.0 * i], dtype=np.float32), write_all=False)
            f.end_frame()
        assert f.nframes == GOLDEN_FRAMES
    # every rank can read (replicated index): partitioned read of its own rows
    with fl.open(path, 'r') as f:
        got = f.read_chunk(GOLDEN_FRAMES - 1, "particles/position", N=rows[rank], M=3, offset=start, r_all=True)
        want = synth.make_frame(GOLDEN_N, GOLDEN_FRAMES - 1)["particles/position"][start:start + rows[rank]]
        assert got.tobytes() == want.tobytes()
    comm.finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] == "auto")
