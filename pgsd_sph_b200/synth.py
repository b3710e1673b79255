"""Synthetic SPH-like particle frames.

Seeded, deterministic inputs shared by the parity tests, ``bench.py`` (both arms) and
``__graft_entry__.smoke()``.  Pure numpy; no reference code involved.

Per particle: position f32x3 uniform[0,L)^3, velocity f32x3 N(0,1), density f32,
pressure f32, typeid u32 in {0,1,2}, id u32 = a seeded permutation of 0..N-1 (dense,
unique, unsorted).  Rank r owns floor(N/P) (+1 if r < N mod P) consecutive rows of that
permuted order -- the split rule of the reference's only multi-rank caller
(/root/reference/pgsd/scripts/benchmark-write.cc:33-37).
"""
import numpy as np

BOX_L = 10.0
SEED0 = 20261018

FIELDS = (  # chunk name, dtype, M  -- call order of SURVEY.md Appendix A.4
    ("particles/position", np.float32, 3),
    ("particles/velocity", np.float32, 3),
    ("particles/typeid", np.uint32, 1),
    ("particles/density", np.float32, 1),
    ("particles/pressure", np.float32, 1),
    ("log/particles/id", np.uint32, 1),
)
BYTES_PER_PARTICLE = 40


def split_rows(n, nprocs):
    """Rows per rank: floor(n/P) (+1 if rank < n mod P) (benchmark-write.cc:33-37)."""
    base, rem = divmod(int(n), int(nprocs))
    return [base + (1 if r < rem else 0) for r in range(nprocs)]


def row_starts(rows):
    """Exclusive prefix sum of rows (benchmark-write.cc:43-45, fl.pyx:598)."""
    out, acc = [], 0
    for r in rows:
        out.append(acc)
        acc += r
    return out


def make_frame(n, frame=0, cheap=False):
    """Return dict of global (unsplit) arrays for one frame.

    ``cheap=True`` replaces the normal draws by uniform ones (4x faster to generate at
    64 Mi particles); the byte layout and dtypes are identical.
    """
    rng = np.random.Generator(np.random.PCG64(SEED0 + int(frame)))
    ids = rng.permutation(n).astype(np.uint32)
    position = (rng.random((n, 3), dtype=np.float32) * np.float32(BOX_L)).astype(np.float32)
    if cheap:
        velocity = (rng.random((n, 3), dtype=np.float32) - np.float32(0.5)).astype(np.float32)
        noise = (rng.random(n, dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    else:
        velocity = rng.standard_normal((n, 3), dtype=np.float32)
        noise = rng.standard_normal(n, dtype=np.float32)
    density = (np.float32(1000.0) * (np.float32(1.0) + np.float32(0.01) * noise)).astype(np.float32)
    pressure = (np.float32(1500.0 ** 2 * 1e-6) * (density - np.float32(1000.0))).astype(np.float32)
    typeid = rng.choice(np.array([0, 1, 2], dtype=np.uint32), size=n, p=[0.8, 0.15, 0.05]).astype(np.uint32)
    return {
        "particles/position": position,
        "particles/velocity": velocity,
        "particles/typeid": typeid,
        "particles/density": density,
        "particles/pressure": pressure,
        "log/particles/id": ids,
    }


def frame_scalars(n, frame=0):
    """The small root-owned chunks written ``all=false`` each frame (Appendix A.4)."""
    return [
        ("configuration/step", np.array([10 * frame], dtype=np.uint64)),
        ("configuration/dimensions", np.array([3], dtype=np.uint8)),
        ("configuration/box", np.array([BOX_L, BOX_L, BOX_L, 0, 0, 0], dtype=np.float32)),
        ("particles/N", np.array([n], dtype=np.uint32)),
    ]


def to_soa(frame):
    """Split a frame into the 10 SoA component arrays a particle solver holds.

    Order: pos_x pos_y pos_z vel_x vel_y vel_z density pressure typeid id -- the blob
    layout oracle/ref_driver.c's bench mode reads.
    """
    p, v = frame["particles/position"], frame["particles/velocity"]
    return [
        np.ascontiguousarray(p[:, 0]), np.ascontiguousarray(p[:, 1]), np.ascontiguousarray(p[:, 2]),
        np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1]), np.ascontiguousarray(v[:, 2]),
        frame["particles/density"], frame["particles/pressure"],
        frame["particles/typeid"], frame["log/particles/id"],
    ]
