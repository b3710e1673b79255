"""Seeded random op scripts for the write-parity tests (and tests/golden/make_golden.py).

The generator stays inside the behaviour the reference defines.  At P > 1 it avoids what
desynchronises the reference's own ranks (found by running it, see DESIGN.md "reference
hazards"): zero-size buffered chunks (non-root ranks leave pgsd_flush_write_buffer early,
pgsd.c:1126-1133), lookups or flushes in the middle of a frame, and name matching on a writable
file (non-root ranks hold no namelist, pgsd.c:2590).
"""
import random

import numpy as np

from opscript import READONLY, Script

DTYPES = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64,
          np.float32, np.float64]


def random_script(seed, nprocs, lookups=True, max_frames=40, max_names=60):
    rng = random.Random(seed)
    nr = np.random.default_rng(seed)
    P = nprocs
    s = Script()
    s.create()
    if rng.random() < 0.4:
        s.setbuf(rng.choice([64, 300, 1000, 4096]))
    if rng.random() < 0.3:
        s.setidx(rng.choice([1, 3, 10]))
    names = [f"n{rng.randrange(10 ** 6)}/{'x' * rng.randrange(0, 40)}" for _ in range(rng.randrange(3, max_names))]
    written = []
    for f in range(rng.randrange(1, max_frames)):
        for _ in range(rng.randrange(0, 8)):
            name = rng.choice(names)
            dt = rng.choice(DTYPES)
            M = rng.choice([1, 1, 1, 2, 3, 4, 7])
            N = rng.choice([0, 1, 1, 2, 5, 17, 100, 1000])
            a = nr.integers(0, 200, size=(N, M)).astype(dt)
            if rng.random() < 0.5:
                mode = rng.choice(["S", "X"])
                rows = None
                if mode == "X":
                    cuts = sorted(rng.randrange(0, N + 1) for _ in range(P - 1))
                    rows = [hi - lo for lo, hi in zip([0] + cuts, cuts + [N])]
                s.chunk(name, a, True, mode, rows)
            else:
                if P > 1 and N == 0:
                    a = nr.integers(0, 200, size=(1, M)).astype(dt)
                s.chunk(name, a, False, "R")
            written.append((f, name))
            if lookups and P == 1 and rng.random() < 0.05:
                s.flush()
            if lookups and P == 1 and rng.random() < 0.05:
                s.find(*rng.choice(written))
        s.end_frame()
        if lookups and rng.random() < 0.08:
            s.flush()
        if lookups and rng.random() < 0.08 and written:
            s.find(*rng.choice(written))
        if rng.random() < 0.1:
            s.nframes()
            s.nnames()
        if lookups and rng.random() < 0.1 and written:
            ff, nn = rng.choice(written)
            s.read(ff, nn, 0)
    s.close()
    s.open(READONLY)
    s.nframes()
    s.nnames()
    for _ in range(10):
        if written:
            ff, nn = rng.choice(written)
            s.find(ff, nn)
            s.read(ff, nn, rng.choice([0, 1]))
    s.match("")
    s.match("n1")
    s.close()
    return s
