"""K3 file stage on the host (pgsd_sph_b200/csrc/file_stage.cpp), without a GPU: pieces written the way the
writer threads write them -- whole pages through per-piece shared mappings, partial pages through pwrite -- must
leave exactly the bytes a plain sequential write leaves, from any number of threads and processes, at any
alignment.  This is the stress test VERDICT r1 asked for after one unexplained byte mismatch of the round-1
scheme (arbitrary byte ranges mapped by several writers); `PGSD_STRESS_ITERS=2000 pytest tests/test_file_stage.py`
is the long run recorded under profiles/."""
import ctypes as C
import multiprocessing as mp
import atexit
import os
import shutil
import threading

import numpy as np
import pytest

from pgsd_sph_b200 import _lib

PAGE = os.sysconf("SC_PAGESIZE")
MODES = {"auto": 0, "pwrite": 1, "mmap": 2}


def pattern(off, n):
    """Bytes of the reference image at file offsets [off, off + n): a function of the offset only."""
    x = np.arange(off, off + n, dtype=np.uint64)
    return (x ^ (x >> np.uint64(9)) ^ (x >> np.uint64(17))).astype(np.uint8)


def cut(off, nbytes, piece):
    """The stager's cuts (file_first_piece_len): interior boundaries on page boundaries of the file."""
    out, done = [], 0
    while done < nbytes:
        ln = min(nbytes - done, piece - ((off + done) % PAGE) if done == 0 else piece)
        out.append((off + done, ln))
        done += ln
    return out


def targets(tmp_path):
    dirs = [str(tmp_path)]
    if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK):
        d = os.path.join("/dev/shm", "pgsd_fs_%d_%s" % (os.getpid(), tmp_path.name))
        os.makedirs(d, exist_ok=True)
        atexit.register(shutil.rmtree, d, True)   # tmp_path is pytest's to clean; this one is ours
        dirs.append(d)
    return dirs


def write_pieces(path, pieces, mode, nthreads):
    lib = _lib.load()
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    errs = []

    def work(t):
        for off, ln in pieces[t::nthreads]:
            buf = pattern(off, ln)
            rc = lib.pgsd_b200_file_stage_write(fd, buf.ctypes.data_as(C.c_void_p), off, ln, mode)
            if rc != 0:
                errs.append(rc)

    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    os.close(fd)
    return errs


@pytest.mark.parametrize("mode", ["pwrite", "mmap", "auto"])
def test_pieces_at_every_alignment_equal_a_sequential_write(tmp_path, mode):
    rng = np.random.default_rng(7)
    for d in targets(tmp_path):
        path = os.path.join(d, "a.bin")
        pieces, off = [], int(rng.integers(1, 9000))
        start = off
        for ln in [0, 1, PAGE - 1, PAGE, PAGE + 1, (1 << 20) - 1, (1 << 20) + 5, 3 * (1 << 20) + 123, 2 * (1 << 20), 7, 5 * (1 << 20) + 4097]:
            pieces.append((off, ln))
            off += ln
        assert write_pieces(path, pieces, MODES[mode], 4) == []
        got = np.fromfile(path, dtype=np.uint8)
        assert len(got) == off
        assert (got[:start] == 0).all()           # the hole before the first piece reads as zeros
        assert got[start:].tobytes() == pattern(start, off - start).tobytes()
        os.unlink(path)


def _rank(args):
    path, regions, rank, mode, nthreads, piece = args
    pieces = []
    for off, nbytes in regions[rank]:
        pieces += cut(off, nbytes, piece)
    return write_pieces(path, pieces, mode, nthreads)


def _one_round(path, seed, nranks, mode, nthreads=4, piece=2 << 20):
    """`nranks` processes x `nthreads` threads write interleaved, unaligned regions of one new file (per-particle
    chunks: rank r's rows follow rank r-1's), rank 0 also writes small "metadata" ranges between them."""
    rng = np.random.default_rng(seed)
    regions, off = [[] for _ in range(nranks)], int(rng.integers(0, 3 * PAGE))
    start = off
    for chunk in range(3):
        for r in range(nranks):
            nbytes = int(rng.integers(1, 4 * (1 << 20)))
            regions[r].append((off, nbytes))
            off += nbytes
        meta = int(rng.integers(1, 200))            # index / small buffered chunks: rank 0, pwrite-sized
        regions[0].append((off, meta))
        off += meta
    if os.path.exists(path):
        os.unlink(path)
    with mp.get_context("fork").Pool(nranks) as pool:
        errs = pool.map(_rank, [(path, regions, r, mode, nthreads, piece) for r in range(nranks)])
    assert all(e == [] for e in errs), errs
    got = np.fromfile(path, dtype=np.uint8)
    want = pattern(start, off - start)
    ok = len(got) == off and (got[:start] == 0).all() and got[start:].tobytes() == want.tobytes()
    os.unlink(path)
    if ok:
        return None
    # describe the damage: which byte ranges differ, what is there, whose region it is
    if len(got) != off:
        return f"size {len(got)} != {off}"
    diff = np.nonzero(got[start:] != want)[0] + start
    runs = np.split(diff, np.nonzero(np.diff(diff) > 1)[0] + 1)
    desc = []
    for r in runs[:6]:
        a, b = int(r[0]), int(r[-1]) + 1
        owner = [(rk, o, nb) for rk in range(nranks) for (o, nb) in regions[rk] if o < b and a < o + nb]
        desc.append(f"[{a},{b}) len {b - a} page_off {a % PAGE} zeros={bool((got[a:b] == 0).all())} owner(rank,off,bytes)={owner}")
    return f"{len(diff)} bytes differ in {len(runs)} run(s): " + "; ".join(desc)


@pytest.mark.parametrize("mode", ["mmap", "auto", "pwrite"])
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_concurrent_ranks_and_threads_leave_the_sequential_image(tmp_path, nranks, mode):
    iters = int(os.environ.get("PGSD_STRESS_ITERS", "2"))
    bad = []
    for d in targets(tmp_path):
        for i in range(iters):
            why = _one_round(os.path.join(d, "s.bin"), 1000 * nranks + i, nranks, MODES[mode])
            if why is not None:
                bad.append((d, i, why))
    assert bad == [], bad


def test_ceiling_probe_runs_on_the_host(tmp_path):
    lib = _lib.load()
    s, thr, mapped = C.c_double(), C.c_int(), C.c_int()
    path = os.path.join(str(tmp_path), "c.bin").encode()
    assert lib.pgsd_b200_file_stage_ceiling(path, 12345, 40 << 20, C.byref(s), C.byref(thr), C.byref(mapped)) == 0
    assert s.value > 0 and thr.value >= 1
    assert os.path.getsize(path) == 12345 + (40 << 20)
