"""Host-only tests of the read-ahead state machine (pgsd_sph_b200/csrc/read_ahead.cpp).  It
is the same code that stages into device memory when PGSD_B200_READ_AHEAD=1 -- here driven with host memory through
pgsd_b200_read_ahead_host_read, so that the CPU suite can check every access order against the file's bytes and
hammer it from several threads.  The pattern it serves is the reference's benchmark-read.cc:46-120 (equally sized
reads at a constant file stride).

The multi-threaded cases are the regression tests of a stall seen on 8 GPUs (profiles/r5_bench_n8_stalled.err):
pgsd.hoomd's frame-prefetch thread and the main thread were in the front end at the same time, one dropped the range
the other was waiting for, the waiter never woke, and the bench hung joining that thread.  (With the front-end lock
and the bounded wait taken out again, test_threads_reading_at_once_with_resets_in_between stalls within seconds.)"""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from pgsd_sph_b200 import _lib

CHUNK = 384 * 1024          # >= the 256 KiB below which reads are never fetched ahead
NCHUNK = 40


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.fixture()
def datafile(tmp_path):
    rng = np.random.default_rng(7)
    data = rng.integers(0, 256, size=CHUNK * NCHUNK + 12345, dtype=np.uint8)
    path = str(tmp_path / "blob.bin")
    data.tofile(path)
    return path, data


def _stats(lib):
    h, i, d = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.pgsd_b200_read_ahead_host_stats(C.byref(h), C.byref(i), C.byref(d))
    return h.value, i.value, d.value


def _read(lib, fd, off, n):
    buf = np.empty(n, dtype=np.uint8)
    rc = lib.pgsd_b200_read_ahead_host_read(fd, buf.ctypes.data_as(C.c_void_p), n, off)
    assert rc == 0, rc
    return buf


def test_sequential_strided_reverse_and_random_orders(lib, datafile):
    path, data = datafile
    fd = os.open(path, os.O_RDONLY)
    try:
        lib.pgsd_b200_read_ahead_host_reset()
        h0 = _stats(lib)[0]
        for k in range(NCHUNK):                                        # stride = size
            assert _read(lib, fd, k * CHUNK, CHUNK).tobytes() == data[k * CHUNK:(k + 1) * CHUNK].tobytes(), k
        h1 = _stats(lib)[0]
        assert h1 - h0 >= NCHUNK - 4                                   # all but the reads that establish the pattern
        part = CHUNK - 4096                                            # a slice of every second chunk, odd offset
        for k in range(0, NCHUNK, 2):
            off = k * CHUNK + 777
            assert _read(lib, fd, off, part).tobytes() == data[off:off + part].tobytes(), k
        h2 = _stats(lib)[0]
        assert h2 - h1 >= NCHUNK // 2 - 4
        for k in reversed(range(NCHUNK)):                              # negative stride
            assert _read(lib, fd, k * CHUNK, CHUNK).tobytes() == data[k * CHUNK:(k + 1) * CHUNK].tobytes(), k
        for k in np.random.default_rng(1).permutation(NCHUNK):         # no pattern
            k = int(k)
            assert _read(lib, fd, k * CHUNK, CHUNK).tobytes() == data[k * CHUNK:(k + 1) * CHUNK].tobytes(), k
        tail = len(data) - NCHUNK * CHUNK                              # small read: never staged
        assert _read(lib, fd, NCHUNK * CHUNK, tail).tobytes() == data[NCHUNK * CHUNK:].tobytes()
        hits, issued, dropped = _stats(lib)
        assert issued >= hits and dropped <= issued
    finally:
        os.close(fd)
        lib.pgsd_b200_read_ahead_host_reset()


def test_replaced_file_is_never_served_from_staging(lib, tmp_path):
    path = str(tmp_path / "r.bin")
    rng = np.random.default_rng(3)
    for gen in range(3):
        data = rng.integers(0, 256, size=CHUNK * 12, dtype=np.uint8)
        data.tofile(path)                                              # same name, same size, new bytes
        fd = os.open(path, os.O_RDONLY)
        try:
            for k in range(8):                                         # leaves fetched ranges behind ...
                assert _read(lib, fd, k * CHUNK, CHUNK).tobytes() == data[k * CHUNK:(k + 1) * CHUNK].tobytes(), (gen, k)
        finally:
            os.close(fd)
            lib.pgsd_b200_read_ahead_host_reset()                      # ... which every close of a handle drops


@pytest.mark.parametrize("nthreads", [2, 4])
def test_threads_reading_at_once_with_resets_in_between(lib, datafile, nthreads):
    """Several threads in the front end at once (ascending, descending, strided, random), a further thread calling
    reset() all the time (what every pgsd_open / pgsd_close does): every read returns the file's bytes and every
    thread terminates."""
    path, data = datafile
    lib.pgsd_b200_read_ahead_host_reset()
    bad, stop = [], threading.Event()

    def reader(t):
        fd = os.open(path, os.O_RDONLY)
        try:
            orders = [list(range(NCHUNK)), list(range(NCHUNK - 1, -1, -1)), list(range(0, NCHUNK, 3)),
                      [int(k) for k in np.random.default_rng(t).permutation(NCHUNK)]]
            for rep in range(6):
                for k in orders[(t + rep) % len(orders)]:
                    got = _read(lib, fd, k * CHUNK, CHUNK)
                    if got.tobytes() != data[k * CHUNK:(k + 1) * CHUNK].tobytes():
                        bad.append((t, rep, k))
        except Exception as e:  # noqa: BLE001
            bad.append(repr(e))
        finally:
            os.close(fd)

    def resetter():
        while not stop.is_set():
            lib.pgsd_b200_read_ahead_host_reset()
            stop.wait(0.002)

    # daemon threads: should the state machine ever stall again, the failed assertion below ends the run
    ts = [threading.Thread(target=reader, args=(t,), daemon=True) for t in range(nthreads)]
    r = threading.Thread(target=resetter, daemon=True)
    r.start()
    for t in ts:
        t.start()
    for t in ts:
        t.join(180)
    stop.set()
    r.join(30)
    assert not any(t.is_alive() for t in ts) and not r.is_alive()
    assert not bad, bad[:5]
    lib.pgsd_b200_read_ahead_host_reset()


def test_two_files_alternating(lib, tmp_path):
    rng = np.random.default_rng(11)
    paths, datas = [], []
    for j in range(2):
        d = rng.integers(0, 256, size=CHUNK * 10, dtype=np.uint8)
        p = str(tmp_path / f"f{j}.bin")
        d.tofile(p)
        paths.append(p)
        datas.append(d)
    fds = [os.open(p, os.O_RDONLY) for p in paths]
    try:
        for k in range(10):
            for j in range(2):
                assert _read(lib, fds[j], k * CHUNK, CHUNK).tobytes() == datas[j][k * CHUNK:(k + 1) * CHUNK].tobytes()
        for k in range(10):
            assert _read(lib, fds[1], k * CHUNK, CHUNK).tobytes() == datas[1][k * CHUNK:(k + 1) * CHUNK].tobytes()
    finally:
        for fd in fds:
            os.close(fd)
        lib.pgsd_b200_read_ahead_host_reset()
