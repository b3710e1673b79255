"""Op-script generator shared by the write-parity tests.

One script + one blob drive BOTH implementations through the same call sequence:
``oracle/_ref/ref_driver script`` (the unmodified reference libpgsd, compiled in place against
the MPI shim) and ``tools/pgsd_replay`` (libpgsd_b200).  The tests then compare the .gsd files,
the return-code logs and the read dumps byte for byte.  Format: see oracle/ref_driver.c.
"""
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")
REPLAY = os.path.join(REPO, "tools", "pgsd_replay")

TYPE_CODES = {  # include/pgsd.h enum pgsd_type (ref: pgsd.h:38-69)
    np.dtype(np.uint8): 1, np.dtype(np.uint16): 2, np.dtype(np.uint32): 3, np.dtype(np.uint64): 4,
    np.dtype(np.int8): 5, np.dtype(np.int16): 6, np.dtype(np.int32): 7, np.dtype(np.int64): 8,
    np.dtype(np.float32): 9, np.dtype(np.float64): 10,
}
CODE_DTYPES = {v: k for k, v in TYPE_CODES.items()}
READWRITE, READONLY, APPEND = 1, 2, 3


class Script:
    """Accumulates ops and the data blob."""

    def __init__(self):
        self.lines = []
        self.blob = bytearray()

    # -- file ops
    def create(self, application="pgsd-b200", schema="hoomd", schema_version=(1 << 16) | 4,
               flags=READWRITE, excl=0):
        self.lines.append(f"create @FILE@ {application} {schema} {schema_version} {flags} {excl}")

    def open(self, flags=READONLY):
        self.lines.append(f"open @FILE@ {flags}")

    def setbuf(self, nbytes):
        self.lines.append(f"setbuf {nbytes}")

    def setidx(self, n):
        self.lines.append(f"setidx {n}")

    def chunk(self, name, array, all_, mode="R", rows=None):
        """mode R: every rank passes the whole array; S: rows split floor/+1; X: explicit rows."""
        a = np.ascontiguousarray(array)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        n, m = a.shape
        while len(self.blob) % 16:
            self.blob.append(0)
        off = len(self.blob)
        self.blob += a.tobytes()
        line = f"chunk {name} {TYPE_CODES[a.dtype]} {m} {int(bool(all_))} {mode} {n} {off}"
        if mode == "X":
            assert rows is not None and sum(rows) == n
            line += " " + " ".join(str(r) for r in rows)
        self.lines.append(line)

    def end_frame(self):
        self.lines.append("end_frame")

    def flush(self):
        self.lines.append("flush")

    def close(self):
        self.lines.append("close")

    def nframes(self):
        self.lines.append("nframes")

    def nnames(self):
        self.lines.append("nnames")

    def find(self, frame, name):
        self.lines.append(f"find {frame} {name}")

    def read(self, frame, name, all_=0):
        self.lines.append(f"read {frame} {name} {int(bool(all_))}")

    def match(self, prefix):
        self.lines.append(f"match {prefix if prefix else '-'}")

    # -- HOOMD-schema frame, call order of SURVEY.md Appendix A.4
    def hoomd_frame(self, frame_arrays, scalars, logs=()):
        for name, arr in scalars:
            self.chunk(name, arr, all_=False, mode="R")
        for name, arr in frame_arrays.items():
            self.chunk(name, arr, all_=True, mode="S")
        for name, arr in logs:
            self.chunk(name, arr, all_=False, mode="R")
        self.end_frame()

    def write(self, workdir, tag, gsd_path):
        ops = os.path.join(workdir, f"{tag}.ops")
        blob = os.path.join(workdir, f"{tag}.blob")
        with open(ops, "w") as f:
            f.write("\n".join(self.lines).replace("@FILE@", gsd_path) + "\n")
        with open(blob, "wb") as f:
            f.write(bytes(self.blob))
        return ops, blob


def have_reference():
    return os.access(REF_DRIVER, os.X_OK)


def have_replay():
    return os.access(REPLAY, os.X_OK)


def run_reference(script, workdir, tag, nprocs, timeout=60):
    """Run the script through the compiled reference; returns (gsd_path, out_prefix)."""
    gsd = os.path.join(workdir, f"{tag}.ref.gsd")
    prefix = os.path.join(workdir, f"{tag}.ref")
    ops, blob = script.write(workdir, tag + ".ref", gsd)
    env = dict(os.environ, PGSD_SHIM_NP=str(nprocs))
    subprocess.run([REF_DRIVER, "script", ops, blob, prefix], check=True, env=env, timeout=timeout,
                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return gsd, prefix


def run_replay(script, workdir, tag, nprocs, device=False, soa=False, auto_offset=False, timeout=60):
    """Run the script through libpgsd_b200; returns (gsd_path, out_prefix)."""
    gsd = os.path.join(workdir, f"{tag}.new.gsd")
    prefix = os.path.join(workdir, f"{tag}.new")
    ops, blob = script.write(workdir, tag + ".new", gsd)
    cmd = [REPLAY, ops, blob, prefix, "--np", str(nprocs)]
    if device:
        cmd.append("--device")
    if soa:
        cmd.append("--soa")
    if auto_offset:
        cmd.append("--auto-offset")
    p = subprocess.run(cmd, timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError(f"pgsd_replay failed ({p.returncode}): {p.stderr.decode()[-2000:]}")
    return gsd, prefix


def read_bytes(path):
    with open(path, "rb") as f:
        return f.read()


def compare_runs(ref, new, nprocs=1):
    """Assert identical .gsd bytes, identical logs, identical read dumps."""
    (ref_gsd, ref_prefix), (new_gsd, new_prefix) = ref, new
    a, b = read_bytes(ref_gsd), read_bytes(new_gsd)
    if a != b:
        n = min(len(a), len(b))
        first = next((i for i in range(n) if a[i] != b[i]), n)
        raise AssertionError(f"gsd files differ: sizes {len(a)} vs {len(b)}, first difference at byte {first}")
    la, lb = read_bytes(ref_prefix + ".log"), read_bytes(new_prefix + ".log")
    assert la == lb, "logs differ:\n" + _log_diff(la.decode(), lb.decode())
    k = 0
    while True:
        found = False
        for suffix in [f".read{k}"] + [f".read{k}.r{r}" for r in range(nprocs)]:
            if os.path.exists(ref_prefix + suffix) or os.path.exists(new_prefix + suffix):
                found = True
                assert read_bytes(ref_prefix + suffix) == read_bytes(new_prefix + suffix), f"read dump {suffix} differs"
        if not found:
            break
        k += 1
    return len(a)


def _log_diff(a, b):
    out = []
    for x, y in zip(a.splitlines(), b.splitlines()):
        if x != y:
            out.append(f"  ref: {x}\n  new: {y}")
    if len(a.splitlines()) != len(b.splitlines()):
        out.append(f"  line counts {len(a.splitlines())} vs {len(b.splitlines())}")
    return "\n".join(out[:20])
