#!/usr/bin/env python
"""Where a config-5 frame's time goes: caller-side time per API call, and (PGSD_B200_TRACE) the K1 / D2H / file-piece
intervals of every frame on one clock.  Development tool."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
trace = os.path.join(os.environ.get("OUT", "/tmp"), "small_trace.json")
if os.environ.get("TRACE", "1") == "1":
    os.environ["PGSD_B200_TRACE"] = trace
import numpy as np
import bench
from pgsd_sph_b200 import _lib, fl, synth
from pgsd_sph_b200.devmem import DeviceArray
lib = _lib.load(); _lib.check(lib.pgsd_b200_device_init(0), "init")
n, frames = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cols = bench.make_soa(n, 0, n, 5)
d = [DeviceArray.from_numpy(c) for c in cols]
path = os.path.join(bench.bench_dir(), "prof_small.gsd")
logs = [("log/value/v%d" % k, np.array([k], dtype=np.float32)) for k in range(8)]
t = {"head": 0.0, "soa": 0.0, "tail": 0.0, "end": 0.0}
with fl.open(path, 'w', 'pgsd-b200', 'hoomd', [1, 4]) as f:
    prep = f.prepare_frame_soa([(nm, [d[j] for j in idx], dt, None, True) for nm, idx, dt in bench.SOA_CHUNKS])
    scal = synth.frame_scalars(n, 0); step = scal[0][1]
    head = f.prepare_chunks([(k, a, None, False) for k, a in scal])
    tail = f.prepare_chunks([(k, a, None, False) for k, a in logs])
    w0 = time.perf_counter()
    for i in range(frames):
        step[0] = 10 * i
        a = time.perf_counter(); f.write_prepared(head)
        b = time.perf_counter(); f.write_frame_soa(prep)
        c = time.perf_counter(); f.write_prepared(tail)
        e = time.perf_counter(); f.end_frame()
        g = time.perf_counter()
        t["head"] += b - a; t["soa"] += c - b; t["tail"] += e - c; t["end"] += g - e
    loop = time.perf_counter() - w0
    f.flush()
    total = time.perf_counter() - w0
st = _lib.Stats(); lib.pgsd_b200_get_stats(st)
print("frames", frames, "loop us/frame %.1f" % (1e6 * loop / frames), "incl. final flush %.1f" % (1e6 * total / frames))
print("caller us/frame:", {k: round(1e6 * v / frames, 2) for k, v in t.items()}, "commit_wait_s", st.commit_wait_s)
os.unlink(path)
lib.pgsd_b200_shutdown()
if os.environ.get("TRACE", "1") != "1":
    sys.exit(0)
ev = json.load(open(trace))["traceEvents"]
for cat in "KDF":
    xs = [e for e in ev if e["cat"] == cat]
    if not xs:
        continue
    dur = np.array([e["dur"] for e in xs]); ts = np.array(sorted(e["ts"] for e in xs))
    gaps = np.diff(ts)
    print(cat, "n=%d dur us: median %.1f p90 %.1f; start-to-start gap us: median %.1f p90 %.1f" % (
        len(xs), np.median(dur), np.percentile(dur, 90), np.median(gaps), np.percentile(gaps, 90)))
# latency K1 start -> file piece end, per frame
k = {e["args"]["frame"]: e for e in ev if e["cat"] == "K"}
fe = {}
for e in ev:
    if e["cat"] == "F":
        fe[e["args"]["frame"]] = max(fe.get(e["args"]["frame"], 0), e["ts"] + e["dur"])
lat = np.array([fe[i] - k[i]["ts"] for i in k if i in fe])
print("K1 start -> last byte in file: median %.0f us, p90 %.0f us" % (np.median(lat), np.percentile(lat, 90)))
