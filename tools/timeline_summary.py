#!/usr/bin/env python
"""Text timeline of consecutive frames from a PGSD_B200_TRACE file (Chrome trace written by the library): per frame
the K1 launch, the span of its D2H pieces and the span of its file pieces on one clock, and how much of each frame's
file stage ran while the next frames' K1 / D2H were already going on.
    PGSD_B200_TRACE=trace.json python bench.py --legs write --steps 3 ; python tools/timeline_summary.py trace.json [first] [count]"""
import json, sys
ev = json.load(open(sys.argv[1]))["traceEvents"]
first = int(sys.argv[2]) if len(sys.argv) > 2 else None
count = int(sys.argv[3]) if len(sys.argv) > 3 else 3
frames = {}
for e in ev:
    f = frames.setdefault(e["args"]["frame"], {"K": [], "D": [], "F": []})
    f[e["cat"]].append((e["ts"], e["ts"] + e["dur"], e["args"]["bytes"]))
ids = sorted(i for i, f in frames.items() if f["K"] and f["D"] and f["F"] and sum(b for _, _, b in f["F"]) > 1e9)
if first is None:
    first = ids[len(ids) // 2 - 1] if len(ids) >= 3 else ids[0]
sel = [i for i in ids if i >= first][:count]
t0 = min(frames[i]["K"][0][0] for i in sel)
print("frame  K1 [ms]            D2H pieces [ms]                 file pieces [ms]                  bytes")
rows = []
for i in sel:
    f = frames[i]
    k = (f["K"][0][0] - t0, f["K"][0][1] - t0)
    d = (min(a for a, _, _ in f["D"]) - t0, max(b for _, b, _ in f["D"]) - t0)
    w = (min(a for a, _, _ in f["F"]) - t0, max(b for _, b, _ in f["F"]) - t0)
    rows.append((i, k, d, w))
    print(f"{i:5d}  {k[0]/1e3:8.2f} -{k[1]/1e3:8.2f}   {d[0]/1e3:8.2f} -{d[1]/1e3:8.2f} ({len(f['D'])} pcs)   "
          f"{w[0]/1e3:8.2f} -{w[1]/1e3:8.2f} ({len(f['F'])} pcs)   {sum(b for _, _, b in f['F'])}")
for (i, k, d, w), (j, k2, d2, w2) in zip(rows, rows[1:]):
    ov = max(0.0, min(w[1], d2[1]) - max(w[0], k2[0]))
    print(f"frame {j}: K1 starts {(w[1] - k2[0])/1e3:.1f} ms before frame {i}'s last file piece ends; "
          f"K1+D2H of {j} overlap the file stage of {i} for {ov/1e3:.1f} ms")
json.dump({"traceEvents": [e for e in ev if e["args"]["frame"] in sel]}, open(sys.argv[1].replace(".json", "_3frames.json"), "w"))
