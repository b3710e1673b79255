"""Stress one seeded op-script through the device write path; report differing byte ranges vs the golden hash."""
import hashlib, json, os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import opscript
from golden.make_golden import seed_nprocs
from randscript import random_script
golden = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
sums = json.load(open(os.path.join(golden, "script_sha256.json")))
seeds = [int(x) for x in sys.argv[1].split(",")]
reps = int(sys.argv[2])
for seed in seeds:
    P = seed_nprocs(seed)
    good, bads = None, []
    for rep in range(reps):
        d = tempfile.mkdtemp()
        gsd, prefix = opscript.run_replay(random_script(seed, P, lookups=(seed % 3 != 0)), d, f"s{seed}", P, device=True, soa=(seed % 2 == 0), timeout=300)
        b = opscript.read_bytes(gsd)
        if hashlib.sha256(b).hexdigest() == sums[str(seed)]["gsd"]:
            good = b
        else:
            bads.append(b)
    print("seed", seed, "P", P, "bad", len(bads), "of", reps, flush=True)
    for b in bads:
        if good is None:
            print("  no good run to compare"); break
        print("  sizes", len(good), len(b))
        m = min(len(good), len(b)); i = 0; runs = []
        while i < m and len(runs) < 12:
            if good[i] != b[i]:
                j = i
                while j < m and good[j] != b[j]: j += 1
                runs.append((i, j - i, good[i:min(j, i + 16)].hex(), b[i:min(j, i + 16)].hex())); i = j
            else:
                i += 1
        for r in runs: print("  diff at", r)
