"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference and the compiled oracle/_ref/ref_driver):

    python tests/golden/make_golden.py

Outputs (committed; the GPU box has no /root/reference):
  v1_known_answer.gsd     the reference's own known-answer file, byte copy of
                          /root/reference/pgsd/pgsd/test/test_gsd_v1.gsd (test_fl.py:613-651:
                          127 int32 names x 5 frames, value = 13 * int(name))
  hoomd_p{1,2,3,8}.gsd    HOOMD-SPH frames (synth.make_frame, N=512, 3 frames, call sequence of
                          SURVEY.md Appendix A.4) written by the UNMODIFIED reference pgsd.c at
                          P shim ranks
  hoomd_p*.log            the reference's return codes / lookups for the same script
  reorder_p2.npz          hoomd_p2.gsd decoded by the reference's OWN Python reader
                          (pgsd.pypgsd + pgsd.hoomd imported from /root/reference with a stub
                          mpi4py) and put in particle-ID order with numpy.argsort(kind='stable')
  script_sha256.json      sha256 of the .gsd file and of the log the reference produces for the
                          seeded random op scripts of tests/randscript.py
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from opscript import Script, READONLY, run_reference, read_bytes  # noqa: E402
from randscript import random_script  # noqa: E402
from pgsd_sph_b200 import synth  # noqa: E402

REFERENCE = os.environ.get("PGSD_REFERENCE_ROOT", "/root/reference")
GOLDEN_N = 512
GOLDEN_FRAMES = 3
SCRIPT_SEEDS = list(range(48))


def hoomd_script(nframes=GOLDEN_FRAMES, n=GOLDEN_N):
    s = Script()
    s.create()
    for f in range(nframes):
        logs = [("log/value/kinetic_energy", np.array([0.5 * f + 1.25], dtype=np.float32)),
                ("log/value/potential_energy", np.array([-3.0 * f], dtype=np.float32))]
        s.hoomd_frame(synth.make_frame(n, f), synth.frame_scalars(n, f), logs)
    s.nframes()
    s.nnames()
    s.close()
    s.open(READONLY)
    for f in range(nframes):
        s.find(f, "particles/position")
        s.find(f, "log/particles/id")
        s.read(f, "particles/velocity", 0)
        s.read(f, "log/particles/id", 1)
    s.find(nframes, "particles/position")
    s.find(0, "no/such/chunk")
    s.match("log/")
    s.match("")
    s.close()
    return s


def seed_nprocs(seed):
    return [1, 2, 3, 8][seed % 4]


def main():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden.py needs the reference checkout at " + REFERENCE)
    shutil.copyfile(os.path.join(REFERENCE, "pgsd/pgsd/test/test_gsd_v1.gsd"), os.path.join(HERE, "v1_known_answer.gsd"))
    work = tempfile.mkdtemp(prefix="golden")
    for P in (1, 2, 3, 8):
        gsd, prefix = run_reference(hoomd_script(), work, f"hoomd_p{P}", P)
        shutil.copyfile(gsd, os.path.join(HERE, f"hoomd_p{P}.gsd"))
        shutil.copyfile(prefix + ".log", os.path.join(HERE, f"hoomd_p{P}.log"))

    # ---- reorder golden through the reference's own Python reader
    stub = os.path.join(work, "stub")
    os.makedirs(os.path.join(stub, "mpi4py"))
    with open(os.path.join(stub, "mpi4py", "__init__.py"), "w") as f:
        f.write("class MPI:\n    pass\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, os.path.join(REFERENCE, "pgsd"))
    sys.dont_write_bytecode = True
    import pgsd.hoomd as ref_hoomd  # the reference module
    import pgsd.pypgsd as ref_pypgsd
    out = {}
    with open(os.path.join(HERE, "hoomd_p2.gsd"), "rb") as fh:
        traj = ref_hoomd.HOOMDTrajectory(ref_pypgsd.PGSDFile(fh))
        assert len(traj) == GOLDEN_FRAMES
        for i in range(len(traj)):
            fr = traj[i]
            ids = fr.log['particles/id']
            o = np.argsort(ids, kind='stable')
            out[f"f{i}/N"] = np.array([fr.particles.N])
            out[f"f{i}/step"] = np.array([fr.configuration.step])
            out[f"f{i}/box"] = np.asarray(fr.configuration.box)
            out[f"f{i}/id"] = ids[o]
            for name in ('typeid', 'mass', 'body', 'position', 'velocity', 'slength', 'density', 'pressure',
                         'energy', 'auxiliary1', 'auxiliary2', 'auxiliary3', 'auxiliary4', 'image'):
                out[f"f{i}/{name}"] = np.asarray(getattr(fr.particles, name))[o]
            for k, v in fr.log.items():
                if not k.startswith('particles/'):
                    out[f"f{i}/log/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reorder_p2.npz"), **out)

    # ---- hashes of the reference's output for the seeded random scripts
    sums = {}
    for seed in SCRIPT_SEEDS:
        P = seed_nprocs(seed)
        gsd, prefix = run_reference(random_script(seed, P, lookups=(seed % 3 != 0)), work, f"s{seed}", P, timeout=30)
        sums[str(seed)] = {"nprocs": P, "gsd": hashlib.sha256(read_bytes(gsd)).hexdigest(),
                           "log": hashlib.sha256(read_bytes(prefix + ".log")).hexdigest(),
                           "bytes": os.path.getsize(gsd)}
    with open(os.path.join(HERE, "script_sha256.json"), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)
    shutil.rmtree(work)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
