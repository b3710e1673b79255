#!/bin/bash
# Build the UNMODIFIED reference libpgsd (C) in place from /root/reference into oracle/_ref/.
# TEST INFRASTRUCTURE ONLY: outputs are git-ignored binaries used as the parity checker and
# as bench.py's CPU "reference" arm.  No reference source is copied into this repository.
# /root/reference does not exist on the GPU box: there the prebuilt oracle/_ref/ files are used.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PGSD_REFERENCE_ROOT:-/root/reference}/pgsd/pgsd"
OUT="$HERE/_ref"
if [ ! -f "$REF/pgsd.c" ]; then
    echo "build_ref.sh: $REF/pgsd.c not found (expected on the GPU box); keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
CFLAGS="-O2 -g -fPIC -w -I$HERE/shim -I$REF"
gcc $CFLAGS -c "$REF/pgsd.c" -o "$OUT/pgsd_ref.o"
gcc $CFLAGS -c "$HERE/shim/mpishim.c" -o "$OUT/mpishim.o"
gcc $CFLAGS -c "$HERE/ref_driver.c" -o "$OUT/ref_driver.o"
gcc -o "$OUT/ref_driver" "$OUT/ref_driver.o" "$OUT/pgsd_ref.o" "$OUT/mpishim.o" -lpthread -lm
gcc -shared -o "$OUT/libpgsd_ref.so" "$OUT/pgsd_ref.o" "$OUT/mpishim.o" -lpthread
# the reference's own C++ benchmark (published numbers: CHANGELOG.md:172-194), unmodified
SCRIPTS="${PGSD_REFERENCE_ROOT:-/root/reference}/pgsd/scripts"
for b in benchmark-write benchmark-read; do
    if [ -f "$SCRIPTS/$b.cc" ]; then
        g++ -O2 -w -I"$HERE/shim" -I"$REF" -o "$OUT/$b" "$SCRIPTS/$b.cc" "$OUT/pgsd_ref.o" "$OUT/mpishim.o" -lpthread
    fi
done
# The reference's pure-Python reader (pypgsd.py + hoomd.py import with numpy only once `mpi4py` resolves to a stub:
# the MPI name is only used in commented-out code, hoomd.py:574-632).  Staged UNMODIFIED into the git-ignored
# oracle/_ref/pyref/ so that bench.py's --impl reference read leg times the reference's own decode on the GPU box,
# where /root/reference does not exist.  Not imported by anything under pgsd_sph_b200/.
PYREF="$OUT/pyref"
rm -rf "$PYREF"
mkdir -p "$PYREF/pgsd" "$PYREF/mpi4py"
for f in __init__.py version.py pypgsd.py hoomd.py; do
    cp "$REF/$f" "$PYREF/pgsd/$f"
done
printf 'class MPI:\n    pass\n' > "$PYREF/mpi4py/__init__.py"
echo "built $OUT/ref_driver and $OUT/libpgsd_ref.so from $REF/pgsd.c"
