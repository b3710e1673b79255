"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the parity tests, smoke() and bench.py's cpu_baseline.

Nothing under pgsd_sph_b200/ imports this package.  Contents:
  _ref/            the UNMODIFIED reference libpgsd compiled in place (build_ref.sh) + ref_driver
  shim/            fork + shared-memory mpi.h so the reference's pgsd.c compiles without an MPI
  reader_oracle.py numpy restatement of the reference's pure-Python reader + HOOMD frame decode
  reorder_oracle.py the oracle definition of the ID reorder: numpy.argsort(kind='stable') + gather
  cast_oracle.py   numpy restatement of the host-side pack / dtype cast in front of write_chunk
"""
