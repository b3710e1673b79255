"""pgsd_sph_b200 -- B200-native PGSD: the per-frame particle-snapshot encode/write path and the
trajectory decode / particle-ID reorder path of krachdd/pgsd-sph behind its own API.

Like the reference package (/root/reference/pgsd/pgsd/__init__.py) the submodules are not imported
by default::

    import pgsd_sph_b200.fl      # drop-in for pgsd.fl     (C ABI: include/pgsd.h)
    import pgsd_sph_b200.hoomd   # drop-in for pgsd.hoomd  (+ reorder='id', device=True)

Everything below these modules lives in libpgsd_b200.so (CUDA kernels for sm_100a + host file
layer); importing a submodule fails loudly when the library has not been built.
"""
import signal
import sys

from . import version  # noqa: F401

# Same courtesy as the reference (__init__.py:19-26): let SIGTERM unwind so open files flush.
try:
    signal.signal(signal.SIGTERM, lambda n, f: sys.exit(1))
except ValueError:
    pass
