"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the host-side pack + dtype cast the reference's
callers run in front of pgsd_write_chunk: numpy.ascontiguousarray (/root/reference/pgsd/pgsd/fl.pyx:571)
and ParticleData.validate's ascontiguousarray(dtype=...) + reshape([N, 3])
(/root/reference/pgsd/pgsd/hoomd.py:206-270).  K1 must produce these bytes.
"""
import numpy as np


def pack_soa(columns, dtype=None):
    """columns: M equally long 1-D arrays (any strides) -> contiguous (N, M) array of dtype."""
    a = np.stack([np.asarray(c) for c in columns], axis=1)
    with np.errstate(all='ignore'):
        return np.ascontiguousarray(a, dtype=dtype if dtype is not None else a.dtype)
