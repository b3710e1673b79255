"""TEST INFRASTRUCTURE ONLY.  The oracle definition of the particle-ID reorder
(BASELINE.json north_star): the reference reader's arrays, then

    o = numpy.argsort(ids, kind='stable');  field[o]  for every per-particle field.

The reference itself never sorts (/root/reference/README.md:29-33); frames come back in file
(rank) order from hoomd.py:724-902.
"""
import numpy as np


def reorder(ids, fields):
    """-> (sorted ids, {name: field[o]}, o)."""
    ids = np.asarray(ids)
    o = np.argsort(ids, kind='stable')
    return ids[o], {k: np.asarray(v)[o] for k, v in fields.items()}, o


def reorder_frame(decoded, id_name='log/particles/id'):
    """Apply the reorder to a dict from reader_oracle.decode_particles(): every array with N rows."""
    n = decoded['N']
    ids = decoded[id_name]
    fields = {k: v for k, v in decoded.items()
              if k not in ('N', id_name) and hasattr(v, 'shape') and len(v) == n
              and (not k.startswith('log/') or k.startswith('log/particles/'))}
    sorted_ids, out, o = reorder(ids, fields)
    res = dict(decoded)
    res.update(out)
    res[id_name] = sorted_ids
    return res


def distributed_ownership(n_global, nranks):
    """Which ids each rank holds after the distributed reorder of one frame (SURVEY.md section 8e, row 3:
    "destination GPU = id / ceil(N/G)", here rounded up to whole id buckets so that ownership is decided by key
    bits): rank r owns [r * S, (r + 1) * S) with S = ceil(ceil(N / C) / ranks) * C and C = 1024 ids per bucket
    (2048 / 4096 when more than 32 Mi / 64 Mi rows make more than 32768 buckets).  -> (S, C)."""
    n = int(n_global)
    bits = max((n - 1).bit_length(), 0)
    L = 10
    while bits - L > 15:
        L += 1
    C = 1 << L
    buckets = -(-n // C)
    return -(-buckets // int(nranks)) * C, C


def reorder_distributed(ids, fields, nranks):
    """Expected result of the distributed reorder, rank by rank: [(first id, sorted ids, {name: rows})]."""
    sorted_ids, out, _ = reorder(ids, fields)
    S, _ = distributed_ownership(len(sorted_ids), nranks)
    shares = []
    for r in range(nranks):
        lo, hi = np.searchsorted(sorted_ids, [r * S, (r + 1) * S])
        shares.append((r * S, sorted_ids[lo:hi], {k: v[lo:hi] for k, v in out.items()}))
    return shares
