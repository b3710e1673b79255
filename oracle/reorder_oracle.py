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
