"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's read path.

Restates, for checking only:
  * pypgsd.PGSDFile  (/root/reference/pgsd/pgsd/pypgsd.py:104-433): header 'QQQQQII64s64s80s'
    (:43-57), index entries 'QQqIHBB' (:59-62), namelist split on NUL (:140-151), index scan that
    stops at location == 0 (:153-175), frame bisect + backwards id scan (:226-256), read_chunk
    (:284-347).
  * HOOMDTrajectory._read_frame (/root/reference/pgsd/pgsd/hoomd.py:724-902) for the per-particle
    fields: chunk of frame i, else frame 0's array when N matches, else the schema default.
Pinned by tests/test_oracle.py against the reference's own known-answer file
(pgsd/pgsd/test/test_gsd_v1.gsd, test_fl.py:613-651) and against frames decoded by the reference's
pypgsd + hoomd modules themselves (tests/golden/make_golden.py).
"""
import numpy as np

HEADER_DTYPE = np.dtype([('magic', '<u8'), ('index_location', '<u8'), ('index_allocated_entries', '<u8'),
                         ('namelist_location', '<u8'), ('namelist_allocated_entries', '<u8'),
                         ('schema_version', '<u4'), ('pgsd_version', '<u4'), ('application', 'S64'),
                         ('schema', 'S64'), ('reserved', 'S80')])
ENTRY_DTYPE = np.dtype([('frame', '<u8'), ('N', '<u8'), ('location', '<i8'), ('M', '<u4'), ('id', '<u2'),
                        ('type', 'u1'), ('flags', 'u1')])
assert HEADER_DTYPE.itemsize == 256 and ENTRY_DTYPE.itemsize == 32
TYPES = {1: np.dtype('uint8'), 2: np.dtype('uint16'), 3: np.dtype('uint32'), 4: np.dtype('uint64'),
         5: np.dtype('int8'), 6: np.dtype('int16'), 7: np.dtype('int32'), 8: np.dtype('int64'),
         9: np.dtype('float32'), 10: np.dtype('float64')}
MAGIC = 0x65DF65DF65DF65DF


class OracleFile:
    """Read-only PGSD/GSD reader with the duck-typed file interface pgsd.hoomd expects."""

    mode = 'r'

    def __init__(self, path):
        self.name = path
        with open(path, 'rb') as f:
            self._raw = f.read()
        if len(self._raw) < 256:
            raise IOError("short file")
        h = np.frombuffer(self._raw, dtype=HEADER_DTYPE, count=1)[0]
        self.header = h
        if int(h['magic']) != MAGIC:
            raise RuntimeError("Not a PGSD file: " + path)
        v = int(h['pgsd_version'])
        if (v < (1 << 16) and v != 3) or v >= (3 << 16):
            raise RuntimeError("Unsupported PGSD file version: " + path)
        # namelist: names in order of appearance = id (pypgsd.py:140-151)
        nl = self._raw[int(h['namelist_location']):int(h['namelist_location']) + 64 * int(h['namelist_allocated_entries'])]
        self.names = {}
        for nm in nl.split(b'\x00'):
            s = nm.decode('utf-8')
            if len(s) != 0:
                self.names[s] = len(self.names)
        # index: entries up to the first location == 0 (pypgsd.py:153-175)
        n_alloc = int(h['index_allocated_entries'])
        idx = np.frombuffer(self._raw, dtype=ENTRY_DTYPE, count=n_alloc, offset=int(h['index_location']))
        zero = np.nonzero(idx['location'] == 0)[0]
        used = int(zero[0]) if len(zero) else n_alloc
        idx = idx[:used]
        ok = (np.isin(idx['type'], list(TYPES)) & (idx['M'] != 0) & (idx['frame'] < n_alloc)
              & (idx['id'] < len(self.names)) & (idx['flags'] == 0))
        if not ok.all() or (used > 1 and (np.diff(idx['frame'].astype(np.int64)) < 0).any()):
            raise RuntimeError("Corrupt PGSD file: " + path)
        self.index = idx

    @property
    def nframes(self):
        return 0 if len(self.index) == 0 else int(self.index[-1]['frame']) + 1

    @property
    def schema(self):
        return bytes(self.header['schema']).rstrip(b'\x00').decode('utf-8')

    @property
    def application(self):
        return bytes(self.header['application']).rstrip(b'\x00').decode('utf-8')

    @property
    def schema_version(self):
        v = int(self.header['schema_version'])
        return (v >> 16, v & 0xffff)

    @property
    def pgsd_version(self):
        v = int(self.header['pgsd_version'])
        return (v >> 16, v & 0xffff)

    def _find_chunk(self, frame, name):
        # bisect to the last entry of the frame, scan backwards for the id (pypgsd.py:226-256)
        if name not in self.names:
            return None
        mid = self.names[name]
        L, R = 0, len(self.index)
        while R - L > 1:
            m = (L + R) // 2
            if frame < self.index[m]['frame']:
                R = m
            else:
                L = m
        cur = L
        while cur >= 0 and cur < len(self.index) and self.index[cur]['frame'] == frame:
            if self.index[cur]['id'] == mid:
                return self.index[cur]
            cur -= 1
        return None

    def chunk_exists(self, frame, name, write_all=False):
        return self._find_chunk(frame, name) is not None

    def read_chunk(self, frame, name, offset=0, r_all=False):
        c = self._find_chunk(frame, name)
        if c is None:
            raise KeyError(f"frame {frame} / chunk {name} not found in: {self.name}")
        dt = TYPES[int(c['type'])]
        size = int(c['N']) * int(c['M']) * dt.itemsize
        if int(c['location']) == 0:
            raise RuntimeError("Corrupt chunk")
        if size == 0:
            return np.array([], dtype=dt)
        raw = self._raw[int(c['location']):int(c['location']) + size]
        if len(raw) != size:
            raise IOError
        a = np.frombuffer(raw, dtype=dt)
        return a if int(c['M']) == 1 else a.reshape([int(c['N']), int(c['M'])])

    def find_matching_chunk_names(self, match, write_all=False):
        return [k for k in self.names if k.startswith(match)]

    def close(self):
        pass


# per-particle fields of the HOOMD-SPH schema and their defaults (hoomd.py:149-184)
PARTICLE_DEFAULTS = {
    'typeid': np.uint32(0), 'mass': np.float32(1.0), 'body': np.int32(-1),
    'position': np.array([0, 0, 0], dtype=np.float32), 'velocity': np.array([0, 0, 0], dtype=np.float32),
    'slength': np.float32(1.0), 'density': np.float32(0.0), 'pressure': np.float32(0.0),
    'energy': np.float32(0.0), 'auxiliary1': np.array([0, 0, 0], dtype=np.float32),
    'auxiliary2': np.array([0, 0, 0], dtype=np.float32), 'auxiliary3': np.array([0, 0, 0], dtype=np.float32),
    'auxiliary4': np.array([0, 0, 0], dtype=np.float32), 'image': np.array([0, 0, 0], dtype=np.int32),
}


def decode_particles(f, idx, _frame0=None):
    """Per-particle arrays of frame idx in FILE order: {'N', fields..., 'log/<name>'...}
    with the frame-0 / default fallback of hoomd.py:852-893."""
    if idx != 0 and _frame0 is None:
        _frame0 = decode_particles(f, 0)
    out = {}
    if f.chunk_exists(idx, 'particles/N'):
        out['N'] = int(f.read_chunk(idx, 'particles/N')[0])
    else:
        out['N'] = _frame0['N'] if _frame0 is not None else 0
    for name, dv in PARTICLE_DEFAULTS.items():
        if f.chunk_exists(idx, 'particles/' + name):
            out[name] = f.read_chunk(idx, 'particles/' + name)
        elif _frame0 is not None and _frame0['N'] == out['N']:
            out[name] = _frame0[name]
        else:
            tmp = np.array([dv])
            a = np.empty(shape=[out['N']] + list(tmp.shape[1:]), dtype=tmp.dtype)
            a[:] = tmp
            out[name] = a
    for log in f.find_matching_chunk_names('log/'):
        if f.chunk_exists(idx, log):
            out[log] = f.read_chunk(idx, log)
        elif _frame0 is not None and log in _frame0:
            out[log] = _frame0[log]
    return out
