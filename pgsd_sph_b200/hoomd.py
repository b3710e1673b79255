"""Read (and write) HOOMD-SPH schema PGSD files -- drop-in for the reference's ``pgsd.hoomd``.

Same classes and call surface as /root/reference/pgsd/pgsd/hoomd.py: :py:func:`open`,
:py:func:`read_log`, :py:class:`HOOMDTrajectory`, :py:class:`Frame`, :py:class:`ParticleData`
(with the SPH fields slength/density/pressure/energy/auxiliary1-4, hoomd.py:111-270),
:py:class:`ConfigurationData`, :py:class:`ConstraintData`, :py:class:`BondData`.  Frame decoding
follows ``_read_frame`` (hoomd.py:724-902): a chunk missing from frame *i* falls back to frame 0's
value when the particle count matches, else to the schema default.

New, on top of the reference surface:

* ``HOOMDTrajectory(file, reorder='id')`` returns every frame in particle-ID order: the decoded
  per-particle arrays (file order = rank order, never sorted by the reference, README.md:29) are
  reordered on the GPU by a stable LSD radix sort of ``log/particles/id`` (K4) and one
  permutation gather of all fields (K5).  Bit-exact with the reference reader followed by
  ``o = numpy.argsort(id, kind='stable'); field[o]``.
* ``device=True`` keeps the per-particle arrays in device memory
  (:py:class:`~pgsd_sph_b200.devmem.DeviceArray`, ``__cuda_array_interface__``).
* :py:meth:`HOOMDTrajectory.append` works (the reference's raises, hoomd.py:568): it writes the
  frame with the call convention of the reference's disabled writer (hoomd.py:597-632) --
  per-particle chunks ``write_all=True`` at this rank's row offset, scalars ``write_all=False``.
"""
import ctypes as C
import json
import logging
import warnings
from collections import OrderedDict

import numpy

from . import _lib, fl
from .devmem import DeviceArray, is_device_array
from .version import __version__

logger = logging.getLogger('pgsd.hoomd')


class ConfigurationData(object):
    """Store configuration data: step, dimensions, box (ref: hoomd.py:45-108)."""

    _default_value = OrderedDict()
    _default_value['step'] = numpy.uint64(0)
    _default_value['dimensions'] = numpy.uint8(3)
    _default_value['box'] = numpy.array([1, 1, 1, 0, 0, 0], dtype=numpy.float32)

    def __init__(self):
        self.step = None
        self.dimensions = None
        self._box = None

    @property
    def box(self):
        """[lx, ly, lz, xy, xz, yz]; setting it also sets ``dimensions`` when that is None."""
        return self._box

    @box.setter
    def box(self, box):
        self._box = box
        try:
            Lz = box[2]
        except TypeError:
            return
        else:
            if self.dimensions is None:
                self.dimensions = 2 if Lz == 0 else 3

    def validate(self):
        """Convert the box to a contiguous float32 array of 6 values."""
        logger.debug('Validating ConfigurationData')
        if self.box is not None:
            self.box = numpy.ascontiguousarray(self.box, dtype=numpy.float32)
            self.box = self.box.reshape([6])


# per-particle fields: name -> (dtype, columns, default); order = file order of the schema
_PARTICLE_FIELDS = OrderedDict([
    ('typeid', (numpy.uint32, 1, 0)),
    ('mass', (numpy.float32, 1, 1.0)),
    ('body', (numpy.int32, 1, -1)),
    ('position', (numpy.float32, 3, 0)),
    ('velocity', (numpy.float32, 3, 0)),
    ('slength', (numpy.float32, 1, 1.0)),
    ('density', (numpy.float32, 1, 0.0)),
    ('pressure', (numpy.float32, 1, 0.0)),
    ('energy', (numpy.float32, 1, 0.0)),
    ('auxiliary1', (numpy.float32, 3, 0)),
    ('auxiliary2', (numpy.float32, 3, 0)),
    ('auxiliary3', (numpy.float32, 3, 0)),
    ('auxiliary4', (numpy.float32, 3, 0)),
    ('image', (numpy.int32, 3, 0)),
])


class ParticleData(object):
    """Store particle data chunks, SPH fields included (ref: hoomd.py:111-270)."""

    _default_value = OrderedDict()
    _default_value['N'] = numpy.uint32(0)
    _default_value['types'] = ['A']
    for _name, (_dt, _m, _dv) in _PARTICLE_FIELDS.items():
        _default_value[_name] = _dt(_dv) if _m == 1 else numpy.array([_dv] * _m, dtype=_dt)
    _default_value['type_shapes'] = [{}]
    del _name, _dt, _m, _dv

    def __init__(self):
        self.N = 0
        self.types = None
        self.type_shapes = None
        for name in _PARTICLE_FIELDS:
            setattr(self, name, None)

    def validate(self):
        """Contiguous arrays of the schema's dtype and shape (N,) / (N, 3); CUDA arrays pass through
        (the dtype cast and packing then happen on the device when the frame is written)."""
        logger.debug('Validating ParticleData')
        for name, (dt, m, _) in _PARTICLE_FIELDS.items():
            v = getattr(self, name)
            if v is None or is_device_array(v):
                continue
            v = numpy.ascontiguousarray(v, dtype=dt)
            setattr(self, name, v.reshape([self.N] if m == 1 else [self.N, m]))
        if self.types is not None and not len(set(self.types)) == len(self.types):
            raise ValueError("Type names must be unique.")


class BondData(object):
    """Store bond data chunks (ref: hoomd.py:273-351); kept for API compatibility."""

    def __init__(self, M):
        self.M = M
        self.N = 0
        self.types = None
        self.typeid = None
        self.group = None
        self._default_value = OrderedDict()
        self._default_value['N'] = numpy.uint32(0)
        self._default_value['types'] = []
        self._default_value['typeid'] = numpy.uint32(0)
        self._default_value['group'] = numpy.array([0] * M, dtype=numpy.int32)

    def validate(self):
        logger.debug('Validating BondData')
        if self.typeid is not None:
            self.typeid = numpy.ascontiguousarray(self.typeid, dtype=numpy.uint32).reshape([self.N])
        if self.group is not None:
            self.group = numpy.ascontiguousarray(self.group, dtype=numpy.int32).reshape([self.N, self.M])
        if self.types is not None and not len(set(self.types)) == len(self.types):
            raise ValueError("Type names must be unique.")


class ConstraintData(object):
    """Store constraint data chunks (ref: hoomd.py:354-421)."""

    def __init__(self):
        self.M = 2
        self.N = 0
        self.value = None
        self.group = None
        self._default_value = OrderedDict()
        self._default_value['N'] = numpy.uint32(0)
        self._default_value['value'] = numpy.float32(0)
        self._default_value['group'] = numpy.array([0] * self.M, dtype=numpy.int32)

    def validate(self):
        logger.debug('Validating ConstraintData')
        if self.value is not None:
            self.value = numpy.ascontiguousarray(self.value, dtype=numpy.float32).reshape([self.N])
        if self.group is not None:
            self.group = numpy.ascontiguousarray(self.group, dtype=numpy.int32).reshape([self.N, self.M])


class Frame(object):
    """System state at one point in time (ref: hoomd.py:424-467)."""

    def __init__(self, num_procs=0):
        self.configuration = ConfigurationData()
        self.particles = ParticleData()
        self.constraints = ConstraintData()
        self.state = {}
        self.log = {}
        self.num_procs = num_procs
        self.part_dist = None  # per-rank particle counts when written by several ranks

    def validate(self):
        self.configuration.validate()
        self.particles.validate()
        self.constraints.validate()


class _HOOMDTrajectoryIterable(object):
    """Iterable over a HOOMDTrajectory object (ref: hoomd.py:470-490)."""

    def __init__(self, trajectory, indices):
        self._trajectory = trajectory
        self._indices = indices
        self._indices_iterator = iter(indices)

    def __next__(self):
        return self._trajectory[next(self._indices_iterator)]

    next = __next__

    def __iter__(self):
        return type(self)(self._trajectory, self._indices)

    def __len__(self):
        return len(self._indices)


class _HOOMDTrajectoryView(object):
    """A view of a HOOMDTrajectory object (ref: hoomd.py:493-516)."""

    def __init__(self, trajectory, indices):
        self._trajectory = trajectory
        self._indices = indices

    def __iter__(self):
        return _HOOMDTrajectoryIterable(self._trajectory, self._indices)

    def __len__(self):
        return len(self._indices)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return type(self)(self._trajectory, self._indices[key])
        return self._trajectory[self._indices[key]]


ID_CHUNK = 'log/particles/id'  # particle id has no schema slot; SURVEY.md section 7 decision


def _packed_device_field(name, obj, n):
    """(ptr, shape, dtype, keepalive) of a device array that the reorder kernels can read as n packed rows."""
    from .devmem import as_device_view
    ptr, shape, dt, strides, keep = as_device_view(obj)
    rows = int(shape[0]) if len(shape) else 1
    if rows != n:
        raise ValueError(f"field {name} has {rows} rows, expected {n}")
    if strides is not None:
        want, acc = [], dt.itemsize
        for d in reversed(shape):
            want.append(acc)
            acc *= int(d)
        if tuple(strides) != tuple(reversed(want)) and n > 1:
            raise ValueError(f"field {name} is not C-contiguous on the device (strides {tuple(strides)}); "
                             "copy it to a packed array first")
    return ptr, shape, dt, keep


def reorder_by_id(ids, arrays, device=False):
    """Reorder per-particle arrays into particle-ID order on the GPU (K4 + K5).

    ``ids`` is a uint32 array of N keys; ``arrays`` a dict of arrays with N rows each.  All in host
    memory (numpy in -> numpy out, copies through pinned buffers inside the call) or all in device
    memory (DeviceArray / CUDA array in -> DeviceArray out).  Returns (sorted ids, reordered dict).
    Equals ``o = numpy.argsort(ids, kind='stable'); {k: v[o]}`` bit for bit.
    """
    lib = _lib.load()
    names = list(arrays.keys())
    if device:
        from .devmem import as_device_view
        kptr, kshape, kdt, _, keep = as_device_view(ids)
        n = int(kshape[0]) if len(kshape) else 1
        if kdt != numpy.dtype(numpy.uint32):
            raise ValueError("particle ids must be uint32")
        sorted_ids = DeviceArray((n,), numpy.uint32)
        outs, fields, keeps = {}, (_lib.Field * max(len(names), 1))(), [keep]
        for i, k in enumerate(names):
            ptr, shape, dt, kp = _packed_device_field(k, arrays[k], n)
            keeps.append(kp)
            row = dt.itemsize * (int(numpy.prod(shape[1:])) if len(shape) > 1 else 1)
            outs[k] = DeviceArray(shape, dt)
            fields[i] = _lib.Field(ptr, outs[k].ptr, row)
        _lib.check(lib.pgsd_b200_reorder_device(n, kptr, sorted_ids.ptr, None, len(names), fields, None),
                   "pgsd_b200_reorder_device")
        _lib.check(lib.pgsd_b200_synchronize(), "synchronize")
        return sorted_ids, outs
    ids = numpy.ascontiguousarray(ids, dtype=numpy.uint32)
    n = ids.shape[0]
    sorted_ids = numpy.empty_like(ids)
    outs, fields, srcs = {}, (_lib.Field * max(len(names), 1))(), []
    for i, k in enumerate(names):
        a = numpy.ascontiguousarray(arrays[k])
        if a.shape[0] != n:
            raise ValueError(f"field {k} has {a.shape[0]} rows, expected {n}")
        srcs.append(a)
        outs[k] = numpy.empty_like(a)
        row = a.dtype.itemsize * (int(numpy.prod(a.shape[1:])) if a.ndim > 1 else 1)
        fields[i] = _lib.Field(a.ctypes.data, outs[k].ctypes.data, row)
    if n:
        _lib.check(lib.pgsd_b200_reorder_host(n, ids.ctypes.data, sorted_ids.ctypes.data, None, len(names), fields),
                   "pgsd_b200_reorder_host")
    return sorted_ids, outs


def reorder_by_id_distributed(ids, arrays):
    """Particle-ID order for ONE frame whose rows are partitioned over the ranks (one process per GPU; every rank
    calls this with the partition it holds, e.g. what ``read_chunk(..., r_all=True, device=True)`` returned).

    ``ids``: uint32 device array of this rank's keys; ``arrays``: dict of device arrays with as many rows.
    Returns ``(first_id, sorted ids, dict)``: this rank's share of the ID-ordered frame -- the rows whose ids
    lie in ``[first_id, first_id + S)`` (``S`` from ``pgsd_b200_reorder_distributed_plan``), as DeviceArrays
    trimmed to the rows actually owned.  Concatenated over the ranks in rank order this equals
    ``o = numpy.argsort(all_ids, kind='stable'); {k: all_v[o]}`` bit for bit.  Raises ``ValueError`` (on every
    rank) when the ids are not unique or not dense enough; gather the frame to one GPU and use
    :py:func:`reorder_by_id` then.  Collective: the records travel GPU to GPU inside the scatter kernel.
    """
    import ctypes as C
    from .devmem import as_device_view
    lib = _lib.load()
    names = list(arrays.keys())
    kptr, kshape, kdt, _, keep = as_device_view(ids)
    n = int(kshape[0]) if len(kshape) else 1
    if kdt != numpy.dtype(numpy.uint32):
        raise ValueError("particle ids must be uint32")
    rank, nranks = lib.pgsd_b200_comm_rank(), lib.pgsd_b200_comm_size()
    n_global, row_start = C.c_uint64(), C.c_uint64()
    _lib.check(lib.pgsd_b200_partition(n, C.byref(n_global), C.byref(row_start)), "pgsd_b200_partition")
    first, cap = C.c_uint64(), C.c_uint64()
    if n_global.value == 0:
        return 0, DeviceArray((0,), numpy.uint32), {k: DeviceArray((0,) + tuple(as_device_view(arrays[k])[1][1:]),
                                                                    as_device_view(arrays[k])[2]) for k in names}
    _lib.check(lib.pgsd_b200_reorder_distributed_plan(n_global.value, nranks, rank, C.byref(first), C.byref(cap)),
               "pgsd_b200_reorder_distributed_plan")
    cap = int(cap.value)
    sorted_ids = DeviceArray((cap,), numpy.uint32)
    outs, fields, keeps = {}, (_lib.Field * max(len(names), 1))(), [keep]
    bad = None
    for i, k in enumerate(names):
        try:
            ptr, shape, dt, kp = _packed_device_field(k, arrays[k], n)
        except ValueError as e:   # the call below is collective: a rank must not leave before it
            bad = bad or e
            ptr, shape, dt, _, kp = as_device_view(arrays[k])
        keeps.append(kp)
        row = dt.itemsize * (int(numpy.prod(shape[1:])) if len(shape) > 1 else 1)
        outs[k] = DeviceArray((cap,) + tuple(shape[1:]), dt)
        fields[i] = _lib.Field(ptr, outs[k].ptr, row)
    n_out, id_first = C.c_uint64(), C.c_uint64()
    if bad is not None:
        # take part in the collective with an empty partition flagged as invalid: every rank gets the error
        lib.pgsd_b200_reorder_distributed(0xffffffffffffffff, kptr, cap, C.byref(n_out), C.byref(id_first), sorted_ids.ptr,
                                          len(names), fields, None)
        raise bad
    rc = lib.pgsd_b200_reorder_distributed(n, kptr, cap, C.byref(n_out), C.byref(id_first), sorted_ids.ptr, len(names),
                                           fields, None)
    if rc == 1:
        raise ValueError("reorder_by_id_distributed: particle ids are not unique or not dense in [0, N)")
    _lib.check(rc, "pgsd_b200_reorder_distributed")
    k_rows = int(n_out.value)

    def trim(a):
        return DeviceArray((k_rows,) + a.shape[1:], a.dtype, ptr=a.ptr, owner=a)
    return int(id_first.value), trim(sorted_ids), {k: trim(v) for k, v in outs.items()}


#: per-particle chunks of the SPH configs and their row widths
SPH_FIELDS = {'position': 3, 'velocity': 3, 'typeid': 1, 'density': 1, 'pressure': 1}


def read_frame_distributed(file, frame, fields=None):
    """Config 3 read-back, end to end: one frame, every rank reads ITS row slice of each per-particle chunk from
    the file straight into device memory (``read_chunk(..., r_all=True, device=True)``, the row-sliced read of
    pgsd.c:2497-2508 with the split rule of benchmark-read.cc:64-76) and the frame is put into particle-ID order
    across the GPUs by :py:func:`reorder_by_id_distributed`.

    ``file``: a :py:class:`pgsd_sph_b200.fl.PGSDFile` opened for reading on every rank; ``fields``: dict
    ``name -> M`` of ``particles/<name>`` chunks (default :py:data:`SPH_FIELDS`).  Returns ``(first_id, ids,
    arrays)``: this rank's share of the ID-ordered frame as DeviceArrays.  Collective.
    """
    from . import synth
    lib = _lib.load()
    fields = dict(SPH_FIELDS if fields is None else fields)
    n = int(file.read_chunk(frame=frame, name='particles/N')[0])
    rank, nranks = lib.pgsd_b200_comm_rank(), lib.pgsd_b200_comm_size()
    rows = synth.split_rows(n, nranks)
    start, mine = synth.row_starts(rows)[rank], rows[rank]
    ids = file.read_chunk(frame=frame, name=ID_CHUNK, N=mine, M=1, offset=start, r_all=True, device=True)
    arrays = {k: file.read_chunk(frame=frame, name='particles/' + k, N=mine, M=m, offset=start, r_all=True, device=True)
              for k, m in fields.items()}
    return reorder_by_id_distributed(ids.reshape(mine), {k: (v if fields[k] > 1 else v.reshape(mine)) for k, v in arrays.items()})


class HOOMDTrajectory(object):
    """Read and write hoomd pgsd files (ref: hoomd.py:519-941).

    Args:
        file (:py:class:`pgsd_sph_b200.fl.PGSDFile`): File to access.
        reorder (None or 'id'): return frames in particle-ID order (GPU radix sort + gather).
        device (bool): keep per-particle arrays on the GPU.
        prefetch (bool or None): read frame i+1's per-particle chunks into device memory on a helper
            thread while frame i is processed (default: on for device / reorder reads of
            read-only single-rank files).
    """

    def __init__(self, file, reorder=None, device=False, prefetch=None):
        if file.mode == 'ab':
            raise ValueError('Append mode not yet supported')
        if reorder not in (None, 'id'):
            raise ValueError("reorder must be None or 'id'")
        self._file = file
        self._initial_frame = None
        self._reorder = reorder
        self._device = bool(device)
        # reorder='id': per-particle chunks go file -> pinned -> HBM directly (no host array in between),
        # are sorted + gathered there, and come back with one D2H copy per field
        self._read_device = self._device or reorder == 'id'
        # sequential access (for frame in traj / traj[i], traj[i+1], ...): while frame i is sorted and
        # copied back, a helper thread already pulls the per-particle chunks of frame i+1 from the file
        # into device memory through a second read-only handle.  Single-rank read-only files only
        # (every library call on a multi-rank communicator is collective).
        if prefetch is None:
            prefetch = self._read_device
        self._prefetch_on = bool(prefetch) and self._read_device and file.mode == 'r' \
            and _lib.load().pgsd_b200_comm_size() == 1
        self._pf = None       # (frame index, thread, {chunk name: DeviceArray}, [error])
        self._pf_file = None
        logger.info('opening HOOMDTrajectory: ' + str(self.file))
        if self.file.schema != 'hoomd':
            raise RuntimeError('PGSD file is not a hoomd schema file: ' + str(self.file))
        version = self.file.schema_version
        if not (version < (2, 0) and version >= (1, 0)):
            raise RuntimeError('Incompatible hoomd schema version ' + str(version) + ' in: ' + str(self.file))
        logger.info('found ' + str(len(self)) + ' frames')

    @property
    def file(self):
        """The underlying file handle."""
        return self._file

    def __len__(self):
        return self.file.nframes

    # ------------------------------------------------------------------ write
    def append(self, frame):
        """Append a frame: every non-None field is written (PGSD rewrites all fields every frame,
        README.md:32-33).  ``frame.part_dist`` (per-rank particle counts) places this rank's rows;
        when it is None the library computes the placement (K2, ``offset='auto'``)."""
        logger.debug('Appending frame to hoomd trajectory: ' + str(self.file))
        frame.validate()
        lib = _lib.load()
        rank, nprocs = lib.pgsd_b200_comm_rank(), lib.pgsd_b200_comm_size()
        f = self.file
        offset = frame.part_dist if frame.part_dist is not None else ('auto' if nprocs > 1 else None)
        cfg = frame.configuration
        if cfg.step is not None:
            f.write_chunk('configuration/step', numpy.array([cfg.step], dtype=numpy.uint64), write_all=False)
        if cfg.dimensions is not None:
            f.write_chunk('configuration/dimensions', numpy.array([cfg.dimensions], dtype=numpy.uint8), write_all=False)
        if cfg.box is not None:
            f.write_chunk('configuration/box', cfg.box, write_all=False)
        p = frame.particles
        n_global = int(numpy.sum(frame.part_dist)) if frame.part_dist is not None else None
        if n_global is None:
            n_local = numpy.array([int(p.N)], dtype=numpy.uint64)
            n_tot, n_start = C.c_uint64(), C.c_uint64()
            _lib.check(lib.pgsd_b200_partition(int(n_local[0]), C.byref(n_tot), C.byref(n_start)), "partition")
            n_global = n_tot.value
        f.write_chunk('particles/N', numpy.array([n_global], dtype=numpy.uint32), write_all=False)
        for name, strings in (('types', p.types), ('type_shapes', p.type_shapes)):
            if strings is None:
                continue
            if name == 'type_shapes':
                strings = [json.dumps(d) for d in strings]
            wid = max(len(w) for w in strings) + 1
            b = numpy.array(strings, dtype=numpy.dtype((bytes, wid)))
            f.write_chunk('particles/' + name, b.view(dtype=numpy.int8).reshape(len(b), wid), write_all=False)
        for name in _PARTICLE_FIELDS:
            data = getattr(p, name)
            if data is not None:
                f.write_chunk('particles/' + name, data, offset, rank, True)
        c = frame.constraints
        if c.N:
            f.write_chunk('constraints/N', numpy.array([c.N], dtype=numpy.uint32), write_all=False)
            for name in ('value', 'group'):
                if getattr(c, name) is not None:
                    f.write_chunk('constraints/' + name, getattr(c, name), write_all=False)
        for log, data in frame.log.items():
            per_particle = log.startswith('particles/')
            if per_particle:
                f.write_chunk('log/' + log, data, offset, rank, True)
            else:
                f.write_chunk('log/' + log, data, write_all=False)
        f.end_frame()

    def extend(self, iterable):
        for item in iterable:
            self.append(item)

    def close(self):
        """Close the file."""
        self._drop_prefetch()
        if self._pf_file is not None:
            self._pf_file.close()
            self._pf_file = None
        self.file.close()
        self._initial_frame = None

    def flush(self):
        """Flush all buffered frames to the file."""
        self._file.flush()

    # ------------------------------------------------------------------ read
    def read_frame(self, idx):
        warnings.warn("Deprecated, trajectory[idx]", DeprecationWarning)
        return self._read_frame(idx)

    _BIG_PREFIXES = ('particles/', 'log/particles/')

    def _start_prefetch(self, idx):
        if not self._prefetch_on or idx >= len(self) or idx < 0:
            return
        if self._pf is not None and self._pf[0] == idx:
            return
        self._drop_prefetch()
        import threading
        if self._pf_file is None:
            self._pf_file = fl.open(self.file.name, 'r')
        names = [n for n in self._per_particle_names()]
        out, err = {}, []

        def work():
            try:
                for n in names:
                    if self._pf_file.chunk_exists(idx, n):
                        out[n] = self._pf_file.read_chunk(idx, n, device=True)
            except Exception as e:  # surfaced on the consuming side as a plain synchronous read
                err.append(e)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        self._pf = (idx, t, out, err)

    def _drop_prefetch(self):
        if self._pf is not None:
            self._pf[1].join()
            for v in self._pf[2].values():
                v.free()
            self._pf = None

    def _take_prefetched(self, idx):
        """Chunks of frame idx read ahead of time ({} if none)."""
        if self._pf is None:
            return {}
        if self._pf[0] != idx:
            self._drop_prefetch()
            return {}
        _, t, out, err = self._pf
        t.join()
        self._pf = None
        if err:
            for v in out.values():
                v.free()
            return {}
        return out

    def _per_particle_names(self):
        names = ['particles/' + n for n in ParticleData._default_value if n not in ('N', 'types', 'type_shapes')]
        names += [n for n in self.file.find_matching_chunk_names('log/particles/', False)]
        return names

    def _read_big(self, idx, name, ahead):
        """A per-particle chunk of frame idx: the prefetched device copy if there is one."""
        v = ahead.pop(name, None)
        if v is not None:
            return v.reshape(v.shape[0]) if len(v.shape) == 2 and v.shape[1] == 1 else v
        return self.file.read_chunk(frame=idx, name=name, offset=0, r_all=False, device=self._read_device)

    def _chunk_or_fallback(self, idx, name, initial, default):
        if self.file.chunk_exists(frame=idx, name=name, write_all=False):
            return self.file.read_chunk(frame=idx, name=name, offset=0, r_all=False), True
        if self._initial_frame is not None:
            return initial(self._initial_frame), False
        return default, False

    def _read_strings(self, idx, name):
        tmp = self.file.read_chunk(frame=idx, name=name, offset=0, r_all=False)
        if tmp.ndim == 1:
            tmp = tmp.reshape([-1, 1])
        tmp = numpy.ascontiguousarray(tmp).view(dtype=numpy.dtype((bytes, tmp.shape[1])))
        return list(a.decode('UTF-8') for a in tmp.reshape([tmp.shape[0]]))

    def _read_frame(self, idx):
        """Decode frame ``idx`` (ref: hoomd.py:724-902)."""
        if idx >= len(self):
            raise IndexError
        logger.debug('reading frame ' + str(idx) + ' from: ' + str(self.file))
        # frame 0 is the fallback source for chunks missing in later frames
        if self._initial_frame is None and idx != 0:
            self._read_frame(0)
        ahead = self._take_prefetched(idx)
        snap = Frame()
        cfg = snap.configuration
        v, hit = self._chunk_or_fallback(idx, 'configuration/step', lambda f0: f0.configuration.step,
                                         cfg._default_value['step'])
        cfg.step = v[0] if hit else v
        v, hit = self._chunk_or_fallback(idx, 'configuration/dimensions', lambda f0: f0.configuration.dimensions,
                                         cfg._default_value['dimensions'])
        cfg.dimensions = v[0] if hit else v
        cfg.box, _ = self._chunk_or_fallback(idx, 'configuration/box', lambda f0: f0.configuration.box,
                                             cfg._default_value['box'])

        read_here = {}  # particles fields read from THIS frame's chunks
        for path in ['particles', 'constraints']:
            container = getattr(snap, path)
            initial = getattr(self._initial_frame, path) if self._initial_frame is not None else None
            container.N = 0
            if self.file.chunk_exists(frame=idx, name=path + '/N', write_all=False):
                container.N = self.file.read_chunk(frame=idx, name=path + '/N', offset=0, r_all=False)[0]
            elif initial is not None:
                container.N = initial.N
            if 'types' in container._default_value:
                if self.file.chunk_exists(frame=idx, name=path + '/types', write_all=False):
                    container.types = self._read_strings(idx, path + '/types')
                else:
                    container.types = initial.types if initial is not None else container._default_value['types']
            if 'type_shapes' in container._default_value and path == 'particles':
                if self.file.chunk_exists(frame=idx, name=path + '/type_shapes', write_all=False):
                    container.type_shapes = [json.loads(s) for s in self._read_strings(idx, path + '/type_shapes')]
                else:
                    container.type_shapes = (initial.type_shapes if initial is not None
                                             else container._default_value['type_shapes'])
            for name in container._default_value:
                if name in ('N', 'types', 'type_shapes'):
                    continue
                if self.file.chunk_exists(frame=idx, name=path + '/' + name, write_all=False):
                    if path == 'particles':
                        container.__dict__[name] = self._read_big(idx, path + '/' + name, ahead)
                    else:
                        container.__dict__[name] = self.file.read_chunk(frame=idx, name=path + '/' + name,
                                                                        offset=0, r_all=False)
                    if path == 'particles':
                        read_here[name] = True
                else:
                    if initial is not None and initial.N == container.N:
                        container.__dict__[name] = initial.__dict__[name]
                        if initial.__dict__.get('_isdefault_' + name):
                            container.__dict__['_isdefault_' + name] = True
                        elif path == 'particles':
                            # rows are in FRAME 0's storage order, not this frame's (see _reordered)
                            container.__dict__['_fallback_' + name] = True
                    else:
                        # the reference fills an N-row array with the default and marks it read-only
                        # (hoomd.py:871-881) -- 76 B/particle of constants per frame 0; a zero-stride
                        # broadcast view has the same shape, dtype, values and read-only flag for free
                        tmp = numpy.array([container._default_value[name]])
                        s = list(tmp.shape)
                        s[0] = int(container.N)
                        container.__dict__[name] = numpy.broadcast_to(tmp[0], s)
                        container.__dict__['_isdefault_' + name] = True

        for log in self.file.find_matching_chunk_names('log/', False):
            if self.file.chunk_exists(frame=idx, name=log, write_all=False):
                if log.startswith('log/particles/'):
                    snap.log[log[4:]] = self._read_big(idx, log, ahead)
                else:
                    snap.log[log[4:]] = self.file.read_chunk(frame=idx, name=log, offset=0, r_all=False)
            elif self._initial_frame is not None and log[4:] in self._initial_frame.log:
                snap.log[log[4:]] = self._initial_frame.log[log[4:]]
                snap.__dict__.setdefault('_log_fallback', set()).add(log[4:])

        for v in ahead.values():  # prefetched but unused (should not happen)
            v.free()
        if self._initial_frame is None and idx == 0:
            self._initial_frame = snap
        self._start_prefetch(idx + 1)
        if self._reorder == 'id':
            return self._reordered(snap)
        return snap

    def _initial_reordered(self):
        """Frame 0 in particle-ID order (host arrays unless device=True), computed once: the source of fallback fields."""
        if getattr(self, '_initial_sorted', None) is None:
            self._initial_sorted = self._reordered(self._initial_frame)
        return self._initial_sorted

    def _reordered(self, snap):
        """Particle-ID order: every array with one row per particle is gathered by the stable
        argsort of this frame's ``log/particles/id`` (constant default fields are left alone)."""
        ids = snap.log.get(ID_CHUNK[4:])
        N = int(snap.particles.N)
        if ids is None:
            raise KeyError("reorder='id' needs the chunk " + ID_CHUNK + " in: " + str(self.file))
        out = Frame()
        out.configuration = snap.configuration
        out.constraints = snap.constraints
        out.state = snap.state
        out.particles.N = snap.particles.N
        out.particles.types = snap.particles.types
        out.particles.type_shapes = snap.particles.type_shapes
        # Per-particle arrays that fell back to frame 0 (chunk absent in this frame) are in frame 0's storage order:
        # gathering them with THIS frame's permutation would attach them to the wrong particles whenever the storage
        # order changed in between (particle migration between ranks).  They are taken from frame 0's own ID-ordered
        # result instead; that needs this frame to carry its own ids (otherwise frame 0's ids order everything).
        log_fb = snap.__dict__.get('_log_fallback', set())
        own_ids = ID_CHUNK[4:] not in log_fb
        first = self._initial_reordered() if (own_ids and snap is not self._initial_frame) else None
        todo = {}
        for name in _PARTICLE_FIELDS:
            v = snap.particles.__dict__[name]
            if snap.particles.__dict__.get('_isdefault_' + name):
                out.particles.__dict__[name] = v
            elif first is not None and snap.particles.__dict__.get('_fallback_' + name):
                out.particles.__dict__[name] = first.particles.__dict__[name]
            else:
                todo['p:' + name] = v
        for k, v in snap.log.items():
            if k.startswith('particles/') and k != ID_CHUNK[4:] and len(v) == N:
                if first is not None and k in log_fb:
                    out.log[k] = first.log[k]
                else:
                    todo['l:' + k] = v
            else:
                out.log[k] = v
        if self._read_device:
            # frame-0 fallbacks held on the host: upload so one device gather covers everything
            todo = {k: (v if is_device_array(v) else DeviceArray.from_numpy(v)) for k, v in todo.items()}
            if not is_device_array(ids):
                ids = DeviceArray.from_numpy(numpy.ascontiguousarray(ids, dtype=numpy.uint32))
        sorted_ids, res = reorder_by_id(ids, todo, device=self._read_device)
        if self._read_device and not self._device:
            from .devmem import download
            host = {k: download(v) for k, v in res.items()}
            sorted_host = download(sorted_ids)
            for v in list(res.values()) + [sorted_ids]:
                v.free()
            sorted_ids, res = sorted_host, host
        out.log[ID_CHUNK[4:]] = sorted_ids
        for k, v in res.items():
            if k[0] == 'p':
                out.particles.__dict__[k[2:]] = v
            else:
                out.log[k[2:]] = v
        return out

    def __getitem__(self, key):
        """Index trajectory frames: an int returns a Frame, a slice a view."""
        if isinstance(key, slice):
            return _HOOMDTrajectoryView(self, range(*key.indices(len(self))))
        elif isinstance(key, (int, numpy.integer)):
            key = int(key)
            if key < 0:
                key += len(self)
            if key >= len(self) or key < 0:
                raise IndexError()
            return self._read_frame(key)
        else:
            raise TypeError

    def __iter__(self):
        return _HOOMDTrajectoryIterable(self, range(len(self)))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.file.close()


def open(name, mode='r', reorder=None, device=False, prefetch=None):
    """Open a hoomd schema PGSD file (ref: hoomd.py:943-989).  ``reorder`` / ``device`` / ``prefetch``:
    see :py:class:`HOOMDTrajectory`."""
    pgsdfileobj = fl.open(name=str(name), mode=mode, application='pgsd.hoomd ' + __version__,
                          schema='hoomd', schema_version=[1, 4])
    return HOOMDTrajectory(pgsdfileobj, reorder=reorder, device=device, prefetch=prefetch)


def read_log(name, scalar_only=False):
    """Read the logged data of a file into a dict of arrays over frames (ref: hoomd.py:992-1075)."""
    with fl.open(name=str(name), mode='r', application='pgsd.hoomd ' + __version__, schema='hoomd',
                 schema_version=[1, 4]) as f:
        names = f.find_matching_chunk_names('log/')
        names.insert(0, 'configuration/step')
        if len(names) == 1:
            warnings.warn('No logged data in file: ' + str(name), RuntimeWarning)
        out = dict()
        for log in names:
            exists0 = f.chunk_exists(frame=0, name=log, write_all=False)
            is_step = log == 'configuration/step'
            if not (exists0 or is_step):
                continue
            tmp = numpy.array([0], dtype=numpy.uint64) if (is_step and not exists0) else f.read_chunk(frame=0, name=log)
            if scalar_only and not tmp.shape[0] == 1:
                continue
            if tmp.shape[0] == 1:
                out[log] = numpy.full(fill_value=tmp[0], shape=(f.nframes,))
            else:
                out[log] = numpy.tile(tmp, (f.nframes,) + tuple(1 for _ in tmp.shape))
        for idx in range(1, f.nframes):
            for log in out.keys():
                if not f.chunk_exists(frame=idx, name=log, write_all=False):
                    continue
                data = f.read_chunk(frame=idx, name=log)
                if len(out[log][idx].shape) == 0:
                    out[log][idx] = data[0]
                else:
                    out[log][idx] = data
    return out
