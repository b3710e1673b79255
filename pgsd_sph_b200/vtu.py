"""pgsd2vtu -- the consumer of ID-ordered frames.

The reference ships this converter only as a listing in its manual
(/root/reference/pgsd/doc/pgsd.tex:1226-1265; the in-tree test_pgsd2vtu.py is a stub): per frame it
splits position / velocity into columns, casts every array to contiguous float64 and hands them to
``pyevtk.hl.pointsToVTK``.  pyevtk is a third-party module that the reference neither pins nor
vendors and that is not installed here, so the .vtu bytes written by :func:`write_vtu` are this
repository's own (VTK XML UnstructuredGrid, appended raw data) -- PARITY UNPINNED.  What is pinned is
the array preparation (:func:`point_arrays`): it must equal the listing's
``numpy.ascontiguousarray(col, dtype=numpy.float64)`` bit for bit.  For device-resident frames the
column split + float32->float64 cast runs on the GPU (K1's strided path), one call per column.
"""
import ctypes as C
import struct

import numpy

from . import _lib
from .devmem import DeviceArray, is_device_array


def _col_f64(a, j=None):
    """Column j of an (N, M) array (or the (N,) array itself) as contiguous float64."""
    if is_device_array(a):
        from .devmem import as_device_view
        from .fl import _NP_TO_PGSD
        ptr, shape, dt, strides, keep = as_device_view(a)
        n = int(shape[0])
        m = int(shape[1]) if len(shape) == 2 else 1
        out = DeviceArray((n,), numpy.float64)
        if n:
            col = (_lib.Column * 1)(_lib.Column(ptr + (j or 0) * dt.itemsize, m))
            _lib.check(_lib.load().pgsd_b200_pack_soa(out.ptr, _lib.TYPE_DOUBLE, n, 1, _NP_TO_PGSD[dt], col, None),
                       "pgsd_b200_pack_soa")
        del keep
        return out
    a = numpy.asarray(a)
    return numpy.ascontiguousarray(a if j is None else a[:, j], dtype=numpy.float64)


def point_arrays(snapshot):
    """(x, y, z, point_data) exactly as the manual's converter prepares them (pgsd.tex:1249-1259)."""
    p = snapshot.particles
    x, y, z = (_col_f64(p.position, j) for j in range(3))
    point_data = {
        'density': _col_f64(p.density),
        'pressure': _col_f64(p.pressure),
        'slength': _col_f64(p.slength),
        'velocity': tuple(_col_f64(p.velocity, j) for j in range(3)),
    }
    return x, y, z, point_data


def _host(a):
    return a.to_numpy() if is_device_array(a) else numpy.ascontiguousarray(a)


def write_vtu(path, x, y, z, point_data):
    """Write points + per-point data as a VTK XML UnstructuredGrid of vertex cells with appended raw
    (uncompressed, little-endian, UInt64 headers) data.  Returns the file name written."""
    x, y, z = _host(x), _host(y), _host(z)
    n = x.shape[0]
    blocks, arrays = [], []

    def add(name, arr, ncomp=1):
        off = sum(8 + b.nbytes for b in blocks)
        blocks.append(arr)
        t = {'float64': 'Float64', 'float32': 'Float32', 'int64': 'Int64', 'uint8': 'UInt8', 'uint32': 'UInt32',
             'int32': 'Int32'}[arr.dtype.name]
        return f'<DataArray type="{t}" Name="{name}" NumberOfComponents="{ncomp}" format="appended" offset="{off}"/>'

    pts = numpy.ascontiguousarray(numpy.stack([x, y, z], axis=1))
    xml_points = add('points', pts, 3)
    xml_cells = [add('connectivity', numpy.arange(n, dtype=numpy.int64)),
                 add('offsets', numpy.arange(1, n + 1, dtype=numpy.int64)),
                 add('types', numpy.ones(n, dtype=numpy.uint8))]
    xml_pd = []
    for name, v in point_data.items():
        if isinstance(v, tuple):
            xml_pd.append(add(name, numpy.ascontiguousarray(numpy.stack([_host(c) for c in v], axis=1)), len(v)))
        else:
            xml_pd.append(add(name, _host(v)))
    if not path.endswith('.vtu'):
        path = path + '.vtu'
    with open(path, 'wb') as f:
        f.write(('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" '
                 'header_type="UInt64">\n<UnstructuredGrid>\n'
                 f'<Piece NumberOfPoints="{n}" NumberOfCells="{n}">\n<Points>\n{xml_points}\n</Points>\n<Cells>\n'
                 + '\n'.join(xml_cells) + '\n</Cells>\n<PointData>\n' + '\n'.join(xml_pd)
                 + '\n</PointData>\n</Piece>\n</UnstructuredGrid>\n<AppendedData encoding="raw">\n_').encode())
        for b in blocks:
            f.write(struct.pack('<Q', b.nbytes))
            f.write(b.tobytes())
        f.write(b'\n</AppendedData>\n</VTKFile>\n')
    return path


def convert(gsd_path, reorder='id', device=True):
    """The manual's loop: one .vtu per frame, named <file>_<count:05d>.vtu; frames in particle-ID
    order so that point i is particle i in every output file.  Returns the list of files."""
    from . import hoomd
    out = []
    with hoomd.open(gsd_path, 'r', reorder=reorder, device=device) as t:
        for count, snapshot in enumerate(t, start=1):
            pname = gsd_path.replace('.gsd', f'_{count:05d}')
            x, y, z, pd = point_arrays(snapshot)
            out.append(write_vtu(pname, x, y, z, pd))
    return out
