"""pgsd2vtu -- the consumer of ID-ordered frames.

The reference ships this converter only as a listing in its manual
(/root/reference/pgsd/doc/pgsd.tex:1226-1265; the in-tree test_pgsd2vtu.py is a stub): per frame it
splits position / velocity into columns, casts every array to contiguous float64 and hands them to
``pyevtk.hl.pointsToVTK(path, x, y, z, pointData)``.  pyevtk is a third-party module that the reference
neither pins nor vendors and that is not installed here.  :func:`write_vtu` takes the same arguments and
lays the file out the way pyevtk's writer does (restated call by call in the test suite's VTU checker from its
published source: VTK XML UnstructuredGrid of vertex cells, Int32 connectivity / offsets, UInt8 types, appended raw
data with UInt64 block sizes, x/y/z interleaved), but no pyevtk output exists to compare with --
CONTAINER PARITY UNPINNED.  What is pinned is the array preparation (:func:`point_arrays`): it must equal
the listing's ``numpy.ascontiguousarray(col, dtype=numpy.float64)`` bit for bit.

Device-resident frames never become numpy arrays: the column split + float32->float64 cast (K1's strided
path) and the x/y/z interleave (K1 again, three unit-stride float64 columns -> (N, 3)) run on the GPU, the
blocks are copied into a page-locked image of the whole file whose constant parts (XML, cell arrays) are
kept between frames, and the image goes to the file through the library's file stage from several threads
(`pgsd_b200_file_stage_write`: page-owned mapped copies on tmpfs, pwrite elsewhere -- K3's host side).
"""
import ctypes as C
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy

from . import _lib
from .devmem import D2H, DeviceArray, is_device_array, pinned_pool

_VTK_TYPE = {'float64': 'Float64', 'float32': 'Float32', 'int64': 'Int64', 'uint64': 'UInt64', 'int32': 'Int32',
             'uint32': 'UInt32', 'int16': 'Int16', 'uint16': 'UInt16', 'int8': 'Int8', 'uint8': 'UInt8'}
_VTK_VERTEX = 1          # cell type id of a vertex
_PIECE = 16 << 20        # file-stage piece: a multiple of the page size


def _writers():
    """Threads that put the image into the file: as the library sizes its own file stage -- the host's cores divided
    by the ranks that share it (torchrun's LOCAL_WORLD_SIZE), 2..8."""
    try:
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    except ValueError:
        ranks = 1
    return max(2, min(8, (os.cpu_count() or 8) // ranks))


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


# ---- file layout ---------------------------------------------------------------------------------------------------
def _dtype_of(a):
    if is_device_array(a):
        from .devmem import as_device_view
        return as_device_view(a)[2]
    return numpy.dtype(a.dtype)


def _layout(n, coord_dtype, point_data):
    """XML text before / after the appended data and the blocks in file order.

    Returns (head bytes, tail bytes, blocks) with blocks = [(kind, key, dtype, ncomp, nbytes)], kind in
    'points' | 'connectivity' | 'offsets' | 'types' | 'data'.  Every block is preceded by its size as UInt64."""
    blocks = []
    offset = 0

    def array(name, dtype, ncomp, kind, key=None):
        nonlocal offset
        dtype = numpy.dtype(dtype)
        s = (f'\n<DataArray Name="{name}" NumberOfComponents="{ncomp}" type="{_VTK_TYPE[dtype.name]}" '
             f'format="appended" offset="{offset}"/>')
        nbytes = n * ncomp * dtype.itemsize
        blocks.append((kind, key, dtype, ncomp, nbytes))
        offset += nbytes + 8
        return s

    xml = ['<?xml version="1.0"?>',
           '\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
           '\n<UnstructuredGrid>',
           f'\n<Piece NumberOfPoints="{n}" NumberOfCells="{n}">',
           '\n<Points>', array('points', coord_dtype, 3, 'points'), '\n</Points>',
           '\n<Cells>', array('connectivity', numpy.int32, 1, 'connectivity'), array('offsets', numpy.int32, 1, 'offsets'),
           array('types', numpy.uint8, 1, 'types'), '\n</Cells>']
    if point_data:
        keys = list(point_data.keys())
        scalars = next((k for k in keys if not isinstance(point_data[k], tuple)), None)
        vectors = next((k for k in keys if isinstance(point_data[k], tuple)), None)
        xml.append('\n<PointData' + (f' scalars="{scalars}"' if scalars else '') + (f' vectors="{vectors}"' if vectors else '') + '>')
        for k in keys:
            v = point_data[k]
            if isinstance(v, tuple):
                if len(v) != 3:
                    raise ValueError(f"point data '{k}': vector data must be a tuple of 3 arrays")
                xml.append(array(k, _dtype_of(v[0]), 3, 'data', k))
            else:
                xml.append(array(k, _dtype_of(v), 1, 'data', k))
        xml.append('\n</PointData>')
    xml += ['\n</Piece>', '\n</UnstructuredGrid>', '\n<AppendedData encoding="raw">\n_']
    return ''.join(xml).encode(), b'\n</AppendedData>\n</VTKFile>', blocks


class _Image:
    """Byte image of one .vtu file for (n, dtypes, keys): XML, block sizes and the cell arrays are written once."""

    def __init__(self, n, coord_dtype, point_data, pinned):
        head, tail, self.blocks = _layout(n, coord_dtype, point_data)
        total = len(head) + sum(8 + b[4] for b in self.blocks) + len(tail)
        self.buf = pinned_pool().empty((total,), numpy.uint8) if pinned else numpy.empty(total, dtype=numpy.uint8)
        self.buf[:len(head)] = numpy.frombuffer(head, dtype=numpy.uint8)
        self.buf[total - len(tail):] = numpy.frombuffer(tail, dtype=numpy.uint8)
        self.at = []  # byte offset of every block's data
        pos = len(head)
        for kind, key, dtype, ncomp, nbytes in self.blocks:
            self.buf[pos:pos + 8] = numpy.frombuffer(struct.pack('<Q', nbytes), dtype=numpy.uint8)
            pos += 8
            self.at.append(pos)
            if kind in ('connectivity', 'offsets'):
                # unaligned destination: fill through a byte view in pieces
                step = 1 << 22
                for s in range(0, n, step):
                    e = min(n, s + step)
                    a = numpy.arange(s + (kind == 'offsets'), e + (kind == 'offsets'), dtype=numpy.int32)
                    self.buf[pos + 4 * s: pos + 4 * e] = a.view(numpy.uint8)
            elif kind == 'types':
                self.buf[pos:pos + n] = _VTK_VERTEX
            pos += nbytes
        self.total = total

    def block_view(self, i):
        return self.buf[self.at[i]: self.at[i] + self.blocks[i][4]]


_images = {}


def _image_for(n, coord_dtype, point_data, pinned):
    sig = (n, numpy.dtype(coord_dtype).name, pinned,
           tuple((k, isinstance(v, tuple), _dtype_of(v[0] if isinstance(v, tuple) else v).name) for k, v in (point_data or {}).items()))
    img = _images.get(sig)
    if img is None:
        _images.clear()  # one image at a time: a 16 Mi-particle file is 1.4 GB
        img = _images[sig] = _Image(n, coord_dtype, point_data, pinned)
    return img


def _interleave_device(cols, n, dtype):
    """Three unit-stride device columns -> one (n, 3) device array (K1)."""
    from .devmem import as_device_view
    from .fl import _NP_TO_PGSD
    out = DeviceArray((n, 3), dtype)
    if n:
        views = [as_device_view(c) for c in cols]
        arr = (_lib.Column * 3)(*[_lib.Column(v[0], 1) for v in views])
        t = _NP_TO_PGSD[numpy.dtype(dtype)]
        _lib.check(_lib.load().pgsd_b200_pack_soa(out.ptr, t, n, 3, t, arr, None), "pgsd_b200_pack_soa")
        del views
    return out


def _fill_block(img, i, src, n):
    """Block i of the image <- src: an array or a tuple of 3 arrays, host or device."""
    kind, key, dtype, ncomp, nbytes = img.blocks[i]
    dst = img.block_view(i)
    parts = src if isinstance(src, tuple) else (src,)
    for c in parts:
        if int(c.shape[0]) != n or len(c.shape) != 1 or _dtype_of(c) != dtype:
            raise ValueError("write_vtu: every array must be 1-D with one value per point and one dtype per block")
    if nbytes == 0:
        return
    if is_device_array(parts[0]):
        from .devmem import as_device_view
        for c in parts:
            st = as_device_view(c)[3]
            if st is not None and tuple(st) != (dtype.itemsize,):
                raise ValueError("write_vtu: device arrays must be contiguous")
        dev = _interleave_device(parts, n, dtype) if ncomp == 3 else parts[0]
        _lib.check(_lib.load().pgsd_b200_memcpy(dst.ctypes.data, as_device_view(dev)[0], nbytes, D2H), "D2H copy")
    elif ncomp == 3:
        out = dst.view(numpy.uint8).reshape(n, 3, dtype.itemsize)
        for j, c in enumerate(parts):
            out[:, j, :] = numpy.ascontiguousarray(c).view(numpy.uint8).reshape(n, dtype.itemsize)
    else:
        dst[:] = numpy.ascontiguousarray(parts[0]).view(numpy.uint8)


def _write_image(path, img):
    """The image -> file: page-aligned pieces through the library's file stage, several threads."""
    lib = _lib.load()
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        base = img.buf.ctypes.data
        pieces = [(o, min(_PIECE, img.total - o)) for o in range(0, img.total, _PIECE)]

        def put(p):
            return lib.pgsd_b200_file_stage_write(fd, C.c_void_p(base + p[0]), p[0], p[1], 0)

        if len(pieces) == 1:
            rcs = [put(pieces[0])]
        else:
            with ThreadPoolExecutor(max_workers=min(_writers(), len(pieces))) as ex:
                rcs = list(ex.map(put, pieces))
        for rc in rcs:
            _lib.check(rc, "pgsd_b200_file_stage_write")
    finally:
        os.close(fd)


def write_vtu(path, x, y, z, point_data=None):
    """``pyevtk.hl.pointsToVTK(path, x, y, z, pointData)``: points + per-point data (arrays, or tuples of three
    arrays for vectors) as a VTK XML UnstructuredGrid of vertex cells with appended raw data.  Arrays may live on the
    host (numpy) or on the device.  Returns the file name written (path + '.vtu')."""
    n = int(x.shape[0])
    if int(y.shape[0]) != n or int(z.shape[0]) != n:
        raise ValueError("write_vtu: x, y, z must have one value per point")
    on_device = is_device_array(x)
    img = _image_for(n, _dtype_of(x), point_data, on_device)
    for i, (kind, key, dtype, ncomp, nbytes) in enumerate(img.blocks):
        if kind == 'points':
            _fill_block(img, i, (x, y, z), n)
        elif kind == 'data':
            _fill_block(img, i, point_data[key], n)
    if not path.endswith('.vtu'):
        path = path + '.vtu'
    _write_image(path, img)
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
