"""TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED.

CPU restatement of the writer behind the reference's pgsd2vtu listing (/root/reference/pgsd/doc/pgsd.tex:1226-1265:
``from pyevtk.hl import pointsToVTK as vtk`` ... ``vtk(pname, x, y, z, pointData=point_data)``).

The algorithm lives in a third-party dependency that is ABSENT from /root/reference and from this image: pyevtk
(PyPI "pyevtk"; the reference names no version -- neither requirements file nor setup metadata mention it -- so the
current release line 1.x is restated: pyevtk/hl.py ``pointsToVTK``, pyevtk/vtk.py ``VtkFile``, pyevtk/xml.py
``XmlWriter``, pyevtk/evtk.py ``writeBlockSize`` / ``writeArrayToFile`` / ``writeArraysToFile``).  It is written
down here from the published source as remembered, call by call in pyevtk's own order, and could NOT be checked
against pyevtk output (no network, no wheel): there is no golden vector, no reference test and no reference output
for this path.  It pins the product's fast writer (pgsd_sph_b200/vtu.py) to one sequential, obviously ordered
statement of the layout -- not to pyevtk itself.

Layout: XML (one element per line, attributes in call order), then ``<AppendedData encoding="raw">`` + newline + '_',
then every array as [UInt64 byte count][raw little-endian data]; a tuple (x, y, z) is written interleaved
x0 y0 z0 x1 ...; cells are vertices: connectivity = 0..n-1 (Int32), offsets = 1..n (Int32), types = 1 (UInt8).
"""
import io
import struct

import numpy as np

_NP_TO_VTK = {'int8': 'Int8', 'uint8': 'UInt8', 'int16': 'Int16', 'uint16': 'UInt16', 'int32': 'Int32', 'uint32': 'UInt32',
              'int64': 'Int64', 'uint64': 'UInt64', 'float32': 'Float32', 'float64': 'Float64'}


class XmlWriter:
    """pyevtk/xml.py: tags are left open until the next element / text so that attributes can follow."""

    def __init__(self, stream):
        self.stream = stream
        self.open_tag = False
        self.current = []
        self.stream.write(b'<?xml version="1.0"?>')

    def open_element(self, tag):
        if self.open_tag:
            self.stream.write(b">")
        self.stream.write(("\n<%s" % tag).encode())
        self.open_tag = True
        self.current.append(tag)
        return self

    def close_element(self, tag=None):
        if tag:
            assert self.current.pop() == tag
            if self.open_tag:
                self.stream.write(b">")
                self.open_tag = False
            self.stream.write(("\n</%s>" % tag).encode())
        else:
            self.stream.write(b"/>")
            self.open_tag = False
            self.current.pop()
        return self

    def add_text(self, text):
        if self.open_tag:
            self.stream.write(b">\n")
            self.open_tag = False
        self.stream.write(text.encode())
        return self

    def add_attributes(self, **kwargs):
        assert self.open_tag
        for key, value in kwargs.items():
            self.stream.write((' %s="%s"' % (key, value)).encode())
        return self


class VtkFile:
    """pyevtk/vtk.py VtkFile for ftype VtkUnstructuredGrid."""

    def __init__(self, stream):
        self.xml = XmlWriter(stream)
        self.offset = 0
        self.appended_open = False
        self.xml.open_element("VTKFile").add_attributes(type="UnstructuredGrid", version="1.0",
                                                         byte_order="LittleEndian", header_type="UInt64")

    def add_data(self, name, data):
        if isinstance(data, tuple):
            assert len(data) == 3
            self._add_header(name, data[0].dtype, data[0].size, 3)
        else:
            self._add_header(name, data.dtype, data.size, 1)

    def _add_header(self, name, dtype, nelem, ncomp):
        self.xml.open_element("DataArray")
        self.xml.add_attributes(Name=name, NumberOfComponents=ncomp, type=_NP_TO_VTK[dtype.name], format="appended",
                                offset=self.offset)
        self.xml.close_element()
        self.offset += nelem * ncomp * dtype.itemsize + 8

    def append_data(self, data):
        if not self.appended_open:
            self.xml.open_element("AppendedData").add_attributes(encoding="raw").add_text("_")
            self.appended_open = True
        s = self.xml.stream
        if isinstance(data, tuple):
            x, y, z = data
            s.write(struct.pack("<Q", 3 * x.size * x.dtype.itemsize))
            s.write(np.stack([x, y, z], axis=1).tobytes())
        else:
            s.write(struct.pack("<Q", data.size * data.dtype.itemsize))
            s.write(np.ascontiguousarray(data).tobytes())
        return self

    def save(self):
        if self.appended_open:
            self.xml.close_element("AppendedData")
        self.xml.close_element("VTKFile")


def points_to_vtk_bytes(x, y, z, data=None):
    """pyevtk/hl.py pointsToVTK(path, x, y, z, data) -> the bytes of path + '.vtu'."""
    assert x.size == y.size == z.size
    npoints = x.size
    offsets = np.arange(start=1, stop=npoints + 1, dtype="int32")
    connectivity = np.arange(npoints, dtype="int32")
    cell_types = np.empty(npoints, dtype="uint8")
    cell_types[:] = 1  # VtkVertex.tid

    out = io.BytesIO()
    w = VtkFile(out)
    w.xml.open_element("UnstructuredGrid")
    w.xml.open_element("Piece").add_attributes(NumberOfPoints=npoints, NumberOfCells=npoints)
    w.xml.open_element("Points")
    w.add_data("points", (x, y, z))
    w.xml.close_element("Points")
    w.xml.open_element("Cells")
    w.add_data("connectivity", connectivity)
    w.add_data("offsets", offsets)
    w.add_data("types", cell_types)
    w.xml.close_element("Cells")
    if data:
        keys = list(data.keys())
        scalars = next((k for k in keys if isinstance(data[k], np.ndarray)), None)
        vectors = next((k for k in keys if isinstance(data[k], tuple)), None)
        w.xml.open_element("PointData")
        if scalars:
            w.xml.add_attributes(scalars=scalars)
        if vectors:
            w.xml.add_attributes(vectors=vectors)
        for k in keys:
            w.add_data(k, data[k])
        w.xml.close_element("PointData")
    w.xml.close_element("Piece")
    w.xml.close_element("UnstructuredGrid")
    w.append_data((x, y, z))
    w.append_data(connectivity).append_data(offsets).append_data(cell_types)
    if data:
        for k in list(data.keys()):
            w.append_data(data[k])
    w.save()
    return out.getvalue()


def parse_vtu(raw):
    """Independent reader for round-trip checks: -> {name: array} (3-component arrays as (n, 3))."""
    import re
    head, rest = raw.split(b'<AppendedData encoding="raw">\n_', 1)
    out = {}
    n = int(re.search(rb'NumberOfPoints="(\d+)"', head).group(1))
    for m in re.finditer(rb'<DataArray Name="([^"]+)" NumberOfComponents="(\d+)" type="([^"]+)" format="appended" offset="(\d+)"/>', head):
        name, ncomp, t, off = m.group(1).decode(), int(m.group(2)), m.group(3).decode(), int(m.group(4))
        dt = np.dtype({v: k for k, v in _NP_TO_VTK.items()}[t])
        size = struct.unpack("<Q", rest[off:off + 8])[0]
        assert size == n * ncomp * dt.itemsize
        a = np.frombuffer(rest[off + 8: off + 8 + size], dtype=dt)
        out[name] = a.reshape(n, ncomp) if ncomp > 1 else a
    return out
