"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

from pgsd_sph_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(REPO, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//.*", "", src)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src):
        n = m.group(1)
        if n.startswith("pgsd_") or n == "is_root":
            names.add(n)
    # typedef'd callback type is not an exported symbol
    names.discard("pgsd_b200_allgather_fn")
    return names


def test_library_loads_and_exports_all_declared_symbols():
    lib = _lib.load()
    declared = declared_functions("pgsd.h") | declared_functions("pgsd_b200.h")
    assert len(declared) >= 50
    for name in sorted(declared):
        assert hasattr(lib, name), f"libpgsd_b200.so does not export {name}"
    # and the ctypes binding covers every declared function
    assert declared <= set(_lib.SIGNATURES), sorted(declared - set(_lib.SIGNATURES))


def test_struct_layouts_match_reference_lp64():
    # ref: pgsd.h:143-174 (256 B header), :182-204 (32 B entry), :297-353 (544 B handle)
    assert ctypes.sizeof(_lib.Header) == 256
    assert ctypes.sizeof(_lib.IndexEntry) == 32
    assert ctypes.sizeof(_lib.Handle) == 544
    H = _lib.Handle
    assert (H.header.offset, H.file_index.offset, H.frame_index.offset, H.buffer_index.offset) == (8, 264, 304, 344)
    assert (H.write_buffer.offset, H.file_names.offset, H.frame_names.offset) == (384, 408, 440)
    assert (H.cur_frame.offset, H.file_size.offset, H.open_flags.offset, H.name_map.offset) == (472, 480, 488, 496)
    assert (H.pending_index_entries.offset, H.maximum_write_buffer_size.offset) == (512, 520)
    assert (H.index_entries_to_buffer.offset, H.rank.offset, H.nprocs.offset) == (528, 536, 540)


def test_scalar_helpers(lib):
    assert lib.pgsd_make_version(2, 0) == 0x20000
    assert lib.pgsd_make_version(1, 4) == 0x10004
    assert [lib.pgsd_sizeof_type(t) for t in range(0, 12)] == [0, 1, 2, 4, 8, 1, 2, 4, 8, 4, 8, 0]
    assert lib.pgsd_b200_comm_size() == 1 and lib.pgsd_b200_comm_rank() == 0
    assert lib.pgsd_b200_comm_kind() == b"single"
    assert lib.is_root()


def test_no_oracle_import_in_product():
    """The product never routes through oracle/ (CPU fallbacks void parity claims)."""
    pkg = os.path.join(REPO, "pgsd_sph_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".pyx", ".pxd", ".cpp", ".cu", ".h")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "oracle/" not in text.replace("oracle/ref_driver.c's", "") or f in ("synth.py",), f
