"""CPU tests of result materialisation: the library's batched PDB writer must produce byte-for-byte what
Structure.write_PDB produces (which follows the reference's line format, libpdb.py:178-195)."""
import ctypes as C

import numpy as np

from golden_utils import GoldenCase


def test_batched_writer_equals_structure_writer(tmp_path):
    from mcsce_b200 import _lib
    from mcsce_b200.chem.tables import load_tables
    from mcsce_b200.structure import _round3_f32

    case = GoldenCase("ptm14")
    s = case.structure()
    t = load_tables()
    work = s.copy()
    first = int(work.res_nums[0])
    for i, res in enumerate(s.residue_types):
        tm = t.templates[res]
        work.append_atoms([tm.names[k] for k in tm.sc_idx], tm.resname, "A", first + i, np.zeros((tm.n_sc, 3)),
                          elements=[tm.elements[k] for k in tm.sc_idx])
    order = np.argsort(work.res_nums, kind="mergesort").astype(np.int32)
    final = work.copy(); final.reorder(order)
    text, line_len, off = final.pdb_template()
    n_atoms = len(work)
    rng = np.random.default_rng(0)
    coords = rng.normal(scale=40.0, size=(6, n_atoms, 3)).astype(np.float32)
    coords[0, :8] = [[0.0005, -0.0005, 0.0015], [999.9994, -999.9994, 12.3455], [1.0, -1.0, 0.0], [0.9995, -0.9995, 99.9995],
                     [123.4565, -123.4565, 7.7775], [-0.0004, 0.0004, -0.0], [500.0005, -500.0005, 0.1235], [2.5, 3.5, -4.5]]
    ok = np.array([1, 1, 0, 1, 1, 1], np.uint8)
    lib = _lib.load_library()
    n = C.c_int32(0)
    rc = lib.mcsce_write_pdbs(6, n_atoms, coords.ctypes.data_as(C.POINTER(C.c_float)), ok.ctypes.data_as(C.c_void_p),
                              order.ctypes.data_as(C.POINTER(C.c_int32)), text, line_len, off, str(tmp_path).encode(), 3, C.byref(n))
    assert rc == 0 and n.value == 5
    assert not (tmp_path / "2.pdb").exists()
    for c in (0, 1, 3, 4, 5):
        ref = final.copy()
        ref.xyz = _round3_f32(coords[c])[order]
        ref.write_PDB(str(tmp_path / "ref.pdb"))
        want = [ln.ljust(80) for ln in (tmp_path / "ref.pdb").read_text().splitlines()]
        got = (tmp_path / ("%d.pdb" % c)).read_text().splitlines()
        assert got == want, (c, next((a, b) for a, b in zip(got, want) if a != b))


def test_round3_matches_reference_text_route():
    from mcsce_b200.structure import _round3_f32

    rng = np.random.default_rng(1)
    x = rng.normal(scale=200.0, size=(5000, 3)).astype(np.float32)
    want = np.round(x, decimals=3).astype("<U8").astype(np.float32)
    assert np.array_equal(_round3_f32(x), want)
    big = np.array([[-1234.5678, 10000.1234, 5.0]], np.float32)
    assert np.array_equal(_round3_f32(big), np.round(big, decimals=3).astype("<U8").astype(np.float32))


def test_pdb_text_written_by_the_reference_round_trips_byte_for_byte(tmp_path):
    """``fix12.npz`` carries a complete structure as the reference's own ``write_PDB`` wrote it (libpdb.py:178-195);
    reading it and writing it back with this package's ``Structure`` must give the same file."""
    from golden_utils import GoldenCase
    from mcsce_b200.structure import Structure

    text = GoldenCase("fix12").pdb
    s = Structure(text).build()
    out = tmp_path / "again.pdb"
    s.write_PDB(str(out))
    assert out.read_text() == text
    assert len(s) == 201 and "HD13" in text
