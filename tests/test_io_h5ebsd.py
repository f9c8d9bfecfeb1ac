"""kikuchipy h5ebsd reader / writer (kikuchipy_b200/io_h5ebsd.py) and the HDF5 parser under it
(kikuchipy_b200/_hdf5.py): against the two kikuchipy h5ebsd files of the reference's data directory
(written by h5py; committed under tests/golden/h5ebsd as data fixtures: the nine nickel patterns of
BASELINE configs[0], chunked and contiguous), against files this package writes, and - when the
reference tree is mounted - against its other HDF5 sample files (gzip-compressed chunks,
variable-length strings).  Semantics follow /root/reference/tests/test_io/test_kikuchipy_h5ebsd.py."""

import os
import warnings

import numpy as np
import pytest

import kikuchipy_b200 as kb
from kikuchipy_b200 import _hdf5, io_h5ebsd

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "h5ebsd")
CHUNKED = os.path.join(FIX, "kikuchipy_patterns.h5")
CONTIGUOUS = os.path.join(FIX, "kikuchipy_patterns_nochunks.h5")
REF_DATA = "/root/reference/src/kikuchipy/data"


def _nickel(golden):
    return golden("config1_nickel_x_1000.npz")["nickel"].reshape(3, 3, 60, 60)


def test_load_reference_file(golden):
    """``test_load``: (3, 3, 60, 60) uint8, the nine nickel patterns, 1.5 um steps, the axes the
    reference's ``ni_small_axes_manager`` fixture describes, header values, per-point PCs."""
    s = kb.load(CHUNKED)
    assert isinstance(s, kb.H5EBSDScan)
    assert s.data.shape == (3, 3, 60, 60) and s.data.dtype == np.uint8
    assert np.array_equal(s.data, _nickel(golden))
    assert s.step_sizes == (1.5, 1.5)
    assert [(a["name"], a["size"], a["scale"], a["units"]) for a in s.axes] == [
        ("y", 3, 1.5, "um"), ("x", 3, 1.5, "um"), ("dy", 60, 1.0, "um"), ("dx", 60, 1.0, "um")]
    assert s.static_background.shape == (60, 60) and s.static_background.dtype == np.uint8
    assert s.detector.shape == (60, 60) and s.detector.pc.shape == (3, 3, 3)
    assert s.detector.sample_tilt == 70.0 and s.detector.binning == 8
    assert abs(s.detector.pc[0, 0, 0] - 0.4214844) < 1e-7
    sem = s.metadata["Acquisition_instrument"]["SEM"]
    assert sem == {"beam_energy": 20.0, "magnification": 200, "microscope": "Hitachi SU-6600", "working_distance": 24.7}
    assert s.metadata["General"] == {"original_filename": "kikuchipy_patterns", "title": "kikuchipy_patterns S..."}
    assert s.original_metadata["manufacturer"] == "kikuchipy" and s.original_metadata["n_rows"] == 3
    assert s.xmap["header"]["phases"]["0"]["name"] == "ni" and s.xmap["data"]["phi1"].shape == (9,)
    assert s.xmap["data"]["is_in_data"].dtype == np.bool_ and s.xmap["data"]["is_in_data"].all()
    # the contiguous (unchunked) file holds the first pattern only: 0-D navigation, squeezed
    s2 = io_h5ebsd.load_h5ebsd(CONTIGUOUS)
    assert s2.data.shape == (60, 60) and np.array_equal(s2.data, s.data[0, 0]) and len(s2.axes) == 2


@pytest.mark.parametrize("names", [["Scan 1", "Scan 2"], ["Scan 1", "Scan 2", "Scan 3"], ["Scan 3"], "Scan 2"])
def test_load_multiple(names):
    """``test_load_multiple``: lists return lists, a missing scan warns (or raises when it is the only
    one asked for), the title carries the scan name."""
    if names == ["Scan 1", "Scan 2", "Scan 3"]:
        with pytest.warns(UserWarning, match="Scan 'Scan 3' is not among "):
            s1, s2 = kb.load_h5ebsd(CHUNKED, scan_group_names=names)
    elif names == ["Scan 3"]:
        with pytest.raises(OSError, match="Scan 'Scan 3' is not among the"):
            kb.load_h5ebsd(CHUNKED, scan_group_names=names)
        return
    elif isinstance(names, list):
        s1, s2 = kb.load_h5ebsd(CHUNKED, scan_group_names=names)
    else:
        s2 = kb.load_h5ebsd(CHUNKED, scan_group_names=names)
        assert s2.metadata["General"]["title"].startswith("kikuchipy_patterns S")
        s1 = kb.load_h5ebsd(CHUNKED)
    assert np.array_equal(s1.data, s2.data)


def test_save_load_cycle_and_padding(tmp_path, golden):
    """``test_load_save_cycle`` / ``test_save_multiple`` / ``test_load_with_padding``."""
    s = kb.load_h5ebsd(CHUNKED)
    out = str(tmp_path / "patterns_out.h5")
    kb.save_h5ebsd(out, s.data, detector=s.detector, static_background=s.static_background,
                   step_sizes=s.step_sizes, metadata=s.metadata, xmap=s.xmap)
    r = kb.load_h5ebsd(out)
    assert np.array_equal(r.data, s.data) and np.array_equal(r.static_background, s.static_background)
    assert np.array_equal(r.detector.pc, s.detector.pc) and r.step_sizes == s.step_sizes
    assert r.metadata["Acquisition_instrument"] == s.metadata["Acquisition_instrument"]
    assert r.xmap["header"]["phases"]["0"]["point_group"] == "m-3m"
    assert np.array_equal(r.xmap["data"]["phi1"], s.xmap["data"]["phi1"])
    # a second scan in the same file; an occupied scan number is refused
    kb.save_h5ebsd(out, s.data[:2], step_sizes=(2.0, 3.0), scan_number=2, add_scan=True)
    with pytest.raises(IOError, match="Invalid scan number"):
        kb.save_h5ebsd(out, s.data, scan_number=2, add_scan=True)
    a, b = kb.load_h5ebsd(out, scan_group_names=["Scan 1", "Scan 2"])
    assert np.array_equal(a.data, s.data) and b.data.shape == (2, 3, 60, 60) and b.step_sizes == (2.0, 3.0)
    assert b.static_background is None and b.detector.pc.shape == (3,) or b.detector.pc.size == 3
    # more map points announced than stored: zero padding with both of the reference's warnings
    with _hdf5.File(out) as f:
        tree = io_h5ebsd._tree_of(f.root)
    tree["Scan 1"]["EBSD"]["Header"]["n_columns"] = np.array([4])
    _hdf5.write(out, tree)
    with pytest.warns(UserWarning) as rec:
        p = kb.load_h5ebsd(out)
    msgs = [str(w.message) for w in rec]
    assert any("Signal shape (60, 60)" in m for m in msgs) and any("Data navigation shape" in m for m in msgs)
    assert p.data.shape == (3, 4, 60, 60) and np.array_equal(p.data.reshape(12, 60, 60)[:9], s.data.reshape(9, 60, 60))
    assert not p.data.reshape(12, 60, 60)[9:].any() and p.detector.pc.shape == (3, 4, 3)
    assert np.all(p.detector.pc[:, 3] == 0.5)


def test_file_checks(tmp_path):
    """``test_check_file_invalid_version`` / ``_no_scan_groups`` / ``test_load_manufacturer`` /
    ``test_read_patterns``."""
    p = str(tmp_path / "x.h5")
    _hdf5.write(p, {"manufacturer": "kikuchipy", "versionn": "0.1"})
    with pytest.raises(IOError, match="Could not find 'version' key in '(.*)'"):
        kb.load_h5ebsd(p)
    _hdf5.write(p, {"manufacturer": "kikuchipy", "version": "0.1"})
    with pytest.raises(IOError, match="(.*) as no top groups"):
        kb.load_h5ebsd(p)
    kb.save_h5ebsd(p, (255 * np.random.default_rng(0).random((10, 3, 5, 5))).astype(np.uint8))
    with _hdf5.File(p) as f:
        tree = io_h5ebsd._tree_of(f.root)
    nope = dict(tree, manufacturer="Nope")
    _hdf5.write(p, nope)
    with pytest.raises(IOError, match="'nope' is not among supported manufacturers"):
        kb.load_h5ebsd(p)
    del tree["Scan 1"]["EBSD"]["Data"]["patterns"]
    tree["Scan 1"]["EBSD"]["Data"]["other"] = np.zeros(3)
    _hdf5.write(p, tree)
    with pytest.raises(KeyError, match="Could not find patterns"):
        kb.load_h5ebsd(p)
    with open(p, "wb") as f:
        f.write(b"not an hdf5 file at all")
    with pytest.raises(IOError, match="is not an HDF5 file"):
        kb.load(p)


def test_navigation_shapes_round_trip(tmp_path):
    """0-, 1- and 2-D navigation shapes squeeze like the reference's ``get_data``."""
    rng = np.random.default_rng(1)
    for shape, want in (((5, 6), (5, 6)), ((4, 5, 6), (4, 5, 6)), ((1, 4, 5, 6), (4, 5, 6)), ((3, 1, 5, 6), (3, 5, 6)),
                        ((2, 3, 5, 6), (2, 3, 5, 6))):
        data = rng.integers(0, 65535, shape).astype(np.uint16)
        p = str(tmp_path / "n.h5")
        kb.save_h5ebsd(p, data)
        r = kb.load_h5ebsd(p)
        assert r.data.shape == want and r.data.dtype == np.uint16 and np.array_equal(r.data.reshape(shape), data)
        assert len(r.axes) == len(want)


def test_hdf5_writer_reader_round_trip(tmp_path):
    """Every type the writer knows, nested groups, a group with more entries than one symbol-table node
    holds, empty groups and big-endian input."""
    rng = np.random.default_rng(2)
    tree = {
        "u8": rng.integers(0, 255, (3, 4, 5)).astype(np.uint8), "i16": np.arange(-5, 5, dtype=np.int16),
        "i64": np.array([-(2 ** 40)]), "f16": np.arange(4, dtype=np.float16), "f32": rng.random((2, 2)).astype(np.float32),
        "f64": rng.random(7), "be": np.arange(5, dtype=">i4"), "text": "Hitachi SU-6600", "empty_text": "",
        "flag": np.array([True, False]), "scalar": 3.5,
        "many": {f"entry {i:04d}": np.array([i]) for i in range(300)},
        "nested": {"a": {"b": {"c": np.arange(3)}}, "empty": {}},
    }
    p = str(tmp_path / "t.h5")
    _hdf5.write(p, tree)
    with _hdf5.File(p) as f:
        assert sorted(f.keys()) == sorted(tree.keys())
        for k in ("u8", "i16", "i64", "f16", "f32", "f64"):
            assert np.array_equal(f[k].read(), tree[k]) and f[k].dtype == tree[k].dtype
        assert np.array_equal(f["be"].read(), np.arange(5))
        assert f["text"][()][0] == b"Hitachi SU-6600" and f["empty_text"][()][0] == b""
        assert np.array_equal(f["flag"].read(), [1, 0]) and f["scalar"].read()[0] == 3.5
        assert len(f["many"].keys()) == 300 and f["many/entry 0299"].read()[0] == 299
        assert np.array_equal(f["nested/a/b/c"].read(), np.arange(3)) and f["nested/empty"].keys() == []
        assert "nested/a/b" in f and "nested/a/x" not in f and f.get("nope") is None
        view = f["u8"].read(copy=False)
        assert not view.flags.writeable and np.array_equal(view, tree["u8"])
        del view


@pytest.mark.skipif(not os.path.isdir(REF_DATA), reason="reference tree not mounted")
def test_parser_on_the_reference_hdf5_files():
    """Files of other writers in the reference's data directory: gzip-compressed chunked 4-D datasets
    and variable-length strings (EMsoft), big-endian scalars (EDAX), and the fixtures are what the
    reference ships."""
    for name, fix in (("patterns.h5", CHUNKED), ("patterns_nochunks.h5", CONTIGUOUS)):
        with open(os.path.join(REF_DATA, "kikuchipy_h5ebsd", name), "rb") as a, open(fix, "rb") as b:
            assert a.read() == b.read()
    with _hdf5.File(os.path.join(REF_DATA, "emsoft_ebsd_master_pattern", "ni_mc_mp_20kv_uint8_gzip_opts9.h5")) as f:
        m = f["EMData/EBSDmaster/mLPNH"]
        assert m.shape == (1, 1, 401, 401) and m.chunks == (1, 1, 101, 101) and m.dtype == np.uint8
        a = m.read()
        assert a[0, 0, 0, :3].tolist() == [149, 127, 118] and a.std() > 1
        # a master pattern is symmetric under inversion of the Lambert square
        assert np.array_equal(a[0, 0], a[0, 0, ::-1, ::-1])
        assert f["NMLfiles/EBSDmasterNML"].read()[0] == b" &EBSDmastervars"
        assert f["EMData/MCOpenCL/accum_z"].shape == (21, 21, 101, 1)
    with _hdf5.File(os.path.join(REF_DATA, "edax_h5ebsd", "patterns.h5")) as f:
        assert f["Scan 1/EBSD/Header/Coordinate System/ID"].read()[0] == 2
        assert f["Scan 1/EBSD/Data/Pattern"].shape == (9, 60, 60)
        with pytest.raises(NotImplementedError, match="datatype class 6"):
            f["Scan 1/EBSD/Header/Phase/1/hkl Families"]
    with pytest.raises(NotImplementedError, match="only kikuchipy h5ebsd files"):
        kb.load_h5ebsd(os.path.join(REF_DATA, "edax_h5ebsd", "patterns.h5"))


@pytest.mark.gpu
def test_load_to_device_and_index(golden):
    """The file's patterns go to the GPU and through dictionary indexing without returning to the host."""
    import torch

    s = kb.load(CHUNKED, device=True)
    assert isinstance(s.data, torch.Tensor) and s.data.is_cuda and tuple(s.data.shape) == (3, 3, 60, 60)
    g = golden("config1_nickel_x_1000.npz")
    assert np.array_equal(s.data.cpu().numpy(), g["nickel"])
    from oracle import di_oracle as orc

    dic = orc.synthetic_dictionary(1000, (60, 60), seed=2)
    res = kb.dictionary_indexing(s.data, dic, keep_n=5, verbose=False)
    ridx, rsc = orc.dictionary_indexing(g["nickel"].reshape(9, 60, 60), dic, keep_n=5)
    assert np.array_equal(res.simulation_indices, ridx) and np.abs(res.scores - rsc).max() < 1e-4
