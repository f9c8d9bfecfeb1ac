"""NORDIF reader (kikuchipy_b200/io_nordif.py) against a NORDIF directory written by the test and,
when the reference tree is mounted (build container), against its sample data set
(/root/reference/src/kikuchipy/data/nordif: the nine nickel patterns)."""

import os
import struct

import numpy as np
import pytest

import kikuchipy_b200 as kb
from kikuchipy_b200 import io_nordif
from oracle import ref_loader

SETTING = "\r\n".join([
    "[NORDIF]\t\t", "Software version\t3.1.2\t", "\t\t", "[Microscope]\t\t", "Manufacturer\tHitachi\t", "Model\tSU-6600\t",
    "Magnification\t200\t#", "Scan direction\tDirect\t", "Accelerating voltage\t20\tkV", "Working distance\t24.7\tmm",
    "Tilt angle\t70\t\xb0", "\t\t", "[Detector angles]\t\t", "Euler 1\t0\t\xb0", "Euler 2\t0\t\xb0", "Euler 3\t0\t\xb0",
    "Azimuthal\t2\t\xb0", "Elevation\t-3.5\t\xb0", "\t\t", "[Acquisition settings]\t\t", "Frame rate\t202\tfps",
    "Resolution\t6x4\tpx", "\t\t", "[Area]\t\t", "Top\t89.200 (223)\t\xb5m (px)", "Left\t60.384 (152)\t\xb5m (px)",
    "Width\t4.500 (11)\t\xb5m (px)", "Height\t4.500 (11)\t\xb5m (px)", "Step size\t1.500\t\xb5m", "Number of samples\t2x3\t#", ""])


def _write_bmp(path, img):
    h, w = img.shape
    stride = (w + 3) & ~3
    rows = np.zeros((h, stride), np.uint8)
    rows[:, :w] = img[::-1]
    palette = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 4, axis=1)
    palette[:, 3] = 0
    with open(path, "wb") as f:
        f.write(b"BM" + struct.pack("<IHHI", 54 + 1024 + rows.size, 0, 0, 54 + 1024))
        f.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 8, 0, rows.size, 2835, 2835, 256, 0))
        f.write(palette.tobytes() + rows.tobytes())


def test_load_nordif_directory(tmp_path):
    rng = np.random.default_rng(0)
    pats = rng.integers(0, 256, (2, 3, 4, 6), dtype=np.uint8)  # ny = 2, nx = 3, sy = 4, sx = 6
    bg = rng.integers(0, 256, (4, 6), dtype=np.uint8)
    pats.tofile(tmp_path / "Pattern.dat")
    (tmp_path / "Setting.txt").write_text(SETTING, encoding="latin-1")
    _write_bmp(tmp_path / "Background acquisition pattern.bmp", bg)
    scan = kb.load_nordif(str(tmp_path / "Pattern.dat"))
    assert np.array_equal(scan.data, pats) and np.array_equal(scan.static_background, bg)
    assert scan.step_sizes == (1.5, 1.5) and scan.detector.shape == (4, 6)
    assert (scan.detector.sample_tilt, scan.detector.tilt, scan.detector.azimuthal) == (70.0, 3.5, 2.0)
    sem = scan.metadata["Acquisition_instrument"]["SEM"]
    assert sem == {"beam_energy": 20.0, "magnification": 200, "microscope": "Hitachi SU-6600", "working_distance": 24.7}
    # explicit sizes, a line scan, a short file (zero padded with a warning), no setting file
    line = kb.load_nordif(str(tmp_path / "Pattern.dat"), scan_size=6, pattern_size=(6, 4))
    assert line.data.shape == (6, 4, 6) and np.array_equal(line.data, pats.reshape(6, 4, 6))
    with pytest.warns(UserWarning, match="zero padding"):
        big = kb.load_nordif(str(tmp_path / "Pattern.dat"), scan_size=(3, 3), pattern_size=(6, 4))
    assert big.data.shape == (3, 3, 4, 6) and not big.data[2].any()
    os.remove(tmp_path / "Setting.txt")
    os.remove(tmp_path / "Background acquisition pattern.bmp")
    with pytest.raises(ValueError, match="No setting file found"):
        kb.load_nordif(str(tmp_path / "Pattern.dat"))
    with pytest.warns(UserWarning):
        bare = kb.load_nordif(str(tmp_path / "Pattern.dat"), scan_size=(3, 2), pattern_size=(6, 4))
    assert bare.static_background is None and bare.detector is None and np.array_equal(bare.data, pats)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_load_reference_sample():
    folder = os.path.join(ref_loader.REFERENCE_ROOT, "src", "kikuchipy", "data", "nordif")
    scan = kb.load_nordif(os.path.join(folder, "Pattern.dat"))
    assert np.array_equal(scan.data, ref_loader.nickel_ebsd_small())
    assert scan.static_background.shape == (60, 60) and scan.static_background.dtype == np.uint8
    assert 50 < scan.static_background.mean() < 200 and scan.step_sizes == (1.5, 1.5)
    assert (scan.detector.sample_tilt, scan.detector.tilt, scan.detector.azimuthal) == (70.0, 0.0, 0.0)
    md, header, sizes, det = io_nordif.read_settings(os.path.join(folder, "Setting.txt"))
    assert sizes == {"ny": 3, "nx": 3, "sy": 60, "sx": 60, "step_y": 1.5, "step_x": 1.5}


@pytest.mark.gpu
def test_load_to_device_and_index(tmp_path):
    """File -> pinned upload -> preprocessing -> indexing without the patterns returning to the host."""
    from oracle import di_oracle as orc
    from oracle import preprocess_oracle as pp

    pats = orc.synthetic_experimental(12, (60, 60), seed=1).reshape(3, 4, 60, 60)
    pats.tofile(tmp_path / "Pattern.dat")
    with pytest.warns(UserWarning):
        scan = kb.load_nordif(str(tmp_path / "Pattern.dat"), scan_size=(4, 3), pattern_size=(60, 60), device=True)
    assert scan.data.is_cuda and tuple(scan.data.shape) == (3, 4, 60, 60)
    clean = kb.remove_dynamic_background(scan.data, "subtract", "spatial")
    assert clean.is_cuda and np.array_equal(clean.cpu().numpy(), pp.remove_dynamic_background(pats, "subtract", "spatial"))
    dic = orc.synthetic_dictionary(400, (60, 60), seed=2)
    res = kb.dictionary_indexing(clean, dic, keep_n=5, verbose=False)
    ridx, rsc = orc.dictionary_indexing(clean.cpu().numpy(), dic, keep_n=5)
    c = orc.compare_topk(ridx, rsc, res.simulation_indices, res.scores)
    assert c["tie_ok"] and c["scores_ok"], c


def _write_edax(path, version, pats, nav=None, is_hex=False, steps=(0.5, 0.25), extra=0):
    sy, sx = pats.shape[-2:]
    with open(path, "wb") as f:
        np.array([version], "uint32").tofile(f)
        if version == 1:
            np.array([sx, sy, 16], "uint32").tofile(f)
        else:
            np.array([sx, sy, 42], "uint32").tofile(f)
            np.array([1], "uint8").tofile(f)
            np.array([nav[1], nav[0]], "uint32").tofile(f)
            np.array([int(is_hex)], "uint8").tofile(f)
            np.array(steps, "float64").tofile(f)
        pats.ravel().tofile(f)
        if extra:
            np.zeros(extra * sy * sx, pats.dtype).tofile(f)


def test_load_edax_binary(tmp_path):
    rng = np.random.default_rng(1)
    p8 = rng.integers(0, 256, (6, 5, 7), dtype=np.uint8)
    p16 = rng.integers(0, 65536, (2, 3, 5, 7)).astype(np.uint16)
    _write_edax(tmp_path / "a.up1", 1, p8)
    a = kb.load_edax_binary(str(tmp_path / "a.up1"))
    assert a.data.dtype == np.uint8 and np.array_equal(a.data, p8) and a.step_sizes == (1, 1)
    assert np.array_equal(kb.load_edax_binary(str(tmp_path / "a.up1"), nav_shape=(2, 3)).data, p8.reshape(2, 3, 5, 7))
    with pytest.raises(ValueError, match="does not match the number of patterns"):
        kb.load_edax_binary(str(tmp_path / "a.up1"), nav_shape=(2, 2))
    _write_edax(tmp_path / "b.up2", 3, p16, nav=(2, 3))
    b = kb.load_edax_binary(str(tmp_path / "b.up2"))
    assert b.data.dtype == np.uint16 and np.array_equal(b.data, p16) and b.step_sizes == (0.25, 0.5)
    _write_edax(tmp_path / "c.up2", 3, p16, nav=(2, 3), is_hex=True, extra=1)
    with pytest.warns(UserWarning, match="hexagonal grid"):
        c = kb.load_edax_binary(str(tmp_path / "c.up2"))
    assert c.data.shape == (7, 5, 7) and np.array_equal(c.data[:6], p16.reshape(6, 5, 7)) and not c.data[6].any()
    _write_edax(tmp_path / "d.up1", 2, p8, nav=(1, 6))
    with pytest.raises(ValueError, match="not 2, can be read"):
        kb.load_edax_binary(str(tmp_path / "d.up1"))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_load_reference_edax_samples():
    folder = os.path.join(ref_loader.REFERENCE_ROOT, "src", "kikuchipy", "data", "edax_binary")
    ni = ref_loader.nickel_ebsd_small().reshape(9, 60, 60)
    up1 = kb.load_edax_binary(os.path.join(folder, "edax_binary.up1"))
    assert np.array_equal(up1.data, ni)
    with pytest.warns(UserWarning, match="hexagonal grid"):
        up2 = kb.load_edax_binary(os.path.join(folder, "edax_binary.up2"))
    assert up2.data.dtype == np.uint16 and up2.data.shape == (10, 60, 60) and np.array_equal(up2.data[:9], ni)
    assert np.allclose(up2.step_sizes, (np.pi / 2, np.pi))


@pytest.mark.gpu
def test_load_edax_to_device(tmp_path):
    rng = np.random.default_rng(2)
    p16 = rng.integers(0, 65536, (2, 3, 20, 24)).astype(np.uint16)
    _write_edax(tmp_path / "b.up2", 3, p16, nav=(2, 3))
    scan = kb.load_edax_binary(str(tmp_path / "b.up2"), device=True)
    assert scan.data.is_cuda and tuple(scan.data.shape) == (2, 3, 20, 24)
    assert np.array_equal(scan.data.cpu().numpy(), p16)
    from oracle import preprocess_oracle as pp

    out = kb.remove_dynamic_background(scan.data, "subtract", "spatial", std=2.5)
    assert np.array_equal(out.cpu().numpy(), pp.remove_dynamic_background(p16, "subtract", "spatial", 2.5))


def _write_oxford(path, pats, version, compressed=False, all_present=True, shuffle=True):
    """An .ebsp file like the reference's test fixture writes (/root/reference/conftest.py,
    ``oxford_binary_file``): pattern records stored in a rolled order, beam positions in um."""
    nr, nc, sr, sc = pats.shape
    n = nr * nc
    item = pats.dtype.itemsize
    header = 16 if version < 5 else 24
    footer = {0: 0, 1: 16}.get(version, 18)
    step = 1.0
    with open(path, "wb") as f:
        if version != 0:
            np.array(-version, dtype="<i8").tofile(f)
        start0 = (0 if version == 0 else 8) + (1 if version > 3 else 0)
        if version > 3:
            np.array(1, dtype="u1").tofile(f)
        starts = np.arange(n, dtype="<i8") * (header + sr * sc * item + footer) + start0 + n * 8
        order = np.arange(n)
        if shuffle:
            starts = np.roll(starts, 1)
            order = np.roll(order, -1)
        if not all_present:  # like the reference's fixture: the first listed point has no pattern
            starts[0] = 0
            order = order[1:]
        starts.tofile(f)
        for i in order:
            r, c = np.unravel_index(i, (nr, nc))
            if version >= 5:
                np.array([c, r], dtype="<i4").tofile(f)
            np.array([int(compressed), sr, sc, sr * sc * item], dtype="<i4").tofile(f)
            pats[r, c].tofile(f)
            if version == 1:
                np.array([c * step, r * step], dtype="<f8").tofile(f)
            elif version > 1:
                np.array(1, dtype=bool).tofile(f)
                np.array(c * step, dtype="<f8").tofile(f)
                np.array(1, dtype=bool).tofile(f)
                np.array(r * step, dtype="<f8").tofile(f)


@pytest.mark.parametrize("version, dtype, nav", [(2, np.uint8, (2, 3)), (1, np.uint16, (2, 3)), (0, np.uint8, (6,)),
                                                 (4, np.uint8, (2, 3)), (5, np.uint16, (2, 3)), (6, np.uint8, (2, 3))])
def test_load_oxford_binary_versions(tmp_path, version, dtype, nav):
    """tests/test_io/test_oxford_binary.py:59-101: versions 0, 1, > 1; uint8 and uint16."""
    rng = np.random.default_rng(version)
    pats = rng.integers(0, np.iinfo(dtype).max, (2, 3, 60, 60)).astype(dtype)
    _write_oxford(tmp_path / "p.ebsp", pats, version, shuffle=version != 0)
    scan = kb.load_oxford_binary(str(tmp_path / "p.ebsp"))
    assert scan.version == version and scan.data.dtype == dtype and scan.data.shape == nav + (60, 60)
    assert np.array_equal(scan.data.reshape(2, 3, 60, 60), pats)
    if version > 0:
        assert np.allclose(scan.original_metadata["beam_x"], [0, 1, 2, 0, 1, 2])
        assert np.allclose(scan.original_metadata["beam_y"], [0, 0, 0, 1, 1, 1])
    if version >= 5:
        assert np.array_equal(scan.original_metadata["map_x"], [0, 1, 2, 0, 1, 2])


def test_load_oxford_binary_missing_and_compressed(tmp_path):
    """tests/test_io/test_oxford_binary.py:47-57, :72-83."""
    pats = np.random.default_rng(0).integers(0, 256, (2, 3, 60, 60), dtype=np.uint8)
    _write_oxford(tmp_path / "c.ebsp", pats, 2, compressed=True)
    with pytest.raises(NotImplementedError, match="Cannot read compressed"):
        kb.load_oxford_binary(str(tmp_path / "c.ebsp"))
    _write_oxford(tmp_path / "m.ebsp", pats, 2, all_present=False)
    scan = kb.load_oxford_binary(str(tmp_path / "m.ebsp"))
    assert scan.data.shape == (5, 60, 60)  # one pattern is missing: a line of the present ones
    assert np.allclose(scan.original_metadata["beam_y"], [0, 1, 1, 1, 0])
    assert np.allclose(scan.original_metadata["beam_x"], [2, 0, 1, 2, 0])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_load_reference_oxford_sample():
    path = os.path.join(ref_loader.REFERENCE_ROOT, "src", "kikuchipy", "data", "oxford_binary", "patterns.ebsp")
    scan = kb.load_oxford_binary(path)
    assert scan.version == 2 and scan.step_sizes == (1.5, 1.5)
    assert np.array_equal(scan.data, ref_loader.nickel_ebsd_small())


def test_load_dispatches_on_extension(tmp_path):
    pats = np.random.default_rng(3).integers(0, 256, (2, 3, 60, 60), dtype=np.uint8)
    _write_oxford(tmp_path / "p.ebsp", pats, 2)
    _write_edax(tmp_path / "p.up1", 1, pats.reshape(6, 60, 60))
    pats.tofile(tmp_path / "p.dat")
    assert np.array_equal(kb.load(str(tmp_path / "p.ebsp")).data, pats)
    assert np.array_equal(kb.load(str(tmp_path / "p.up1"), nav_shape=(2, 3)).data, pats)
    with pytest.warns(UserWarning):
        assert np.array_equal(kb.load(str(tmp_path / "p.dat"), scan_size=(3, 2), pattern_size=(60, 60)).data, pats)
    with pytest.raises(IOError, match="No filename matches"):
        kb.load(str(tmp_path / "missing.dat"))
    (tmp_path / "p.h5").write_bytes(b"x")
    with pytest.raises(IOError, match="is not an HDF5 file"):
        kb.load(str(tmp_path / "p.h5"))
    kb.save_h5ebsd(str(tmp_path / "q.h5"), pats)
    assert np.array_equal(kb.load(str(tmp_path / "q.h5")).data, pats)
    (tmp_path / "p.tif").write_bytes(b"x")
    with pytest.raises(IOError, match="only .dat, .up1"):
        kb.load(str(tmp_path / "p.tif"))
