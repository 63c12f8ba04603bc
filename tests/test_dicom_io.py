"""Host-side DICOM I/O of the volume driver (ducosy_gan_b200/dicom_io.py).  pydicom is absent here, so the
files are assembled by hand byte by byte following PS3.5/PS3.10 ("parity unpinned" against pydicom); what is pinned is the
behaviour generate.py asks of the files: stored values, Rows/Columns, RescaleSlope/Intercept defaults, the tags synthesis()
rewrites (generate.py:264-287) and that every other element survives untouched."""
import os
import struct

import numpy as np
import pytest

from ducosy_gan_b200 import dicom_io as dio


def _el_explicit(g, e, vr, value):
    if len(value) % 2:
        value += b" "
    if vr in (b"OB", b"OW", b"SQ", b"UN", b"UT"):
        return struct.pack("<HH2sHI", g, e, vr, 0, len(value)) + value
    return struct.pack("<HH2sH", g, e, vr, len(value)) + value


def _el_implicit(g, e, value, length=None):
    if len(value) % 2 and length is None:
        value += b" "
    return struct.pack("<HHI", g, e, len(value) if length is None else length) + value


def _file(px, explicit=True, signed=True, slope="1", intercept="-1024", with_rescale=True, series="NCCT chest"):
    ts = dio.EXPLICIT_LE if explicit else dio.IMPLICIT_LE
    meta = _el_explicit(0x0002, 0x0002, b"UI", b"1.2.840.10008.5.1.4.1.1.2\0") + _el_explicit(0x0002, 0x0010, b"UI", ts.encode() + (b"\0" if len(ts) % 2 else b""))
    head = _el_explicit(0x0002, 0x0000, b"UL", struct.pack("<I", len(meta)))
    E = (lambda g, e, vr, v: _el_explicit(g, e, vr, v)) if explicit else (lambda g, e, vr, v: _el_implicit(g, e, v))
    # a sequence of undefined length with one undefined-length item holding one element (nested encoding follows the file's)
    inner = E(0x0008, 0x0100, b"SH", b"CODE")
    item = struct.pack("<HHI", 0xFFFE, 0xE000, 0xFFFFFFFF) + inner + struct.pack("<HHI", 0xFFFE, 0xE00D, 0)
    seq_val = item + struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
    seq = (struct.pack("<HH2sHI", 0x0008, 0x1140, b"SQ", 0, 0xFFFFFFFF) if explicit else struct.pack("<HHI", 0x0008, 0x1140, 0xFFFFFFFF)) + seq_val
    body = E(0x0008, 0x0060, b"CS", b"CT") + E(0x0008, 0x103E, b"LO", series.encode()) + seq + E(0x0009, 0x1001, b"UN" if explicit else b"", b"private!")
    body += E(0x0010, 0x0010, b"PN", b"Anon^Patient") + E(0x0020, 0x0013, b"IS", b"7")
    body += E(0x0028, 0x0010, b"US", struct.pack("<H", px.shape[0])) + E(0x0028, 0x0011, b"US", struct.pack("<H", px.shape[1]))
    body += E(0x0028, 0x0100, b"US", struct.pack("<H", 16)) + E(0x0028, 0x0103, b"US", struct.pack("<H", 1 if signed else 0))
    body += E(0x0028, 0x1050, b"DS", b"40") + E(0x0028, 0x1051, b"DS", b"400")
    if with_rescale:
        body += E(0x0028, 0x1052, b"DS", intercept.encode()) + E(0x0028, 0x1053, b"DS", slope.encode())
    body += E(0x7FE0, 0x0010, b"OW", px.astype("<i2" if signed else "<u2").tobytes())
    return b"\0" * 128 + b"DICM" + head + meta + body


@pytest.mark.parametrize("explicit", [True, False])
@pytest.mark.parametrize("signed", [True, False])
def test_read_write_round_trip(tmp_path, explicit, signed):
    rng = np.random.Generator(np.random.PCG64(1))
    px = rng.integers(0 if not signed else -1000, 3000, size=(16, 24)).astype(np.int16)
    p = tmp_path / "a.dcm"
    p.write_bytes(_file(px, explicit, signed))
    ds = dio.read_dicom(str(p))
    assert (ds.rows, ds.cols, ds.pixel_representation) == (16, 24, 1 if signed else 0)
    assert (ds.slope, ds.intercept, ds.series_description) == (1.0, -1024.0, "NCCT chest")
    assert ds.transfer_syntax == (dio.EXPLICIT_LE if explicit else dio.IMPLICIT_LE)
    assert np.array_equal(ds.pixel_array().astype(np.int64), px.astype(np.int64))
    new = (px // 2 + 5).astype(np.int16)
    out = tmp_path / "o.dcm"
    dio.write_dicom(str(out), ds, new.view(np.uint16) if not signed else new)
    back = dio.read_dicom(str(out))
    assert back.transfer_syntax == dio.EXPLICIT_LE                               # generate.py:110
    assert np.array_equal(back.pixel_array().astype(np.int64), new.astype(np.int64))
    assert back.series_description == "DuCoSyGAN sCECT v2"                       # generate.py:283
    fmt = "<h" if signed else "<H"
    assert struct.unpack(fmt, back.find((0x0028, 0x0106)).value)[0] == int(new.min())   # generate.py:273-278
    assert struct.unpack(fmt, back.find((0x0028, 0x0107)).value)[0] == int(new.max())
    assert back.find((0x0028, 0x0106)).vr == (b"SS" if signed else b"US")
    assert back.find((0x0028, 0x1050)).value.strip() == b"-375.0" and back.find((0x0028, 0x1051)).value.strip() == b"1250"
    # everything else is carried over: same tags in order, same values
    keep = lambda d: [(e.tag, e.value.rstrip(b" \0")) for e in d.elements if e.tag not in
                      ((0x0008, 0x103E), (0x0028, 0x0106), (0x0028, 0x0107), (0x0028, 0x1050), (0x0028, 0x1051), (0x7FE0, 0x0010))]
    assert keep(back) == keep(ds)
    tags = [e.tag for e in back.elements]
    assert tags == sorted(tags)
    seq = back.find((0x0008, 0x1140))
    assert seq.length == 0xFFFFFFFF and seq.vr == (b"SQ" if explicit else b"UN")
    assert back.find((0x0010, 0x0010)).vr == b"PN" and back.find((0x0009, 0x1001)).vr == b"UN"
    glen = struct.unpack("<I", back.meta[0].value)[0]
    assert back.meta[0].tag == (0x0002, 0x0000) and glen == sum(8 + len(e.value) + (4 if e.vr in dio._LONG_VRS else 0) for e in back.meta[1:])


def test_rescale_defaults_and_errors(tmp_path):
    px = np.zeros((8, 8), np.int16)
    p = tmp_path / "n.dcm"
    p.write_bytes(_file(px, with_rescale=False))
    ds = dio.read_dicom(str(p))
    assert (ds.slope, ds.intercept) == (1.0, 0.0)                                # generate.py:140-145 defaults
    (tmp_path / "bad.dcm").write_bytes(b"not dicom" * 40)
    with pytest.raises(dio.DicomError):
        dio.read_dicom(str(tmp_path / "bad.dcm"))
    jpeg = _file(px).replace(dio.EXPLICIT_LE.encode() + b"\0", b"1.2.840.10008.1.2.4.70")
    (tmp_path / "j.dcm").write_bytes(jpeg)
    with pytest.raises(dio.DicomError):
        dio.read_dicom(str(tmp_path / "j.dcm"))


def test_read_series_orders_by_filename_and_write_series_names(tmp_path):
    src = tmp_path / "POST VUE"
    src.mkdir()
    vols = []
    for i, name in enumerate(["IM0003.dcm", "IM0001.dcm", "IM0002.dcm"]):
        px = np.full((8, 8), int(name[5]) * 100, np.int16)
        (src / name).write_bytes(_file(px))
    vol, slices = dio.read_series(str(src), workers=2)
    assert [os.path.basename(s.path) for s in slices] == ["IM0001.dcm", "IM0002.dcm", "IM0003.dcm"]   # sorted(glob) generate.py:88
    assert vol.shape == (3, 8, 8) and [int(vol[i, 0, 0]) for i in range(3)] == [100, 200, 300]
    out = tmp_path / "out"
    dio.write_series(slices, vol + 1, str(out), workers=2)
    assert sorted(os.listdir(out)) == ["0000.dcm", "0001.dcm", "0002.dcm"]                              # generate.py:285
    assert int(dio.read_dicom(str(out / "0002.dcm")).pixel_array()[0, 0]) == 301


class _FakeSynth:
    """Stands in for DualHUSynthesizer on the CPU: records the (slope, intercept) groups it is called with."""
    device = "cpu"

    def __init__(self):
        self.calls = []

    def synthesize_volume(self, vol, slope, intercept, postprocess=False):
        self.calls.append((tuple(vol.shape), slope, intercept))
        return vol + 7


def test_synthesize_series_groups_slices_by_rescale_and_rejects_other_sizes(tmp_path):
    import torch
    src = tmp_path / "s"
    src.mkdir()
    px = np.zeros((512, 512), np.int16)
    for i, (sl, ic) in enumerate([("1", "-1024"), ("1", "-1024"), ("1", "-1000"), ("2", "-1000"), ("2", "-1000")]):
        (src / f"{i:03d}.dcm").write_bytes(_file(px + i, slope=sl, intercept=ic))
    fake = _FakeSynth()
    merged, slices = dio.synthesize_series(fake, str(src), str(tmp_path / "o"), postprocess=False, workers=2)
    assert fake.calls == [((2, 512, 512), 1.0, -1024.0), ((1, 512, 512), 1.0, -1000.0), ((2, 512, 512), 2.0, -1000.0)]
    assert torch.equal(merged[:, 0, 0], torch.arange(5, dtype=torch.int16) + 7)
    assert int(dio.read_dicom(str(tmp_path / "o" / "0004.dcm")).pixel_array()[5, 5]) == 11
    small = tmp_path / "small"
    small.mkdir()
    (small / "a.dcm").write_bytes(_file(np.zeros((8, 8), np.int16)))
    with pytest.raises(dio.DicomError):
        dio.synthesize_series(fake, str(small), str(tmp_path / "o2"), postprocess=False)


def test_read_series_rejects_unsigned_values_beyond_int16(tmp_path):
    src = tmp_path / "u"
    src.mkdir()
    px = np.zeros((8, 8), np.uint16)
    px[0, 0] = 40000
    raw = _file(px.view(np.int16), signed=False)
    (src / "a.dcm").write_bytes(raw)
    with pytest.raises(dio.DicomError):
        dio.read_series(str(src))
