"""Volume-at-a-time DICOM driver: what the reference's ``generate()`` + ``synthesis()`` do per slice through
three intermediate DICOM folders (generate.py:57-126,190-292: 2 dcmread + 2 deepcopy + 2 save_as per slice, then 3 more reads and
1 write), done once per series: read every slice of a series into ONE pinned int16 volume -> ``DualHUSynthesizer`` (GPU) ->
write every merged slice once.

pydicom is not available in this image, so the small part of DICOM the path needs is implemented here (PARITY UNPINNED against
pydicom: no reference fixture exists; covered by round-trip and hand-assembled byte tests):
  * Part-10 files, uncompressed little-endian transfer syntaxes (Implicit VR LE 1.2.840.10008.1.2, Explicit VR LE ...1.2.1);
    compressed / big-endian syntaxes raise.
  * read: Rows, Columns, BitsAllocated, PixelRepresentation, RescaleSlope / RescaleIntercept (defaults 1 / 0 exactly as
    generate.py:140-145), SeriesDescription, PixelData; every other element is carried as raw bytes.
  * write (generate.py:264-287): Explicit VR LE like ``output_dcm.file_meta.TransferSyntaxUID = ExplicitVRLittleEndian``
    (generate.py:110); PixelData replaced; SmallestImagePixelValue / LargestImagePixelValue (0028,0106/0107, VR US or SS by
    PixelRepresentation), WindowWidth 1250 / WindowCenter -375.0, SeriesDescription "DuCoSyGAN sCECT v2"; files named
    ``{idx:04d}.dcm``.  Elements of an implicit-VR source whose VR is not in the small table below are written as UN
    (the standard's rule for unknown VRs), sequences keep their implicit-VR item encoding under UN with their original length.
Slices whose size is not the generators' 512 x 512 raise: the reference resizes them with torchvision (generate.py:52,94-100),
which this driver does not reproduce.
"""
from __future__ import annotations

import glob
import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np
import torch

IMPLICIT_LE = "1.2.840.10008.1.2"
EXPLICIT_LE = "1.2.840.10008.1.2.1"
_LONG_VRS = {b"OB", b"OD", b"OF", b"OL", b"OV", b"OW", b"SQ", b"UC", b"UN", b"UR", b"UT", b"SV", b"UV"}
_UNDEF = 0xFFFFFFFF
_ITEM, _ITEM_END, _SEQ_END = (0xFFFE, 0xE000), (0xFFFE, 0xE00D), (0xFFFE, 0xE0DD)
# VRs of the tags a CT image usually carries (used only when converting an implicit-VR source to explicit VR)
_VR = {
    (0x0008, 0x0005): b"CS", (0x0008, 0x0008): b"CS", (0x0008, 0x0016): b"UI", (0x0008, 0x0018): b"UI", (0x0008, 0x0020): b"DA",
    (0x0008, 0x0021): b"DA", (0x0008, 0x0022): b"DA", (0x0008, 0x0023): b"DA", (0x0008, 0x0030): b"TM", (0x0008, 0x0031): b"TM",
    (0x0008, 0x0032): b"TM", (0x0008, 0x0033): b"TM", (0x0008, 0x0050): b"SH", (0x0008, 0x0060): b"CS", (0x0008, 0x0070): b"LO",
    (0x0008, 0x0080): b"LO", (0x0008, 0x0090): b"PN", (0x0008, 0x1030): b"LO", (0x0008, 0x103E): b"LO", (0x0008, 0x1090): b"LO",
    (0x0010, 0x0010): b"PN", (0x0010, 0x0020): b"LO", (0x0010, 0x0030): b"DA", (0x0010, 0x0040): b"CS", (0x0018, 0x0015): b"CS",
    (0x0018, 0x0050): b"DS", (0x0018, 0x0060): b"DS", (0x0018, 0x1030): b"LO", (0x0018, 0x1100): b"DS", (0x0018, 0x1120): b"DS",
    (0x0018, 0x1151): b"IS", (0x0018, 0x1210): b"SH", (0x0018, 0x5100): b"CS", (0x0020, 0x000D): b"UI", (0x0020, 0x000E): b"UI",
    (0x0020, 0x0010): b"SH", (0x0020, 0x0011): b"IS", (0x0020, 0x0012): b"IS", (0x0020, 0x0013): b"IS", (0x0020, 0x0032): b"DS",
    (0x0020, 0x0037): b"DS", (0x0020, 0x0052): b"UI", (0x0020, 0x1041): b"DS", (0x0028, 0x0002): b"US", (0x0028, 0x0004): b"CS",
    (0x0028, 0x0010): b"US", (0x0028, 0x0011): b"US", (0x0028, 0x0030): b"DS", (0x0028, 0x0100): b"US", (0x0028, 0x0101): b"US",
    (0x0028, 0x0102): b"US", (0x0028, 0x0103): b"US", (0x0028, 0x1050): b"DS", (0x0028, 0x1051): b"DS", (0x0028, 0x1052): b"DS",
    (0x0028, 0x1053): b"DS", (0x0028, 0x1054): b"LO", (0x7FE0, 0x0010): b"OW",
}


class DicomError(RuntimeError):
    pass


@dataclass
class Element:
    tag: tuple
    vr: bytes | None            # None: implicit-VR source, VR unknown
    length: int                 # value length as encoded (may be 0xFFFFFFFF)
    value: bytes                # value bytes (for undefined length: everything up to and including the sequence delimiter)


@dataclass
class DicomSlice:
    path: str
    preamble: bytes
    meta: list = field(default_factory=list)        # file-meta elements (group 0002, explicit VR)
    elements: list = field(default_factory=list)    # top-level dataset elements in file order, PixelData included
    transfer_syntax: str = EXPLICIT_LE
    rows: int = 0
    cols: int = 0
    bits_allocated: int = 16
    pixel_representation: int = 1
    slope: float = 1.0
    intercept: float = 0.0
    series_description: str = ""

    def find(self, tag):
        for e in self.elements:
            if e.tag == tag:
                return e
        return None

    @property
    def pixel_dtype(self):
        return np.dtype("<i2") if self.pixel_representation == 1 else np.dtype("<u2")

    def pixel_array(self) -> np.ndarray:
        e = self.find((0x7FE0, 0x0010))
        if e is None:
            raise DicomError(f"{self.path}: no PixelData")
        need = self.rows * self.cols * 2
        if e.length == _UNDEF or len(e.value) < need:
            raise DicomError(f"{self.path}: PixelData is encapsulated or short ({len(e.value)} < {need} bytes)")
        return np.frombuffer(e.value, dtype=self.pixel_dtype, count=self.rows * self.cols).reshape(self.rows, self.cols)


# ------------------------------------------------------------------ parsing
def _skip_items(buf, pos, explicit):
    """pos: first byte after an undefined-length element header -> position after its sequence delimitation item."""
    n = len(buf)
    while pos + 8 <= n:
        g, e, ln = struct.unpack_from("<HHI", buf, pos)
        pos += 8
        if (g, e) == _SEQ_END:
            return pos
        if (g, e) != _ITEM:
            raise DicomError(f"unexpected tag ({g:04X},{e:04X}) inside a sequence")
        if ln != _UNDEF:
            pos += ln
            continue
        while True:      # undefined-length item: nested data set up to the item delimitation item
            if pos + 8 > n:
                raise DicomError("truncated sequence item")
            g2, e2 = struct.unpack_from("<HH", buf, pos)
            if (g2, e2) == _ITEM_END:
                pos += 8
                break
            _, vr2, ln2, hdr = _read_header(buf, pos, explicit)
            pos += hdr
            pos = _skip_items(buf, pos, explicit and vr2 != b"UN") if ln2 == _UNDEF else pos + ln2
    raise DicomError("sequence without a delimiter")


def _read_header(buf, pos, explicit):
    """-> (tag, vr | None, length, header bytes)."""
    g, e = struct.unpack_from("<HH", buf, pos)
    if explicit and g != 0xFFFE:
        vr = bytes(buf[pos + 4:pos + 6])
        if vr in _LONG_VRS:
            return (g, e), vr, struct.unpack_from("<I", buf, pos + 8)[0], 12
        return (g, e), vr, struct.unpack_from("<H", buf, pos + 6)[0], 8
    return (g, e), None, struct.unpack_from("<I", buf, pos + 4)[0], 8


def _parse_elements(buf, pos, end, explicit, stop_group=None):
    out = []
    while pos + 8 <= end:
        tag, vr, ln, hdr = _read_header(buf, pos, explicit)
        if stop_group is not None and tag[0] != stop_group:
            break
        pos += hdr
        if ln == _UNDEF:
            stop = _skip_items(buf, pos, explicit and vr != b"UN")    # UN of undefined length = implicit-VR encoded sequence (PS3.5 6.2.2)
        else:
            stop = pos + ln
            if stop > end:
                raise DicomError(f"element ({tag[0]:04X},{tag[1]:04X}) runs past the end of the file")
        out.append(Element(tag, vr, ln, bytes(buf[pos:stop])))
        pos = stop
    return out, pos


def _text(e):
    return e.value.decode("ascii", errors="replace").strip(" \0")


def read_dicom(path: str) -> DicomSlice:
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 132 or buf[128:132] != b"DICM":
        raise DicomError(f"{path}: not a DICOM Part-10 file (no DICM prefix)")
    ds = DicomSlice(path=path, preamble=buf[:128])
    ds.meta, pos = _parse_elements(buf, 132, len(buf), True, stop_group=0x0002)
    for e in ds.meta:
        if e.tag == (0x0002, 0x0010):
            ds.transfer_syntax = _text(e)
    if ds.transfer_syntax not in (IMPLICIT_LE, EXPLICIT_LE):
        raise DicomError(f"{path}: transfer syntax {ds.transfer_syntax} is not supported (uncompressed little endian only)")
    ds.elements, _ = _parse_elements(buf, pos, len(buf), ds.transfer_syntax == EXPLICIT_LE)
    us = lambda e: struct.unpack_from("<H", e.value)[0]
    for e in ds.elements:
        if e.tag == (0x0028, 0x0010): ds.rows = us(e)
        elif e.tag == (0x0028, 0x0011): ds.cols = us(e)
        elif e.tag == (0x0028, 0x0100): ds.bits_allocated = us(e)
        elif e.tag == (0x0028, 0x0103): ds.pixel_representation = us(e)
        elif e.tag == (0x0028, 0x1052): ds.intercept = float(_text(e).split("\\")[0])
        elif e.tag == (0x0028, 0x1053): ds.slope = float(_text(e).split("\\")[0])
        elif e.tag == (0x0008, 0x103E): ds.series_description = _text(e)
    if ds.bits_allocated != 16:
        raise DicomError(f"{path}: BitsAllocated {ds.bits_allocated} unsupported (16-bit CT only)")
    return ds


# ------------------------------------------------------------------ writing
def _pad(value: bytes, vr: bytes) -> bytes:
    if len(value) % 2:
        value += b"\0" if vr in (b"UI", b"OB", b"UN", b"OW") else b" "
    return value


def _encode(tag, vr, value, length=None):
    """One element in Explicit VR LE; ``length`` overrides len(value) (undefined-length elements carried over)."""
    ln = len(value) if length is None else length
    if vr in _LONG_VRS:
        return struct.pack("<HH2sHI", tag[0], tag[1], vr, 0, ln) + value
    if ln > 0xFFFF:
        raise DicomError(f"element ({tag[0]:04X},{tag[1]:04X}) too long for VR {vr!r}")
    return struct.pack("<HH2sH", tag[0], tag[1], vr, ln) + value


def write_dicom(path: str, src: DicomSlice, pixels: np.ndarray, series_description="DuCoSyGAN sCECT v2", window_width=1250,
                window_center=-375.0):
    """generate.py:264-287 for one slice: ``src`` with PixelData := pixels and the tags the reference updates."""
    if pixels.shape != (src.rows, src.cols):
        raise DicomError(f"pixel array {pixels.shape} does not match Rows x Columns {src.rows} x {src.cols}")
    px = np.ascontiguousarray(pixels.astype(src.pixel_dtype, copy=False))
    mm_vr, mm_fmt = (b"US", "<H") if src.pixel_representation == 0 else (b"SS", "<h")
    replace = {
        (0x0008, 0x103E): (b"LO", _pad(series_description.encode("ascii"), b"LO")),
        (0x0028, 0x0106): (mm_vr, struct.pack(mm_fmt, int(px.min()))),
        (0x0028, 0x0107): (mm_vr, struct.pack(mm_fmt, int(px.max()))),
        (0x0028, 0x1050): (b"DS", _pad(str(window_center).encode("ascii"), b"DS")),
        (0x0028, 0x1051): (b"DS", _pad(str(window_width).encode("ascii"), b"DS")),
        (0x7FE0, 0x0010): (b"OW", px.tobytes()),
    }
    body = []
    pending = sorted(replace)
    for e in src.elements:
        while pending and pending[0] < e.tag:          # elements the source does not carry are inserted in tag order
            t = pending.pop(0)
            body.append(_encode(t, *replace[t]))
        if e.tag in replace:
            if pending and pending[0] == e.tag:
                pending.pop(0)
            body.append(_encode(e.tag, *replace[e.tag]))
            continue
        vr = e.vr
        if vr is None:                                  # implicit-VR source
            vr = _VR.get(e.tag, b"UN")
            if e.length == _UNDEF or (vr not in _LONG_VRS and len(e.value) > 0xFFFF):
                vr = b"UN"                              # sequences keep their implicit-VR items under UN
        body.append(_encode(e.tag, vr, e.value if e.length == _UNDEF else _pad(e.value, vr), e.length if e.length == _UNDEF else None))
    for t in pending:
        body.append(_encode(t, *replace[t]))
    meta = [e for e in src.meta if e.tag not in ((0x0002, 0x0000), (0x0002, 0x0010))]
    meta.append(Element((0x0002, 0x0010), b"UI", 0, _pad(EXPLICIT_LE.encode("ascii"), b"UI")))
    meta.sort(key=lambda e: e.tag)
    meta_bytes = b"".join(_encode(e.tag, e.vr, e.value) for e in meta)
    head = _encode((0x0002, 0x0000), b"UL", struct.pack("<I", len(meta_bytes)))
    with open(path, "wb") as f:
        f.write(src.preamble + b"DICM" + head + meta_bytes + b"".join(body))


# ------------------------------------------------------------------ series level
def read_series(folder: str, workers: int = 8):
    """``sorted(glob(folder/*.dcm))`` (generate.py:88) -> (pinned int16 volume [S,H,W], list of DicomSlice)."""
    paths = sorted(glob.glob(os.path.join(folder, "*.dcm")))
    if not paths:
        raise DicomError(f"{folder}: no .dcm files")
    with ThreadPoolExecutor(max_workers=workers) as pool:
        slices = list(pool.map(read_dicom, paths))
    H, W = slices[0].rows, slices[0].cols
    pin = torch.cuda.is_available()
    vol = torch.empty((len(slices), H, W), dtype=torch.int16, pin_memory=pin)
    dst = vol.numpy()
    for i, s in enumerate(slices):
        if (s.rows, s.cols) != (H, W):
            raise DicomError(f"{s.path}: slice size {s.rows}x{s.cols} differs from the series' {H}x{W}")
        a = s.pixel_array()
        if s.pixel_representation == 0 and a.max(initial=0) > 32767:
            raise DicomError(f"{s.path}: unsigned stored values above 32767 cannot go through the int16 path")
        dst[i] = a.view(np.int16) if a.dtype != np.int16 else a
    return vol, slices


def write_series(slices, volume, out_dir: str, workers: int = 8, **tags):
    """Merged volume -> ``out_dir/{idx:04d}.dcm`` (generate.py:264-287)."""
    os.makedirs(out_dir, exist_ok=True)
    vol = volume.cpu().numpy() if torch.is_tensor(volume) else np.asarray(volume)
    if len(slices) != vol.shape[0]:
        raise DicomError("write_series: one DicomSlice per volume slice is required")

    def one(i):
        px = vol[i]
        if slices[i].pixel_representation == 0:
            px = px.view(np.uint16)
        write_dicom(os.path.join(out_dir, f"{i:04d}.dcm"), slices[i], px, **tags)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(one, range(len(slices))))


def synthesize_series(synthesizer, ncct_folder: str, out_dir: str, postprocess: bool = True, workers: int = 8):
    """One patient of generate.py's ``generate()`` + ``synthesis()``: NCCT series folder -> synthetic CECT series folder.
    Slices are grouped by (RescaleSlope, RescaleIntercept) -- one synthesizer call per run of equal values (a CT series has
    one); ``postprocess`` applies the volume smoothing of generate.py:254-263 on the device."""
    vol, slices = read_series(ncct_folder, workers)
    if vol.shape[1:] != (512, 512):
        raise DicomError(f"{ncct_folder}: {vol.shape[1]}x{vol.shape[2]} slices; the generators take 512x512 (the reference's "
                         "torchvision resize, generate.py:52,94-100, is not reproduced)")
    merged = torch.empty_like(vol)
    lo = 0
    while lo < len(slices):
        hi = lo + 1
        key = (slices[lo].slope, slices[lo].intercept)
        while hi < len(slices) and (slices[hi].slope, slices[hi].intercept) == key:
            hi += 1
        merged[lo:hi] = synthesizer.synthesize_volume(vol[lo:hi], key[0], key[1], postprocess=False)
        lo = hi
    if postprocess:
        from .postprocess import postprocess_volume
        dev = synthesizer.device
        merged = postprocess_volume(merged.to(dev)).cpu()
    write_series(slices, merged, out_dir, workers)
    return merged, slices
