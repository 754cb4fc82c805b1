"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader and writer — just enough for the reference's
data plane (utils/mri_dataset.py:28,124-127 read volumes with nibabel's get_fdata(); eval.py:51-57
writes the argmax label volume as float32 with an identity affine).  nibabel is not available in
this image, and the hot path keeps volumes resident in HBM anyway, so host I/O is a thin numpy
layer.  Header layout: NIfTI-1 specification (348-byte header, vox_offset 352)."""
from __future__ import annotations

import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load(path) -> np.ndarray:
    """Return the volume as float64 with scl_slope / scl_inter applied (== nibabel get_fdata())."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: too short for a NIfTI-1 header")
    end = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(end + "i", raw[:4])[0] != 348:
        raise ValueError(f"{path}: sizeof_hdr != 348 (not NIfTI-1)")
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype = struct.unpack(end + "h", raw[70:72])[0]
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    slope, inter = struct.unpack(end + "2f", raw[112:120])
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=max(vox_offset, 352)).reshape(shape, order="F")
    out = data.astype(np.float64)
    if slope not in (0.0, 1.0) or inter != 0.0:
        if slope != 0.0 and np.isfinite(slope):
            out = out * slope + inter
    while out.ndim > 3 and out.shape[-1] == 1:
        out = out[..., 0]
    return np.ascontiguousarray(out)


def save(path, vol: np.ndarray, affine=None) -> None:
    """Write `vol` (cast to float32 like eval.py:54) with the given 4x4 affine (identity default)."""
    vol = np.asarray(vol, dtype=np.float32)
    aff = np.eye(4, dtype=np.float32) if affine is None else np.asarray(affine, dtype=np.float32)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [vol.ndim] + list(vol.shape) + [1] * (7 - vol.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, 16)            # datatype float32
    struct.pack_into("<h", hdr, 72, 32)            # bitpix
    struct.pack_into("<8f", hdr, 76, 1.0, *([1.0] * 7))   # pixdim
    struct.pack_into("<f", hdr, 108, 352.0)        # vox_offset
    struct.pack_into("<2f", hdr, 112, 1.0, 0.0)    # scl_slope, scl_inter
    struct.pack_into("<h", hdr, 254, 1)            # sform_code = 1 (scanner)
    for r in range(3):
        struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(v) for v in aff[r]])
    hdr[344:348] = b"n+1\x00"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\x00" * 4)
        f.write(np.asfortranarray(vol).tobytes(order="F"))
