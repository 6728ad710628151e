"""Minimal NIfTI-1 single-file reader / writer (gzip + struct; nibabel is not a dependency).

Covers what the reference's drivers do with nibabel around the solvers
(synthetic_gens/generatePatient_FK_2c.py:55-65: `nib.load(path).get_fdata()` and
`nib.save(nib.Nifti1Image(array, np.eye(4)), path)`): 3-D / 4-D volumes of the common scalar
types, scl_slope / scl_inter applied on load, identity affine written on save.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def load(path, raw=False):
    """Returns the volume as float64 with scl_slope/scl_inter applied (nibabel's get_fdata()), or the
    stored dtype when raw=True.  Array order is (x, y, z[, t]) like nibabel's."""
    with _open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 352:
        raise ValueError(f"{path}: too short for a NIfTI-1 header")
    endian = "<"
    if struct.unpack("<i", buf[0:4])[0] != 348:
        endian = ">"
        if struct.unpack(">i", buf[0:4])[0] != 348:
            raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if buf[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI-1 magic")
    dim = struct.unpack(endian + "8h", buf[40:56])
    datatype = struct.unpack(endian + "h", buf[70:72])[0]
    vox_offset, slope, inter = struct.unpack(endian + "3f", buf[108:120])
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    n = int(np.prod(shape))
    a = np.frombuffer(buf, dtype=dt, count=n, offset=int(vox_offset)).reshape(shape, order="F")
    if raw:
        return a
    out = a.astype(np.float64)
    if slope != 0.0 and not (slope == 1.0 and inter == 0.0) and np.isfinite(slope):
        out = out * float(slope) + float(inter)
    return out


def save(array, path, voxel_size=(1.0, 1.0, 1.0)):
    """Writes `array` with an identity affine (sform), like nib.Nifti1Image(array, np.eye(4)).
    int64 is narrowed to int32 as the reference's save_nifti does (generatePatient_FK_2c.py:57-61);
    bool becomes uint8."""
    a = np.asarray(array)
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    key = a.dtype.str[1:]
    if key not in _CODES:
        raise ValueError(f"unsupported dtype {a.dtype}")
    if not (1 <= a.ndim <= 7):
        raise ValueError("NIfTI-1 stores 1 to 7 dimensions")
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    pixdim = [1.0] + [float(v) for v in voxel_size][:3] + [1.0] * 4
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, _CODES[key])
    struct.pack_into("<h", hdr, 72, a.dtype.itemsize * 8)          # bitpix
    struct.pack_into("<8f", hdr, 76, *pixdim[:8])
    struct.pack_into("<f", hdr, 108, 352.0)                        # vox_offset
    struct.pack_into("<f", hdr, 112, 1.0)                          # scl_slope
    struct.pack_into("<f", hdr, 116, 0.0)                          # scl_inter
    hdr[123] = 2                                                   # xyzt_units: mm
    struct.pack_into("<h", hdr, 252, 0)                            # qform_code
    struct.pack_into("<h", hdr, 254, 2)                            # sform_code: aligned
    struct.pack_into("<4f", hdr, 280, pixdim[1], 0.0, 0.0, 0.0)    # srow_x
    struct.pack_into("<4f", hdr, 296, 0.0, pixdim[2], 0.0, 0.0)    # srow_y
    struct.pack_into("<4f", hdr, 312, 0.0, 0.0, pixdim[3], 0.0)    # srow_z
    hdr[344:348] = b"n+1\x00"
    data = np.asarray(a, dtype=a.dtype.newbyteorder("<")).tobytes(order="F")
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(b"\x00\x00\x00\x00")                               # no header extensions
        f.write(data)
