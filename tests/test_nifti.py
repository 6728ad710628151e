"""NIfTI-1 reader / writer used by tools/generate_patient.py (CPU only)."""
import gzip
import os
import struct

import numpy as np
import pytest

from tumorgrowthtoolkit_b200 import nifti


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32, np.int16, np.uint8])
@pytest.mark.parametrize("ext", [".nii", ".nii.gz"])
def test_roundtrip(tmp_path, dtype, ext):
    rng = np.random.default_rng(0)
    a = (rng.uniform(0, 100, (5, 7, 3))).astype(dtype)
    p = str(tmp_path / ("v" + ext))
    nifti.save(a, p)
    assert np.array_equal(nifti.load(p, raw=True), a)
    assert nifti.load(p).dtype == np.float64 and np.array_equal(nifti.load(p), a.astype(np.float64))


def test_header_fields_and_int64_narrowing(tmp_path):
    a = np.arange(24, dtype=np.int64).reshape(2, 3, 4)
    p = str(tmp_path / "seg.nii.gz")
    nifti.save(a, p)
    raw = gzip.open(p, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:347] == b"n+1"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 2, 3, 4)
    assert struct.unpack("<h", raw[70:72])[0] == 8                     # int32, like the reference's save_nifti
    assert struct.unpack("<4f", raw[280:296]) == (1.0, 0.0, 0.0, 0.0)  # identity sform
    b = nifti.load(p, raw=True)
    assert b.dtype == np.int32 and np.array_equal(b, a)
    # Fortran order on disk: the first stored voxels run along x
    first = np.frombuffer(raw, dtype="<i4", count=2, offset=352)
    assert list(first) == [int(a[0, 0, 0]), int(a[1, 0, 0])]


def test_scl_slope_and_inter_applied(tmp_path):
    a = np.arange(8, dtype=np.int16).reshape(2, 2, 2)
    p = str(tmp_path / "s.nii")
    nifti.save(a, p)
    raw = bytearray(open(p, "rb").read())
    struct.pack_into("<f", raw, 112, 0.5)
    struct.pack_into("<f", raw, 116, 2.0)
    open(p, "wb").write(bytes(raw))
    assert np.array_equal(nifti.load(p), a * 0.5 + 2.0)


def test_rejects_garbage(tmp_path):
    p = str(tmp_path / "x.nii")
    open(p, "wb").write(b"\x00" * 400)
    with pytest.raises(ValueError):
        nifti.load(p)


@pytest.mark.skipif(not os.path.exists("/root/reference/dataset/WM.nii.gz"), reason="reference dataset not present")
def test_reads_the_reference_atlas_files(atlas_labels):
    wm = nifti.load("/root/reference/dataset/WM.nii.gz")
    assert wm.shape == (193, 229, 193) and 0.0 <= wm.min() and wm.max() <= 1.0 + 1e-6     # scl_slope/inter applied
    seg = nifti.load("/root/reference/dataset/sub-mni152_tissue-with-antsN4_space-sri.nii.gz")
    assert np.array_equal(np.rint(seg).astype(np.uint8), atlas_labels)
