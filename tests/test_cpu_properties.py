"""Property tests (hypothesis) of the host-side logic around the hot path and of the oracle's data plane: slice
sharding, cube padding (mri_dataset.py:85-98), NIfTI-1 I/O, slicing <-> scatter inverses
(eval.py:176-190), resampling on integer grids.  No GPU, no compute calls into the CUDA library."""
import gzip
import os
import struct

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import pmu_oracle as O

import pmu_b200  # noqa: E402  (tests/conftest.py puts the repo root on sys.path)
from pmu_b200 import nifti_io
from pmu_b200.multiplanar import padded_dims, plane_affine, reduce_scatter_accumulators, shard_slices

dims_st = st.tuples(st.integers(1, 40), st.integers(1, 40), st.integers(1, 40))
planes_st = st.sampled_from([(0, 1, 2), (0,), (1,), (2,), (0, 2), (2, 1), (1, 0, 2)])


@settings(max_examples=200, deadline=None)
@given(dims=dims_st, planes=planes_st, world=st.integers(1, 9))
def test_shard_slices_is_an_ordered_balanced_partition(dims, planes, world):
    total = sum(dims[p] for p in planes)
    chunk = -(-total // world)
    flat = [(p, s) for p in planes for s in range(dims[p])]          # the reference's index_map order for these planes
    seen, sizes = [], []
    for r in range(world):
        part = shard_slices(dims, planes, r, world)
        assert list(part) == [p for p in planes if p in part]         # planes stay in order inside a rank
        mine = [(p, s) for p, (a, b) in part.items() for s in range(a, b)]
        assert all(0 <= a < b <= dims[p] for p, (a, b) in part.items())
        sizes.append(len(mine))
        seen += mine
    assert seen == flat                                               # no gaps, no duplicates, same order
    assert max(sizes) <= chunk and sum(sizes) == total
    assert all(s == chunk for s in sizes[: total // chunk])           # every rank before the last non-empty one is full


def test_shard_slices_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_slices((4, 4, 4), (0, 1, 2), 2, 2)
    with pytest.raises(ValueError):
        shard_slices((4, 4, 4), (0, 1, 2), -1, 2)


@settings(max_examples=200, deadline=None)
@given(dims=dims_st)
def test_padded_dims_matches_the_oracle_pad(dims):
    """pad_dimensions pads ONLY the arg-min axis, at its high end, up to the max extent (mri_dataset.py:85-98)."""
    v = np.arange(np.prod(dims), dtype=np.float64).reshape(dims) + 1.0
    padded = O.pad_dimensions(v)
    assert padded_dims(dims) == padded.shape
    assert np.array_equal(padded[: dims[0], : dims[1], : dims[2]], v)
    assert padded.sum() == v.sum()                                    # the padding is zeros
    assert sorted(padded.shape)[-1] == max(dims)


@pytest.mark.parametrize("plane", [0, 1, 2])
def test_plane_affine_is_the_oracle_identity_grid(plane):
    assert np.array_equal(np.array(plane_affine(plane), np.float32), O.identity_affine(plane))


@settings(max_examples=40, deadline=None)
@given(dims=st.tuples(st.integers(1, 9), st.integers(1, 9), st.integers(1, 9)), plane=st.integers(0, 2),
       C=st.integers(1, 4), seed=st.integers(0, 2 ** 16))
def test_scatter_is_the_inverse_of_slicing(dims, plane, C, seed):
    """slices_to_volume + the permutes of eval.py:176-190 put slice s of `plane` back at index s of that axis, for every
    channel: scatter_plane(plane, slices(vol_c)) == vol_c."""
    rng = np.random.default_rng(seed)
    vols = rng.standard_normal((C,) + dims).astype(np.float32)
    per_c = [O.plane_slices(vols[c], plane, normalise=False)[:, 0] for c in range(C)]       # [S, H, W] each
    stacked = torch.from_numpy(np.stack(per_c, 1))                                          # [S, C, H, W]
    back = O.scatter_plane(plane, stacked).numpy()                                           # [x, C, y, z]
    assert back.shape == (dims[0], C, dims[1], dims[2])
    for c in range(C):
        assert np.array_equal(back[:, c], vols[c])


@settings(max_examples=30, deadline=None)
@given(dims=st.tuples(st.integers(2, 8), st.integers(2, 8), st.integers(2, 8)), plane=st.integers(0, 2),
       shift=st.tuples(st.integers(-3, 3), st.integers(-3, 3), st.integers(-3, 3)), mode=st.sampled_from(["nearest", "trilinear"]),
       seed=st.integers(0, 2 ** 16))
def test_resampling_on_an_integer_shifted_grid_is_a_shifted_copy_with_zero_fill(dims, plane, shift, mode, seed):
    rng = np.random.default_rng(seed)
    vol = rng.random(dims).astype(np.float32)
    aff = O.identity_affine(plane).copy()
    aff[:3] = np.array(shift, np.float32)
    H, W = [dims[a] for a in range(3) if a != plane]
    out = O.resample_slices(vol, aff, 0, dims[plane], H, W, mode)
    big = np.zeros(tuple(d + 6 for d in dims), np.float32)
    big[3:3 + dims[0], 3:3 + dims[1], 3:3 + dims[2]] = vol
    sh = big[3 + shift[0]: 3 + shift[0] + dims[0], 3 + shift[1]: 3 + shift[1] + dims[1], 3 + shift[2]: 3 + shift[2] + dims[2]]
    want = O.plane_slices(sh, plane, normalise=False)[:, 0]
    assert np.array_equal(out, want)


def _raw_nifti(data: np.ndarray, code: int, endian: str = "<", slope: float = 1.0, inter: float = 0.0, vox_offset: float = 352.0,
               extra_dims=()):
    hdr = bytearray(348)
    struct.pack_into(endian + "i", hdr, 0, 348)
    shape = list(data.shape) + list(extra_dims)
    struct.pack_into(endian + "8h", hdr, 40, len(shape), *(shape + [1] * (7 - len(shape))))
    struct.pack_into(endian + "h", hdr, 70, code)
    struct.pack_into(endian + "h", hdr, 72, data.dtype.itemsize * 8)
    struct.pack_into(endian + "f", hdr, 108, vox_offset)
    struct.pack_into(endian + "2f", hdr, 112, slope, inter)
    hdr[344:348] = b"n+1\x00"
    body = np.asfortranarray(data.astype(data.dtype.newbyteorder(endian))).tobytes(order="F")
    return bytes(hdr) + b"\x00" * (int(vox_offset) - 348) + body


@pytest.mark.parametrize("dtype,code", [(np.uint8, 2), (np.int16, 4), (np.int32, 8), (np.float32, 16), (np.float64, 64),
                                        (np.int8, 256), (np.uint16, 512)])
@pytest.mark.parametrize("endian", ["<", ">"])
def test_nifti_reader_datatypes_endianness_and_scaling(tmp_path, dtype, code, endian):
    rng = np.random.default_rng(3)
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    data = (rng.integers(max(info.min, -1000), min(info.max, 1000), (3, 4, 5)).astype(dtype) if info
            else rng.standard_normal((3, 4, 5)).astype(dtype))
    p = tmp_path / "v.nii"
    p.write_bytes(_raw_nifti(data, code, endian))
    got = nifti_io.load(str(p))
    assert got.dtype == np.float64 and got.flags["C_CONTIGUOUS"]
    assert np.array_equal(got, data.astype(np.float64))
    # scl_slope / scl_inter (nibabel get_fdata applies them), a larger vox_offset, a trailing singleton 4th dimension
    p.write_bytes(_raw_nifti(data, code, endian, slope=0.5, inter=-2.0, vox_offset=400.0, extra_dims=(1,)))
    got = nifti_io.load(str(p))
    assert got.shape == (3, 4, 5)
    np.testing.assert_allclose(got, data.astype(np.float64) * 0.5 - 2.0, rtol=0, atol=1e-12)
    # slope 0 means "no scaling" in NIfTI-1
    p.write_bytes(_raw_nifti(data, code, endian, slope=0.0, inter=7.0))
    assert np.array_equal(nifti_io.load(str(p)), data.astype(np.float64))


def test_nifti_writer_header_and_gzip(tmp_path):
    v = np.random.default_rng(1).random((4, 5, 6))
    aff = np.diag([2.0, 3.0, 4.0, 1.0])
    p = tmp_path / "w.nii.gz"
    nifti_io.save(str(p), v, affine=aff)
    raw = gzip.open(p, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == 348 and raw[344:348] == b"n+1\x00"
    assert struct.unpack("<8h", raw[40:56])[:4] == (3, 4, 5, 6)
    assert struct.unpack("<h", raw[70:72])[0] == 16 and struct.unpack("<f", raw[108:112])[0] == 352.0
    assert struct.unpack("<4f", raw[280:296]) == (2.0, 0.0, 0.0, 0.0)
    assert len(raw) == 352 + 4 * v.size
    # voxel order on disk is Fortran (x fastest), as NIfTI prescribes
    assert np.array_equal(np.frombuffer(raw, np.float32, offset=352).reshape(v.shape, order="F"), v.astype(np.float32))
    with pytest.raises(ValueError):
        (tmp_path / "bad.nii").write_bytes(_raw_nifti(np.zeros((2, 2, 2), np.float32), 1234))
        nifti_io.load(str(tmp_path / "bad.nii"))


def test_reduce_scatter_accumulators_single_rank_and_bad_split():
    acc = torch.arange(2 * 6 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 6, 2, 3, 3)
    out, (x0, x1) = reduce_scatter_accumulators(acc, 0, 1)
    assert out is acc and (x0, x1) == (0, 6)
    with pytest.raises(ValueError):
        reduce_scatter_accumulators(acc, 0, 4)          # X = 6 is not divisible by 4
