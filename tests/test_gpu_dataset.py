"""GPU parity tests of the rows SURVEY.md §8(f) ranks 2 and 4: the HBM-resident MRI_Dataset (index map with the
background filter as a GPU reduction, slice / mask items), arbitrary view vectors through the affine gather, and the
latent-grid sweep of visualize_sampling.py — against the real reference's data-plane fixture
(tests/golden/dataplane_ref.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import pmu_oracle as O

pytestmark = pytest.mark.gpu

CASES = ["cube", "pad0", "pad1", "pad2"]


@pytest.fixture(scope="module")
def pmu():
    import pmu_b200
    return pmu_b200


@pytest.fixture(scope="module")
def trainer_sd():
    return O.make_state_dict(seed=0)


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "dataplane_ref.npz"))


@pytest.mark.parametrize("case", CASES)
def test_dataset_index_map_and_items_match_reference(pmu, ref, case):
    """index_map (filter=True / False), image_dims and every item of the real MRI_Dataset (run unmodified with stub
    nibabel, make_golden_dataplane.py).  Masks bit for bit; images bit for bit against the reference arithmetic on
    the fp32-held volume (oracle, pinned to the same fixture) and to one fp32 ulp against the fp64-held fixture."""
    img, mask = ref[f"{case}_img"], ref[f"{case}_mask"]
    vols = {"scan0.nii": (img, mask)}
    ds_all = pmu.MRI_Dataset(None, None, 3, filter=False, volumes=vols)
    ds_f = pmu.MRI_Dataset(None, None, 3, filter=True, volumes=vols)
    assert np.array_equal(np.array(ds_all.index_map), ref[f"{case}_all_index_map"])
    assert np.array_equal(np.array(ds_f.index_map), ref[f"{case}_filt_index_map"])
    assert tuple(ds_all.image_dims) == tuple(int(v) for v in ref[f"{case}_all_image_dims"])
    assert len(ds_all) == len(ref[f"{case}_all_index_map"]) and len(ds_f) == len(ref[f"{case}_filt_index_map"])
    assert ds_all.fp32_exact == [False]                      # random fp64 data: rounded once on upload
    padded32 = O.pad_dimensions(img).astype(np.float32)
    assert np.array_equal(ds_all.volume(0).cpu().numpy(), padded32)
    i = 0
    for view in range(3):
        want_img, want_mask = ref[f"{case}_all_view{view}_image"], ref[f"{case}_all_view{view}_mask"]
        exact = O.plane_slices(padded32.astype(np.float64), view)
        for s in range(want_img.shape[0]):
            item = ds_all[i]
            assert item["image"].shape == (1,) + want_img.shape[2:] and item["image"].dtype == torch.float32
            assert np.array_equal(item["mask"].cpu().numpy(), want_mask[s])
            got = item["image"].cpu().numpy()
            assert np.array_equal(got, exact[s])
            np.testing.assert_allclose(got, want_img[s], rtol=4e-7, atol=0)
            i += 1
    # batch assembly == stacking the items (DataLoader + default_collate), runs of consecutive slices and singletons
    idx = [0, 1, 2, len(ds_all) - 1, 3, 5]
    b = ds_all.gather_batch(idx)
    for j, k in enumerate(idx):
        it = ds_all[k]
        assert torch.equal(b["image"][j], it["image"]) and torch.equal(b["mask"][j], it["mask"])
    from torch.utils.data import DataLoader
    first = next(iter(DataLoader(ds_all, batch_size=4, shuffle=False, num_workers=0)))
    assert first["image"].is_cuda and torch.equal(first["image"], ds_all.gather_batch(range(4))["image"])
    got = list(ds_f.batches(4, shuffle=True, seed=3))
    assert sum(x["image"].shape[0] for x in got) == len(ds_f)


def test_dataset_integer_volume_is_bit_exact_with_reference_arithmetic(pmu):
    """int16-valued data (the usual NIfTI case): the fp32 cache is lossless, items equal the reference's
    fp64 divide + .float() exactly; two scans; the filter drops background-only slices of each view."""
    g = np.random.default_rng(5)
    vols = {}
    for k in range(2):
        img = g.integers(-200, 3000, size=(12, 12, 12)).astype(np.float64)
        img[3] = 0.0                                     # all-zero slice: no divide
        img[:, 5, :] = -np.abs(img[:, 5, :]) - 1.0       # negative maximum
        mask = np.zeros((12, 12, 12))
        mask[4:8, 2:9, 3:7] = g.integers(0, 3, size=(4, 7, 4))
        vols[f"s{k}.nii"] = (img, mask)
    ds = pmu.MRI_Dataset(None, None, 3, filter=True, volumes=vols)
    assert ds.fp32_exact == [True, True]
    want_map = []
    for scan, key in enumerate(sorted(vols)):
        m = vols[key][1]
        for view in range(3):
            for s in range(12):
                if O.sample_slice(m, view, s).max() > 0:
                    want_map.append((scan, view, s))
    assert ds.index_map == want_map and 0 < len(ds) < 72
    for i in range(0, len(ds), 5):
        scan, view, s = ds.index_map[i]
        img, m = vols[sorted(vols)[scan]]
        item = ds[i]
        assert np.array_equal(item["image"].cpu().numpy(), O.normalised_slice(img, view, s))
        assert np.array_equal(item["mask"].cpu().numpy()[0], O.sample_slice(m, view, s).astype(np.float32))


def test_dataset_from_nifti_files_and_predictor(pmu, tmp_path, trainer_sd):
    """imgs_dir / masks_dir on disk (built-in NIfTI reader), then the cached volume feeds MultiPlanarPredictor
    without another host round trip."""
    from pmu_b200 import nifti_io
    vol, lab = O.phantom(16, seed=11)
    os.makedirs(tmp_path / "images"); os.makedirs(tmp_path / "labels")
    nifti_io.save(str(tmp_path / "images" / "a.nii"), vol)
    nifti_io.save(str(tmp_path / "labels" / "a.nii"), lab.astype(np.float32))
    ds = pmu.MRI_Dataset(str(tmp_path / "images"), str(tmp_path / "labels"), 3, filter=True)
    assert ds.ids == ["a.nii"] and ds.image_dims == (16, 16, 16) and ds.fp32_exact == [True]
    keep = [(0, v, s) for v in range(3) for s in range(16) if O.sample_slice(lab, v, s).max() > 0]
    assert ds.index_map == keep
    pred = pmu.MultiPlanarPredictor(trainer_sd, "cuda", precision="f16", n_samples=2, slice_batch=16)
    eps = torch.randn(3, 16, 2, 6, generator=torch.Generator().manual_seed(2))
    a = pred.predict(ds.volume(0), eps=eps)
    b = pred.predict(vol, eps=eps)
    assert torch.equal(a["mean"], b["mean"])


def test_cpu_device_is_refused(pmu):
    with pytest.raises(RuntimeError):
        pmu.MRI_Dataset(None, None, 3, device="cpu", volumes={"a": (np.zeros((4, 4, 4)), np.zeros((4, 4, 4)))})


@pytest.mark.parametrize("view", [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 2, 3), (-1, 0.5, 0.25)])
def test_view_vectors_through_the_affine_gather(pmu, view):
    """Standard view vectors reproduce sample_slice exactly; oblique ones are the oracle's trilinear / nearest
    resampling on the same grid, bit for bit, and the grid's frame is orthonormal and centred."""
    vol, lab = O.phantom(0, seed=9, dims=(40, 40, 40))
    ds = pmu.MRI_Dataset(None, None, 3, filter=False, volumes={"a": (vol, lab)})
    aff, (H, W), n = pmu.view_affine(view, vol.shape)
    A = np.array(aff, np.float32)
    std = [tuple(v) for v in np.eye(3, dtype=int)]
    if tuple(view) in std:
        p = std.index(tuple(view))
        assert np.array_equal(A, O.identity_affine(p))
        got = ds.oblique_slices(0, view)
        assert np.array_equal(got.cpu().numpy(), O.plane_slices(vol, p))
        return
    nn, u, v = A[3:6].astype(np.float64), A[6:9].astype(np.float64), A[9:12].astype(np.float64)
    np.testing.assert_allclose([nn @ nn, u @ u, v @ v], 1.0, atol=1e-6)
    np.testing.assert_allclose([nn @ u, nn @ v, u @ v], 0.0, atol=1e-6)
    mid = A[:3] + (n - 1) / 2.0 * (nn + u + v)
    np.testing.assert_allclose(mid, (np.array(vol.shape) - 1) / 2.0, atol=1e-4)
    for mode in ("trilinear", "nearest"):
        want = O.resample_slices(vol, A, 3, 20, H, W, mode)
        got = ds.oblique_slices(0, view, s0=3, ns=20, interp=mode, normalise=False)
        assert np.array_equal(got.cpu().numpy()[:, 0], want), mode
    got = ds.oblique_slices(0, view, s0=3, ns=20)
    assert np.array_equal(got.cpu().numpy()[:, 0], O.normalise_slices(O.resample_slices(vol, A, 3, 20, H, W, "trilinear")))
    m = ds.oblique_slices(0, view, s0=3, ns=20, mask=True)
    assert np.array_equal(m.cpu().numpy()[:, 0], O.resample_slices(lab.astype(np.float32), A, 3, 20, H, W, "nearest"))


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_latent_grid_sweep(pmu, trainer_sd, precision):
    """visualize_sampling.py:21-31: the G x G grid around the prior mean equals G*G separate sample_at(z) calls
    (the reference's loop), and the oracle's fcomb at the same z."""
    tr = pmu.ProbUNetTrainer(torch.device("cuda"), 1, 3, latent_dim=6, beta=10, precision=precision)
    tr.net.load_state_dict(trainer_sd, strict=True)
    tr.net.eval()
    x = torch.rand(1, 1, 32, 32, generator=torch.Generator().manual_seed(3)).cuda()
    mask = torch.zeros(1, 1, 32, 32).cuda()
    imgs, logits, z = tr.latent_grid(x, mask, n_preds=3, sigma_scale=40.0)
    assert logits.shape == (1, 3, 3, 3, 32, 32) and z.shape == (1, 3, 3, 6) and imgs.shape == (1, 3, 3, 3, 32, 32)
    mu = tr.net.prior_latent_space.base_dist.loc[0]
    sigma = tr.net.prior_latent_space.base_dist.scale[0] * 40.0
    feat_cpu = tr.net.unet_features.cpu()
    for i, z0 in enumerate(range(-1, 2)):
        for j, z1 in enumerate(range(-1, 2)):
            zz = torch.stack([1 * z0 * sigma[0] + mu[0], 1 * z1 * sigma[1] + mu[1], mu[2], mu[3], mu[4], mu[5]])
            assert torch.allclose(z[0, i, j], zz, rtol=0, atol=1e-6)
            with torch.no_grad():
                one = tr.predict(x, mask, z=zz)                   # the reference's per-cell call
            assert torch.allclose(one[0], logits[0, i, j], rtol=1e-5, atol=1e-5)
            want = O.fcomb(trainer_sd, feat_cpu, z[0, i, j].cpu()[None])
            assert float((want[0] - logits[0, i, j].cpu()).abs().max()) < 2e-4 * max(1.0, float(want.abs().max()))
    # colour table == the reference's per-pixel loop
    lab = torch.argmax(logits[0, 1, 1], 0).cpu()
    colors = torch.tensor([[0., 0., 0.], [0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])
    assert torch.equal(imgs[0, 1, 1].cpu(), colors[lab].permute(2, 0, 1))
    tm = torch.randint(0, 3, (2, 1, 8, 8)).float().cuda()
    assert torch.equal(tr.mask_to_image(tm).cpu(), colors[tm.cpu().squeeze(1).long()].permute(0, 3, 1, 2))
