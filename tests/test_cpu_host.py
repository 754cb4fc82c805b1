"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, host
logic (sharding, padding), oracle data-plane identities, and the N>1 reduce path over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import pmu_b200
from oracle import pmu_oracle as O
from pmu_b200 import _lib
from pmu_b200.multiplanar import padded_dims, reduce_accumulators, reduce_scatter_accumulators, shard_slices


def test_library_loads_and_exports_header_symbols():
    if not os.path.exists(_lib.LIB_PATH):
        from pmu_b200.build import build_native
        build_native()
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib._SIG) == set(syms)          # the ctypes table covers the whole header
    assert lib.pmu_version() == 100


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA, never compute on the CPU."""
    x = torch.zeros(4, 4, 4)
    with pytest.raises(RuntimeError):
        pmu_b200.ops.plane_max(x)
    net = pmu_b200.ProbabilisticUnet(1, 3, [4, 8], 2, 2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net.forward(torch.zeros(1, 1, 8, 8), None, training=False)
    with pytest.raises(RuntimeError):
        pmu_b200.MultiPlanarPredictor(net, device="cpu")


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "probabilistic-multiplanar-unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    # the entry points and the helper scripts stay oracle-free too: only tests/ (incl. tests/tools/), smoke() and
    # bench.py's cpu_baseline / --impl reference legs may touch it
    others = [os.path.join(root, f) for f in ("eval.py", "predict.py", "pmu_b200/__init__.py")]
    others += [os.path.join(root, "scripts", f) for f in os.listdir(os.path.join(root, "scripts"))]
    for f in others:
        if os.path.isfile(f) and f.endswith((".py", ".sh", ".cu")):
            src = open(f).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_schema_matches_reference():
    sd = O.make_state_dict((8, 16, 32), latent_dim=4, no_convs_fcomb=3)
    net = pmu_b200.ProbabilisticUnet(1, 3, [8, 16, 32], latent_dim=4, no_convs_fcomb=3)
    net.load_state_dict(sd, strict=True)
    with pytest.raises(ValueError):
        pmu_b200.ProbabilisticUnet(1, 3, [32, 64, 128, 192])   # reference default crashes too (App. B #9)


def test_padded_dims_matches_oracle():
    for d in [(24, 40, 40), (40, 24, 40), (40, 40, 24), (16, 16, 16), (8, 12, 16)]:
        assert padded_dims(d) == O.pad_dimensions(np.zeros(d)).shape


def test_shard_slices_partition():
    for dims in [(256, 256, 256), (24, 40, 40), (5, 7, 3)]:
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                for p, (a, b) in shard_slices(dims, (0, 1, 2), r, world).items():
                    seen += [(p, s) for s in range(a, b)]
            assert seen == O.index_map(dims)     # same order as the reference's index_map, no gaps/dups


def test_oracle_slicing_identities():
    vol, _ = O.phantom(0, dims=(6, 7, 8))
    for p in range(3):
        a = O.plane_slices(vol, p, normalise=False)[:, 0]
        b = O.resample_slices(vol, O.identity_affine(p), 0, vol.shape[p], a.shape[1], a.shape[2], "nearest")
        c = O.resample_slices(vol, O.identity_affine(p), 0, vol.shape[p], a.shape[1], a.shape[2], "trilinear")
        assert np.array_equal(a, b) and np.array_equal(a, c)
        # scatter is the inverse of slicing (eval.py:176-190)
        back = O.scatter_plane(p, torch.from_numpy(a)[:, None])[:, 0].numpy()
        assert np.array_equal(back, vol)


def test_oracle_trilinear_matches_grid_sample():
    vol, _ = O.phantom(0, dims=(9, 10, 11))
    aff = np.array([0.3, -0.2, 0.4, 0.9, 0.1, 0.0, -0.1, 0.95, 0.05, 0.02, 0.0, 1.05], np.float32)
    out = O.resample_slices(vol, aff, 0, 8, 10, 11, "trilinear")
    a = aff.astype(np.float64).reshape(4, 3)
    sg, rg, cg = np.meshgrid(np.arange(8.0), np.arange(10.0), np.arange(11.0), indexing="ij")
    qx, qy, qz = [a[0, ax] + sg * a[1, ax] + rg * a[2, ax] + cg * a[3, ax] for ax in range(3)]   # the affine grid, exactly
    D0, D1, D2 = vol.shape
    grid = torch.from_numpy(np.stack([2 * qz / (D2 - 1) - 1, 2 * qy / (D1 - 1) - 1, 2 * qx / (D0 - 1) - 1], -1))[None].float()
    ref = torch.nn.functional.grid_sample(torch.from_numpy(vol)[None, None], grid, mode="bilinear",
                                          padding_mode="zeros", align_corners=True)[0, 0].numpy()
    np.testing.assert_allclose(out, ref, atol=1e-5)


def test_oracle_fusion_reduces_to_reference_average():
    g = torch.Generator().manual_seed(3)
    v = [torch.softmax(torch.randn(5, 3, 5, 5, generator=g), 1) for _ in range(3)]
    mean, var, ent = O.finalize(v[0] + v[1] + v[2], v[0] ** 2 + v[1] ** 2 + v[2] ** 2, 3)
    torch.testing.assert_close(mean, (v[0] + v[1] + v[2]) / 3.0)   # eval.py:193
    assert (var >= 0).all() and (ent >= 0).all() and (ent <= np.log(3) + 1e-6).all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, dims, ret):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    per_slice = {p: torch.rand(dims[p], 2, 2, *[dims[a] for a in range(3) if a != p], generator=g) for p in range(3)}
    acc = torch.zeros(2, dims[0], 2, dims[1], dims[2])
    for p, (a, b) in shard_slices(dims, (0, 1, 2), rank, world).items():
        for j in range(2):
            full = torch.zeros_like(per_slice[p][:, j])
            full[a:b] = per_slice[p][a:b, j]
            acc[j] += O.scatter_plane(p, full)
    ref = torch.zeros_like(acc)
    for p in range(3):
        for j in range(2):
            ref[j] += O.scatter_plane(p, per_slice[p][:, j])
    # slab-sharded exchange (reduce-scatter along x; gloo takes the all-reduce fallback): every rank owns its x-slab
    err_slab = 0.0
    if dims[0] % world == 0:
        mine, (x0, x1) = reduce_scatter_accumulators(acc.clone(), rank, world, None)
        err_slab = float((mine - ref[:, x0:x1]).abs().max())
        assert (x1 - x0) * world == dims[0]
    reduce_accumulators(acc, world, None, dst=0)
    errs = [torch.tensor([err_slab])]
    gathered = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(gathered, errs[0])
    if rank == 0:
        ret.put(max(float((acc - ref).abs().max()), max(float(g) for g in gathered)))
    dist.destroy_process_group()


def test_sharded_accumulators_reduce_over_gloo():
    """world_size-2 run of the N>1 host path: each rank scatters its slice chunk, ONE reduce
    joins them, result equals the single-rank accumulators."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, (6, 5, 7), ret)) for r in range(2)]
    for p in procs:
        p.start()
    err = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6


def _gloo_grad_worker(rank, world, port, ret):
    """Data-parallel training exchange (train_dp.py) on 2 gloo ranks: every rank differentiates its
    half batch of the ORACLE objective with the KL term weighted beta / world, gradients are
    SUM-all-reduced in buckets; rank 0 compares with the single-process gradient of
    sum_b CE + beta * mean_b KL evaluated with the same per-shard BatchNorm statistics."""
    import torch.distributed as dist
    from pmu_b200.train_dp import allreduce_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    beta = 10.0
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 1, 16, 16, generator=g)
    segm = torch.randint(0, 3, (4, 1, 16, 16), generator=g).float()
    eps = torch.randn(4, 6, generator=g)

    def params(seed=0):
        sd = O.make_state_dict((4, 8), num_classes=3, latent_dim=6, no_convs_fcomb=3, seed=seed)
        ps = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items() if v.dtype == torch.float32 and "running_" not in k
              and "outc" not in k}
        full = dict(sd); full.update(ps)
        return ps, full

    ps, sd = params()
    sl = slice(rank * 2, rank * 2 + 2)
    r = O.elbo(sd, x[sl], segm[sl], eps[sl], beta=beta / world, bn_train=True)
    (-r["elbo"]).backward()
    n_coll = allreduce_gradients(ps.values(), bucket_bytes=4096)      # small buckets: several collectives
    if rank == 0:
        ps1, sd1 = params()
        total = 0
        for rr in range(world):
            s2 = slice(rr * 2, rr * 2 + 2)
            o = O.elbo(sd1, x[s2], segm[s2], eps[s2], beta=beta, bn_train=True)
            total = total + o["reconstruction_loss"] + beta * o["kl"] / world
        total.backward()
        err = max(float((ps[k].grad - ps1[k].grad).abs().max()) / max(float(ps1[k].grad.abs().max()), 1e-3) for k in ps)
        ret.put((err, n_coll))
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_over_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_grad_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    err, n_coll = ret.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-4, err
    assert n_coll > 1


def _gloo_tape_worker(rank, world, port, ret):
    """The range-announcing gradient tape of the tensor-core training step (train_engine._Tape with a _FlatLayout) on 2 gloo
    ranks: gradients are put in completion order, every announced range is all-reduced at once (what GraphedTrainStep does
    inside its CUDA graph); the result must equal an all-reduce of every gradient after the fact, the ranges must tile the
    buffer, and an out-of-order put must fall back to one final range without losing anything."""
    import torch.distributed as dist
    from pmu_b200 import train_engine as te
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11)
    params = [torch.nn.Parameter(torch.zeros(*shp)) for shp in ((64, 33), (7,), (128, 64, 3, 3), (5, 5), (256, 130), (3,))]
    grads = [[torch.randn(p.shape, generator=g) for p in params] for _ in range(world)]       # every rank draws all, uses its own
    layout = te._FlatLayout(params)
    out = {}
    for name, order in (("in_order", list(range(len(params)))), ("shuffled", [0, 2, 1, 3, 5, 4])):
        ranges = []

        def exchange(buf, lo, hi):
            ranges.append((lo, hi))
            dist.all_reduce(buf[lo:hi])

        tape = te._Tape(layout, "cpu", exchange, bucket_bytes=64 * 1024)
        for i in order:
            if i % 2 == 0:                                    # written in place through out(p) ...
                tape.out(params[i]).copy_(grads[rank][i])
                tape.put(params[i])
            else:                                             # ... or handed over and copied in
                tape.put(params[i], grads[rank][i].clone())
        tape.finish()
        want = [sum(grads[r][i] for r in range(world)) for i in range(len(params))]
        err = max(float((tape.g[id(p)] - w).abs().max()) for p, w in zip(params, want))
        tiles = ranges[0][0] == 0 and ranges[-1][1] == layout.total and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        out[name] = (err, len(ranges), tiles, tape.in_order)
    if rank == 0:
        ret.put(out)
    dist.destroy_process_group()


def test_flat_gradient_ranges_allreduce_over_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_tape_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    err, n, tiles, in_order = out["in_order"]
    assert err < 1e-6 and n >= 3 and tiles and in_order, out
    err, n, tiles, in_order = out["shuffled"]
    assert err < 1e-6 and tiles and not in_order, out


def test_nifti_roundtrip(tmp_path):
    from pmu_b200 import nifti_io
    v = np.random.default_rng(0).random((5, 6, 7))
    for name in ("a.nii", "b.nii.gz"):
        p = str(tmp_path / name)
        nifti_io.save(p, v)
        w = nifti_io.load(p)
        assert w.shape == v.shape and w.dtype == np.float64
        np.testing.assert_array_equal(w, v.astype(np.float32).astype(np.float64))
    with pytest.raises(ValueError):
        (tmp_path / "bad.nii").write_bytes(b"x" * 400)
        nifti_io.load(str(tmp_path / "bad.nii"))


def test_view_affine_frames():
    """Host logic of the generalised views: standard axes give the oracle's identity grids; any other vector gives an
    orthonormal right-handed frame centred on the volume, with unit spacing."""
    import pmu_b200
    dims = (40, 48, 56)
    for p in range(3):
        aff, hw, n = pmu_b200.view_affine(np.eye(3, dtype=int)[p], dims)
        assert np.array_equal(np.array(aff, np.float32), O.identity_affine(p))
        assert n == dims[p] and hw == tuple(d for a, d in enumerate(dims) if a != p)
    for view in [(1, 1, 0), (1, 2, 3), (-1, 0.5, 0.25), (0, 0, 2)]:
        aff, hw, n = pmu_b200.view_affine(view, dims)
        A = np.array(aff, np.float64)
        nn, u, v = A[3:6], A[6:9], A[9:12]
        np.testing.assert_allclose([nn @ nn, u @ u, v @ v], 1.0, atol=1e-6)
        np.testing.assert_allclose([nn @ u, nn @ v, u @ v], 0.0, atol=1e-6)
        np.testing.assert_allclose(np.cross(nn, u), v, atol=1e-6)
        np.testing.assert_allclose(nn, np.array(view, float) / np.linalg.norm(view), atol=1e-6)
        assert hw == (56, 56) and n == 56
        np.testing.assert_allclose(A[:3] + (n - 1) / 2.0 * (nn + u + v), (np.array(dims) - 1) / 2.0, atol=1e-4)
    with pytest.raises(ValueError):
        pmu_b200.view_affine((0, 0, 0), dims)


def test_dataset_and_pipeline_refuse_cpu():
    import pmu_b200
    with pytest.raises(RuntimeError):
        pmu_b200.MRI_Dataset(None, None, 3, device="cpu", volumes={"a": (np.zeros((4, 4, 4)), np.zeros((4, 4, 4)))})


@pytest.mark.parametrize("name", ["golden_cfg2_lattice.npz", "golden_cfg3_lattice.npz"])
def test_fullsize_lattice_fixtures_are_consistent(golden_dir, name):
    """The whole-volume oracle runs behind the full-size GPU tests: probabilities sum to one, population variance and
    natural-log entropy are inside their ranges, the checksums agree with the lattice's scale."""
    g = np.load(os.path.join(golden_dir, name))
    n = len(range(int(g["offset"]), int(g["D"]), int(g["step"])))
    assert g["mean"].shape == (n, n, n, 3) and g["var"].shape == (n, n, n, 3) and g["entropy"].shape == (n, n, n)
    np.testing.assert_allclose(g["mean"].sum(-1), 1.0, atol=1e-5)
    assert g["var"].min() >= 0 and g["var"].max() <= 0.25
    assert g["entropy"].min() >= 0 and g["entropy"].max() <= np.log(3) + 1e-6
    p = np.clip(g["mean"], 1e-12, 1)
    np.testing.assert_allclose(-(p * np.log(p)).sum(-1), g["entropy"], atol=1e-5)
    np.testing.assert_allclose(g["mean_sum"].sum(), float(g["D"]) ** 3, rtol=1e-6)
    assert int(g["labels_hist"].sum()) == int(g["D"]) ** 3


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/pmu_b200.h against the ctypes table of _lib.py: same parameter count and the same
    class of type per position (pointer / int / int64_t / float) — a drifted binding corrupts arguments silently."""
    import ctypes
    import re
    from pmu_b200 import _lib
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = re.findall(r"\b(?:const\s+char\s*\*|int)\s+(pmu_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert sorted(n for n, _ in protos) == _lib.header_symbols()
    assert sorted(_lib._SIG) == _lib.header_symbols()            # the table binds exactly what the header declares

    def kind_of_c(param):
        param = param.strip()
        if "*" in param or "[" in param:
            return "ptr"
        t = param.rsplit(" ", 1)[0].replace("const", "").strip()
        return {"int": "int", "int32_t": "int", "int64_t": "int64", "float": "float"}[t]

    def kind_of_ctypes(t):
        if t is ctypes.c_void_p or t is ctypes.c_char_p or hasattr(t, "_type_") and not isinstance(t._type_, str):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_int32: "int", ctypes.c_int64: "int64", ctypes.c_float: "float"}[t]

    for name, params in protos:
        plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        _, args = _lib._SIG[name]
        assert len(plist) == len(args), (name, len(plist), len(args))
        for i, (p, a) in enumerate(zip(plist, args)):
            assert kind_of_c(p) == kind_of_ctypes(a), (name, i, p.strip(), a)


def test_every_compute_entry_point_cites_the_reference():
    """include/pmu_b200.h: the comment in front of every compute prototype names the reference file (file:line) whose
    arithmetic it replaces; housekeeping entry points (error string, version, device, launch context, fill) have no counterpart."""
    import re
    from pmu_b200 import _lib
    src = open(_lib.HEADER_PATH).read()
    tokens = re.findall(r"/\*.*?\*/|\b(?:const\s+char\s*\*|int)\s+pmu_[a-z0-9_]+\s*\([^)]*\)\s*;", src, flags=re.S)
    last, missing = "", []
    for t in tokens:
        if t.startswith("/*"):
            last = t
        elif not re.search(r"[a-z_]+\.py:\d+", last):
            missing.append(re.search(r"(pmu_[a-z0-9_]+)", t).group(1))
    assert set(missing) <= {"pmu_last_error", "pmu_version", "pmu_device_info", "pmu_set_device", "pmu_fill_f32",
                            "pmu_ctx_create", "pmu_ctx_destroy", "pmu_ctx_bind", "pmu_ctx_stats"}, missing


def test_python_constants_match_the_header():
    """Workspace sizes the Python wrappers allocate follow include/pmu_b200.h (PMU_RED_MAX_BLOCKS rows per reduction)."""
    import re
    from pmu_b200 import _lib, ops
    src = open(_lib.HEADER_PATH).read()
    assert int(re.search(r"#define\s+PMU_RED_MAX_BLOCKS\s+(\d+)", src).group(1)) == ops.RED_MAX_BLOCKS
    assert int(re.search(r"#define\s+PMU_POOL_MAX\s+(\d+)", src).group(1)) == ops.POOL_MAX
    assert int(re.search(r"#define\s+PMU_POOL_AVG_CEIL\s+(\d+)", src).group(1)) == ops.POOL_AVG_CEIL


def test_missing_library_raises_loudly(monkeypatch, tmp_path):
    """The product never builds or falls back on its own: with no libpmu_b200.so, loading — and therefore every op —
    raises a RuntimeError that says how to build it."""
    from pmu_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libpmu_b200.so"))
    with pytest.raises(RuntimeError, match="not built|not found"):
        _lib.load()
    with pytest.raises(RuntimeError):
        _lib.check(-1, "anything")          # even error reporting needs the library
