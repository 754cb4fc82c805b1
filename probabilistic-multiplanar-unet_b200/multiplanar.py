"""Multi-planar probabilistic inference: volume -> voxel-space mean / variance / entropy.

The hot path of the reference's eval.py main loop (eval.py:132-214) re-designed for one B200
(or N of them): the volume stays resident in HBM, slices are gathered in batches per plane
(K1), U-Net + prior run ONCE per slice (the reference re-runs the whole network per sample,
eval.py:148-152), N latent samples go through the fused fcomb/softmax/accumulate kernel (K3+K4)
and the per-slice sums are scatter-added into [x,C,y,z] accumulators (K4).  Multi-GPU: the flat
(plane, slice) list — exactly the reference's index_map order, utils/mri_dataset.py:37-49 — is
cut into contiguous chunks, one per rank; ONE collective (sum-reduce of the accumulators) joins
them.

Semantics follow SURVEY.md Appendix A (eval-mode BN, probabilities averaged over samples and
planes, population variance, natural-log entropy).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .engine import PackedNet


def padded_dims(dims: Sequence[int]) -> Tuple[int, int, int]:
    """MRI_Dataset.pad_dimensions (mri_dataset.py:85-98): only the arg-min axis is padded, up to
    the max extent."""
    d = list(int(v) for v in dims)
    diff = max(d) - min(d)
    if diff:
        d[int(np.argmin(d))] += diff
    return tuple(d)


def shard_slices(dims: Sequence[int], planes: Sequence[int], rank: int, world: int) -> Dict[int, Tuple[int, int]]:
    """Contiguous chunk of the plane-major flat slice list for `rank` -> {plane: (s0, s1)}."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    total = sum(int(dims[p]) for p in planes)
    chunk = -(-total // world)
    g0, g1 = min(total, rank * chunk), min(total, (rank + 1) * chunk)
    out: Dict[int, Tuple[int, int]] = {}
    base = 0
    for p in planes:
        lo, hi = max(g0, base), min(g1, base + int(dims[p]))
        if hi > lo:
            out[p] = (lo - base, hi - base)
        base += int(dims[p])
    return out


def plane_affine(plane: int):
    """12 floats [o, n, u, v] of the standard view `plane`: q = o + s*n + r*u + c*v reproduces
    MRI_Dataset.sample_slice (mri_dataset.py:72-77) — the identity resampling grid."""
    e = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]
    n, u, v = {0: (e[0], e[1], e[2]), 1: (e[1], e[0], e[2]), 2: (e[2], e[0], e[1])}[plane]
    return [0.0, 0.0, 0.0] + n + u + v


def reduce_accumulators(acc: torch.Tensor, world: int, group=None, dst: Optional[int] = 0) -> torch.Tensor:
    """The single exchange step: sum the [2,X,C,Y,Z] accumulators over ranks (NCCL on GPUs over
    NVLink/NVSwitch; gloo in the CPU tests).  dst=None -> all-reduce."""
    if world <= 1:
        return acc
    import torch.distributed as dist
    if dst is None:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return acc


def reduce_scatter_accumulators(acc: torch.Tensor, rank: int, world: int, group=None):
    """Slab-sharded form of the exchange step: every rank ends up with the SUM of its own x-slab
    [2, X/world, C, Y, Z] of the accumulators — `ncclReduceScatter` along x, one call per moment (Σp, Σp² are two
    contiguous [X, ...] tensors), so that each rank finalises and delivers only its slab.  X must be divisible by
    world.  Backends without reduce-scatter (gloo in the CPU tests) fall back to all-reduce + slicing."""
    X = acc.shape[1]
    if X % world:
        raise ValueError(f"slab-sharded output needs X = {X} divisible by world = {world}")
    xs = X // world
    if world <= 1:
        return acc, (0, X)
    import torch.distributed as dist
    out = torch.empty((2, xs) + tuple(acc.shape[2:]), dtype=acc.dtype, device=acc.device)
    if dist.get_backend(group) == "nccl":
        for j in range(2):
            dist.reduce_scatter_tensor(out[j], acc[j], op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        out.copy_(acc[:, rank * xs:(rank + 1) * xs])
    return out, (rank * xs, (rank + 1) * xs)


class MultiPlanarPredictor:
    """3-plane, N-sample probabilistic prediction of a volume on the current CUDA device."""

    def __init__(self, state_dict, device="cuda", precision: str = "f16", n_samples: int = 16,
                 planes: Sequence[int] = (0, 1, 2), slice_batch: int = 32, interp: str = "exact",
                 affines: Optional[Dict[int, Sequence[float]]] = None, out_hw: Optional[Tuple[int, int]] = None,
                 rank: int = 0, world_size: int = 1, process_group=None, output: str = "rank0",
                 upload: str = "each", graph: Optional[bool] = None):
        if hasattr(state_dict, "state_dict"):
            state_dict = state_dict.state_dict()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MultiPlanarPredictor runs on CUDA only (no CPU fallback)")
        self.net = PackedNet(state_dict, self.device, precision)
        self.n_samples, self.planes, self.slice_batch = int(n_samples), tuple(planes), int(slice_batch)
        if interp not in ops.INTERP:
            raise ValueError(f"interp must be one of {list(ops.INTERP)}")
        if affines is None:
            affines = {p: plane_affine(p) for p in self.planes}     # resample onto the standard grids
        self.identity_grid = all(list(map(float, affines[p])) == plane_affine(p) for p in self.planes)
        self.interp, self.affines, self.out_hw = interp, affines, out_hw
        self.rank, self.world, self.group = int(rank), int(world_size), process_group
        if output not in ("rank0", "slab"):
            raise ValueError("output must be 'rank0' (one reduce, results on rank 0) or 'slab' (reduce-scatter along x, "
                             "every rank keeps its x-slab)")
        self.output = output
        if upload not in ("each", "broadcast"):
            raise ValueError("upload must be 'each' (every rank copies the volume from its own host buffer) or 'broadcast' "
                             "(submit(): rank 0 copies it once, the other ranks receive it over NVLink)")
        self.upload = upload
        # graph: submit() on several GPUs replays the whole slice pass of a volume as ONE CUDA graph (accumulate_graphed)
        # instead of ~45 launches per slice batch — the host thread of every rank goes idle.  Default (None): on when
        # world_size > 1, where it is the faster e2e path (8 GPUs, 256^3 x 16 samples: 16.3 -> 15.5 ms per volume,
        # profiles/r02a_n8_bench*.json.log); one GPU keeps the streaming path (results leave x-slab by x-slab).
        self.graph = (int(world_size) > 1) if graph is None else bool(graph)
        self.C = self.net.fcomb["C"]
        self.L = self.net.fcomb["L"]

    # ------------------------------------------------------------------
    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def _to_device_volume(self, vol) -> torch.Tensor:
        if isinstance(vol, np.ndarray):
            vol = torch.from_numpy(np.ascontiguousarray(vol, dtype=np.float32))
        vol = vol.to(self.device, torch.float32, non_blocking=True)
        pd = padded_dims(vol.shape)
        if tuple(vol.shape) != pd:
            big = torch.zeros(pd, dtype=torch.float32, device=self.device)
            big[: vol.shape[0], : vol.shape[1], : vol.shape[2]] = vol
            vol = big
        return vol.contiguous()

    def accumulate(self, vol: torch.Tensor, eps: torch.Tensor, acc: torch.Tensor, only_plane: Optional[int] = None,
                   plane0_last: bool = False, on_slab=None, cnt: Optional[torch.Tensor] = None) -> int:
        """Run this rank's slices and add their sums into acc [2,X,C,Y,Z]; returns #slices done.
        only_plane restricts the pass to one view (per-view volumes of eval.py:176-190).
        plane0_last processes the x-slicing view after the others, so that after every plane-0 batch the voxels
        of that x-slab are complete; on_slab(x0, x1) is then called (streaming finalise + device->host copy)."""
        dims = tuple(vol.shape)
        net, N = self.net, self.n_samples
        my = shard_slices(dims, self.planes, self.rank, self.world)
        # on the standard grids (exact slicing, or nearest / trilinear resampling onto them) a slice's max is
        # the volume's per-plane max: one pass over the volume, normalisation fused into the gather
        exact = self.interp == "exact" or self.identity_grid
        maxes = ops.plane_max(vol) if exact else None
        offs = (0, dims[0], dims[0] + dims[1])
        done = 0
        order = list(enumerate(self.planes))
        if plane0_last:
            order = [(pi, p) for pi, p in order if p != 0] + [(pi, p) for pi, p in order if p == 0]
        for pi, p in order:
            if p not in my or (only_plane is not None and p != only_plane):
                continue
            s_lo, s_hi = my[p]
            # K1 once per plane: all of this rank's slices of the plane in one launch (full-bandwidth
            # granularity: 134 MB of traffic per 256^3 plane instead of 33 MB per batch)
            if exact:
                xs = ops.slice_gather(vol, p, s_lo, s_hi - s_lo, interp=self.interp,
                                      affine=None if self.interp == "exact" else self.affines[p],
                                      slice_max_in=maxes[offs[p]: offs[p] + dims[p]])
            else:
                xs, mx = ops.slice_gather(vol, p, s_lo, s_hi - s_lo, interp=self.interp, affine=self.affines[p],
                                          hw=self.out_hw, want_max=True)
                ops.slice_normalize_(xs, mx)
            for s0 in range(s_lo, s_hi, self.slice_batch):
                ns = min(self.slice_batch, s_hi - s0)
                x = xs[s0 - s_lo: s0 - s_lo + ns]
                feat = net.unet_features(x)
                mu, ls = net.gaussian("prior", x)
                sigma = torch.exp(ls)          # Normal(scale=exp(log_sigma)), probabilistic_unet.py:113
                sums = net.fcomb_sums(feat, mu, sigma, eps[pi, s0:s0 + ns].contiguous())
                if exact:
                    ops.scatter_accum_(sums, p, s0, dims, acc[0], acc[1])
                else:
                    # non-identity grid (App. A step 6): nearest-voxel scatter with a per-voxel count of the samples
                    ops.scatter_accum_affine_(sums, self.affines[p], s0, dims, acc[0], acc[1], cnt, float(N))
                done += ns
                if on_slab is not None and p == 0:
                    on_slab(s0, s0 + ns)
        return done

    @torch.no_grad()
    def accumulate_graphed(self, vol: torch.Tensor, eps: torch.Tensor, acc: torch.Tensor) -> int:
        """acc = 0; accumulate(vol, eps, acc) — captured into a CUDA graph on first use for this (vol, eps, acc) buffer
        triple and replayed afterwards: one launch per volume instead of ~2100 (256^3 on one GPU), no host work between
        kernels.  The buffers are baked into the graph: refill them in place (vol.copy_(...)) between calls.  The
        graph's intermediates (gathered slices, activations of one slice batch) live in a memory pool shared by all
        graphs of this predictor, which is safe because they are replayed on one stream, one after the other.
        Falls back to the eager path while ops.PROFILE is recording per-launch events."""
        if ops.PROFILE is not None:
            acc.zero_()
            return self.accumulate(vol, eps, acc)
        key = (vol.data_ptr(), eps.data_ptr(), acc.data_ptr(), tuple(vol.shape), tuple(eps.shape))
        graphs = self.__dict__.setdefault("_graphs", {})
        ent = graphs.get(key)
        if ent is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):          # eager warm-up off the capture: module loads, allocator growth
                acc.zero_()
                self.accumulate(vol, eps, acc)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            pool = next(iter(graphs.values()))[0].pool() if graphs else None
            l0 = ops.LAUNCHES
            # thread_local: the NCCL watchdog thread of a multi-GPU job may query events while we capture
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                acc.zero_()
                n = self.accumulate(vol, eps, acc)
            ent = graphs[key] = (g, n, ops.LAUNCHES - l0)
        g, n, launches = ent
        g.replay()
        ops.LAUNCHES += launches                   # bench.py's gpu_launches counts kernels, not graph launches
        return n

    # ------------------------------------------------------------------ pipelined serving path
    @torch.no_grad()
    def submit(self, vol_host: torch.Tensor, eps: torch.Tensor, host_out: Dict[str, torch.Tensor],
               want_labels: bool = False, depth: int = 2) -> int:
        """Asynchronous predict for a stream of volumes: enqueue one volume and return a ticket at once.

        vol_host: PINNED fp32 host tensor [d0,d1,d2] (already cubic / padded) — with upload="broadcast" and
        world_size > 1 only rank 0's buffer is read (one PCIe upload, then one NCCL broadcast over NVLink; the other
        ranks pass a tensor of the same shape, or a torch.Size / tuple of dims); eps: device tensor [P, D, N, L];
        host_out: PINNED host tensors {"mean", "var", "entropy"[, "labels"]} — the whole volume on one GPU, this rank's
        x-slab with output="slab", rank 0 only with output="rank0".  `wait(ticket)` (or `wait()` for everything
        submitted) blocks the host until the results are in host_out.

        Three streams: the host->device copy of volume k+1 runs on a copy stream while volume k computes on the current
        stream, and the device->host copies of volume k's results run on a second copy stream behind the next volume's
        kernels.  `depth` slots of device buffers (volume, accumulators, outputs) rotate; a slot is rewritten only after
        the compute that read it / the copy-out that drained it has finished (CUDA events, no host synchronisation).
        Same kernels, same results as predict(host_out=...) — only the scheduling differs."""
        if self.interp != "exact" and not self.identity_grid:
            raise NotImplementedError("voxel fusion is defined for the standard axis-aligned grids")
        bcast = self.upload == "broadcast" and self.world > 1
        reads_host = not bcast or self.rank == 0
        if reads_host and not (isinstance(vol_host, torch.Tensor) and vol_host.dtype == torch.float32 and vol_host.is_pinned()):
            raise ValueError("submit() takes a pinned fp32 host tensor (torch.Tensor.pin_memory())")
        dims = tuple(vol_host.shape) if isinstance(vol_host, torch.Tensor) else tuple(int(d) for d in vol_host)
        if padded_dims(dims) != dims:
            raise ValueError("submit() takes an already padded volume (pad_dimensions); use predict() otherwise")
        P, N = len(self.planes), self.n_samples
        slab = self.output == "slab" and self.world > 1
        if slab and dims[0] % self.world:
            raise ValueError(f"slab-sharded output needs X = {dims[0]} divisible by world = {self.world}")
        pipe = getattr(self, "_pipe", None)
        if pipe is None or pipe["dims"] != dims or pipe["depth"] != depth or pipe["labels"] != want_labels:
            self.wait()
            dev = self.device
            xs = dims[0] // self.world if slab else dims[0]
            mine = self.rank == 0 or slab
            slots = []
            for _ in range(depth):
                s = {"vol": torch.empty(dims, dtype=torch.float32, device=dev),
                     "acc": torch.empty(2, dims[0], self.C, dims[1], dims[2], dtype=torch.float32, device=dev),
                     "in": torch.cuda.Event(), "compute": None, "out": None}
                if mine:
                    s["mean"] = torch.empty(xs, self.C, dims[1], dims[2], dtype=torch.float32, device=dev)
                    s["var"] = torch.empty_like(s["mean"])
                    s["entropy"] = torch.empty(xs, dims[1], dims[2], dtype=torch.float32, device=dev)
                    s["labels"] = torch.empty(xs, dims[1], dims[2], dtype=torch.uint8, device=dev) if want_labels else None
                slots.append(s)
            pipe = self._pipe = {"dims": dims, "depth": depth, "labels": want_labels, "slots": slots, "k": 0,
                                 "h2d": torch.cuda.Stream(dev), "pending": {}}
        k = pipe["k"]
        pipe["k"] = k + 1
        s = pipe["slots"][k % depth]
        main, h2d, d2h = torch.cuda.current_stream(self.device), pipe["h2d"], self._side_stream()
        # ---- copy-in on its own stream (waits for the compute that last read this slot's volume) ----
        if s["compute"] is not None:
            h2d.wait_event(s["compute"])
        with torch.cuda.stream(h2d):
            if reads_host:
                s["vol"].copy_(vol_host, non_blocking=True)
            if bcast:
                # issued from the copy stream: NCCL orders the broadcast behind the upload, and the compute stream only
                # waits for the slot's "in" event (same collective order on every rank: broadcast k, exchange k, ...)
                import torch.distributed as dist
                dist.broadcast(s["vol"], group=self.group, group_src=0)
            s["in"].record(h2d)
        main.wait_event(s["in"])
        if s["out"] is not None:
            main.wait_event(s["out"])                     # the slot's outputs have left for the host
        acc = s["acc"]
        graphed = self.graph and not (self.world == 1 and 0 in self.planes)
        if not graphed:
            acc.zero_()
        keys = [k_ for k_ in ("mean", "var", "entropy", "labels") if k_ in host_out and s.get(k_) is not None]

        def copy_out(x0, x1):
            d2h.wait_stream(main)
            with torch.cuda.stream(d2h):
                for k_ in keys:
                    host_out[k_][x0:x1].copy_(s[k_][x0:x1], non_blocking=True)

        if self.world == 1 and 0 in self.planes:
            def on_slab(x0, x1):
                ops.fuse_finalize(acc[0][x0:x1], acc[1][x0:x1], float(P * N), want_labels=want_labels,
                                  out=(s["mean"][x0:x1], s["var"][x0:x1], s["entropy"][x0:x1],
                                       None if s["labels"] is None else s["labels"][x0:x1]))
                copy_out(x0, x1)
            self.accumulate(s["vol"], eps, acc, plane0_last=True, on_slab=on_slab)
        else:
            if graphed:
                self.accumulate_graphed(s["vol"], eps, acc)
            else:
                self.accumulate(s["vol"], eps, acc)
            if slab:
                part, _ = reduce_scatter_accumulators(acc, self.rank, self.world, self.group)
            else:
                reduce_accumulators(acc, self.world, self.group, dst=0)
                part = acc
            if self.rank == 0 or slab:
                ops.fuse_finalize(part[0], part[1], float(P * N), want_labels=want_labels,
                                  out=(s["mean"], s["var"], s["entropy"], s["labels"]))
                copy_out(0, s["mean"].shape[0])
        s["compute"] = torch.cuda.Event()
        s["compute"].record(main)
        s["out"] = torch.cuda.Event()
        s["out"].record(d2h)
        pipe["pending"][k] = s["out"]
        return k

    @torch.no_grad()
    def _predict_oblique(self, vol, eps, seed, want_labels, keep_sums):
        """predict() on NON-identity slice grids (arbitrary view vectors — the reference's use_standard_axis=False TODO,
        utils/mri_dataset.py:60-71): slices are resampled (nearest / trilinear) on the
        affine grids, every pixel's sums go to its nearest voxel and a per-voxel count replaces the constant P * N of
        the standard views: mean = S1 / cnt, var = S2 / cnt - mean^2.  Voxels no slice pixel lands on read 0
        ("count" in the result is the per-voxel tensor).  Multi-GPU: the counts are reduced with the accumulators."""
        vol = self._to_device_volume(vol)
        dims = tuple(vol.shape)
        P, N, L = len(self.planes), self.n_samples, self.L
        if eps is None:
            g = torch.Generator(device=self.device).manual_seed(seed)
            eps = torch.randn(P, max(dims), N, L, generator=g, device=self.device)
        else:
            eps = eps.to(self.device, torch.float32, non_blocking=True)
        acc = torch.zeros(2, dims[0], self.C, dims[1], dims[2], dtype=torch.float32, device=self.device)
        cnt = torch.zeros(dims, dtype=torch.float32, device=self.device)
        self.accumulate(vol, eps, acc, cnt=cnt)
        reduce_accumulators(acc, self.world, self.group, dst=0)
        reduce_accumulators(cnt, self.world, self.group, dst=0)
        out: Dict[str, torch.Tensor] = {"count": cnt}
        if self.rank == 0:
            mean, var, ent, lab = ops.fuse_finalize_counted(acc[0], acc[1], cnt, want_labels=want_labels)
            out.update(mean=mean, var=var, entropy=ent)
            if want_labels:
                out["labels"] = lab
        if keep_sums:
            out["S1"], out["S2"] = acc[0], acc[1]
        return out

    def wait(self, ticket: Optional[int] = None) -> None:
        """Block the host until `ticket`'s results (default: everything submitted) are complete in host memory."""
        pipe = getattr(self, "_pipe", None)
        if pipe is None:
            return
        for k in sorted(pipe["pending"]):
            if ticket is None or k <= ticket:
                pipe["pending"].pop(k).synchronize()

    @torch.no_grad()
    def predict(self, vol, eps: Optional[torch.Tensor] = None, seed: int = 4321, want_labels: bool = False,
                keep_sums: bool = False, per_plane: bool = False, host_out: Optional[Dict[str, torch.Tensor]] = None
                ) -> Dict[str, torch.Tensor]:
        """vol: [d0,d1,d2] fp32 (numpy / CPU / CUDA).  eps: [P, D, N, L] standard-normal draws
        (host or device); generated on the device from `seed` when omitted.  Returns mean / var
        [x,C,y,z], entropy [x,y,z] (valid on rank 0 when world_size > 1; with output="slab" every rank gets its
        own x-slab, "x_range" = (x0, x1), from a reduce-scatter instead of a reduce).  per_plane=True also
        returns "plane_means": the per-view probability volumes volume1/2/3 of eval.py:176-190.
        host_out = {"mean", "var", "entropy"[, "labels"]} of PINNED host tensors: the results are also delivered to
        the host; on one GPU the x-slicing view runs last, every finished x-slab is finalised and copied out on a
        side stream while the next slice batch computes (the 470 MB device->host copy of a 256^3 result leaves
        the critical path), and the call returns with the copies complete on the current stream."""
        if self.interp != "exact" and not self.identity_grid:
            return self._predict_oblique(vol, eps, seed, want_labels, keep_sums)
        vol = self._to_device_volume(vol)
        dims = tuple(vol.shape)
        P, N, L = len(self.planes), self.n_samples, self.L
        Dmax = max(dims)
        if eps is None:
            g = torch.Generator(device=self.device).manual_seed(seed)
            eps = torch.randn(P, Dmax, N, L, generator=g, device=self.device)
        else:
            eps = eps.to(self.device, torch.float32, non_blocking=True)
        acc = torch.zeros(2, dims[0], self.C, dims[1], dims[2], dtype=torch.float32, device=self.device)
        out: Dict[str, torch.Tensor] = {"count": float(P * N)}
        if per_plane:
            plane_means = []
            for p in self.planes:
                pacc = torch.zeros_like(acc)
                self.accumulate(vol, eps, pacc, only_plane=p)
                reduce_accumulators(pacc, self.world, self.group, dst=0)
                if self.rank == 0:
                    plane_means.append(ops.fuse_finalize(pacc[0], pacc[1], float(N), want_var=False, want_entropy=False)[0])
                acc += pacc
            out["plane_means"] = plane_means
        elif host_out is not None and self.world == 1 and 0 in self.planes:
            # streaming path: finalise + copy out x-slabs as the last view completes them
            mean = torch.empty_like(acc[0])
            var = torch.empty_like(acc[0])
            ent = torch.empty(dims, dtype=torch.float32, device=self.device)
            lab = torch.empty(dims, dtype=torch.uint8, device=self.device) if want_labels else None
            main, side = torch.cuda.current_stream(self.device), self._side_stream()
            keys = [("mean", mean), ("var", var), ("entropy", ent)] + ([("labels", lab)] if want_labels else [])

            def on_slab(x0, x1):
                ops.fuse_finalize(acc[0][x0:x1], acc[1][x0:x1], float(P * N), want_labels=want_labels,
                                  out=(mean[x0:x1], var[x0:x1], ent[x0:x1], None if lab is None else lab[x0:x1]))
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for k, t in keys:
                        if k in host_out:
                            host_out[k][x0:x1].copy_(t[x0:x1], non_blocking=True)

            self.accumulate(vol, eps, acc, plane0_last=True, on_slab=on_slab)
            main.wait_stream(side)
            out.update(mean=mean, var=var, entropy=ent)
            if want_labels:
                out["labels"] = lab
            if keep_sums:
                out["S1"], out["S2"] = acc[0], acc[1]
            return out
        elif self.output == "slab" and self.world > 1:
            # slab-sharded outputs: reduce-scatter along x, every rank finalises (and copies out) its own x-slab
            self.accumulate(vol, eps, acc)
            mine, (x0, x1) = reduce_scatter_accumulators(acc, self.rank, self.world, self.group)
            mean, var, ent, lab = ops.fuse_finalize(mine[0], mine[1], float(P * N), want_labels=want_labels)
            out.update(mean=mean, var=var, entropy=ent, x_range=(x0, x1))
            if want_labels:
                out["labels"] = lab
            if host_out is not None:
                for k in ("mean", "var", "entropy", "labels"):
                    if k in host_out and k in out:
                        host_out[k][: x1 - x0].copy_(out[k], non_blocking=True)
            if keep_sums:
                out["S1"], out["S2"] = mine[0], mine[1]
            return out
        else:
            self.accumulate(vol, eps, acc)
            reduce_accumulators(acc, self.world, self.group, dst=0)
        if self.rank == 0:
            mean, var, ent, lab = ops.fuse_finalize(acc[0], acc[1], float(P * N), want_labels=want_labels)
            out.update(mean=mean, var=var, entropy=ent)
            if want_labels:
                out["labels"] = lab
            if host_out is not None:
                for k in ("mean", "var", "entropy", "labels"):
                    if k in host_out and k in out:
                        host_out[k].copy_(out[k], non_blocking=True)
        if keep_sums:
            out["S1"], out["S2"] = acc[0], acc[1]
        return out
