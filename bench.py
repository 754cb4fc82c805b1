#!/usr/bin/env python
"""bench.py — volumes/sec of the multi-planar probabilistic inference hot path.

    python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): a synthetic
256^3 fp32 volume, 3 planes x 16 z-samples, axis-aligned slicing, trainer-architecture
ProbabilisticUnet([64..1024], C=3, L=6, fcomb 4 convs) with random-init weights, mean /
variance / entropy fusion.  One "step" = one whole volume.

  value : volumes/s, volume + latents resident in HBM, outputs left in HBM (CUDA events, max
          over ranks).  N > 1: the slice list of ONE volume is sharded over the ranks and the
          accumulators are reduce-scattered along x, every rank finalising its x-slab
          ("scaling": "strong"; --resident-output rank0: sum-reduced to rank 0).
  e2e   : same metric through the public API with the volume in pinned host memory and mean / var /
          entropy read back to pinned host memory, every step: MultiPlanarPredictor.submit()/wait()
          (copies of neighbouring volumes overlap compute); sync_ms_per_step is the one-at-a-time
          MultiPlanarPredictor.predict(host_out=...) figure.
  roofline : the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs of all its
          launches in one step / their summed CUDA-event durations, against the measured
          sustained bf16 peak of MEASURED_PEAKS.json.  hbm_kernels adds the same for the
          gather and scatter-accumulate kernels (GB/s vs measured HBM copy bandwidth).
  cpu_baseline : the oracle port of the reference path on the host cores, on a bounded sample
          of the same workload (N = 1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "volumes/sec (256^3, 3 planes x 16 samples)"
UNIT = "volumes/s"


def workload(D, N, interp):
    """config.workload — the SAME string on both arms (the reference arm's sampling is described in its cpu_baseline)."""
    return (f"{D}^3 fp32 volume, 3 planes x {N} z-samples, {interp} resampling on the standard plane grids, trainer model "
            f"[64,128,256,512,1024] C=3 L=6 fcomb=4, mean/var/entropy fusion")


def config_dict(D, N, interp):
    """config — the SAME dict on both arms (what differs between the arms lives in `run`, `cpu_baseline`, `e2e`)."""
    return {"workload": workload(D, N, interp),
            "l2": "per-step working set (GBs of activations; the CPU arm: every slice's activations) >> 126 MB L2; no explicit flush"}


def load_traffic():
    """DRAM bytes per launch of our kernels from the last committed `ncu --set full` capture
    (profiles/*_traffic.json, written by scripts/summarize_ncu.py)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return {}
    try:
        return json.load(open(files[-1]))
    except Exception:
        return {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.gpu = [], threading.Event(), gpu_index
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------
def cpu_reference_rate(D, N, slices_per_plane, threads=None):
    """Oracle port of the reference CPU path (variant B of BASELINE.md: forward once per slice +
    N x fcomb + softmax + accumulate, batch 1 like eval.py:105) on a stratified sample of
    `slices_per_plane` slices per plane; returns (volumes/s extrapolated, seconds, cores)."""
    import numpy as np
    import torch
    from oracle import pmu_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    sd = O.make_state_dict(seed=0)
    vol, _ = O.phantom(D, seed=1234)
    eps = torch.randn(3, D, N, 6, generator=torch.Generator().manual_seed(4321))
    # equally spaced slices, centred in their strata (never only index 0, which is background in the phantom)
    idx = [min(D - 1, int((i + 0.5) * D / slices_per_plane)) for i in range(slices_per_plane)]
    C = 3
    with torch.no_grad():
        x = torch.from_numpy(O.plane_slices(vol, 0, idx[0], 1))
        O.unet_features(sd, x)                       # warm-up (thread pools, oneDNN primitives)
        t0 = time.perf_counter()
        for p in range(3):
            H, W = [vol.shape[a] for a in range(3) if a != p]
            s1 = torch.zeros(len(idx), C, H, W)
            s2 = torch.zeros_like(s1)
            for j, s in enumerate(idx):
                x = torch.from_numpy(O.plane_slices(vol, p, s, 1))
                feat = O.unet_features(sd, x)
                mu, ls = O.gaussian_head(sd, "prior", x)
                sigma = torch.exp(ls)
                for n in range(N):
                    pr = torch.softmax(O.fcomb(sd, feat, mu + sigma * eps[p, s:s + 1, n]), 1)
                    s1[j] += pr[0]
                    s2[j] += pr[0] * pr[0]
            O.scatter_plane(p, s1).contiguous()
        dt = time.perf_counter() - t0
    n_done = 3 * len(idx)
    return (n_done / dt) / (3.0 * D), dt, cores, n_done


def cpu_variant_a_rate(D, N, threads=None):
    """For information (SURVEY.md §8d "variant A"): the literal sample loop of eval.py:148-152 — N FULL predict() calls
    (U-Net + prior + fcomb) per slice — on the middle slice of each plane; returns (volumes/s extrapolated, seconds)."""
    import torch
    from oracle import pmu_oracle as O
    if threads:
        torch.set_num_threads(threads)
    sd = O.make_state_dict(seed=0)
    vol, _ = O.phantom(D, seed=1234)
    eps = torch.randn(3, D, N, 6, generator=torch.Generator().manual_seed(4321))
    with torch.no_grad():
        t0 = time.perf_counter()
        for p in range(3):
            s = D // 2
            x = torch.from_numpy(O.plane_slices(vol, p, s, 1))
            acc = 0
            for n in range(N):
                feat = O.unet_features(sd, x)
                mu, ls = O.gaussian_head(sd, "prior", x)
                acc = acc + torch.softmax(O.fcomb(sd, feat, mu + torch.exp(ls) * eps[p, s:s + 1, n]), 1)
        dt = time.perf_counter() - t0
    return (3 / dt) / (3.0 * D), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, N = args.size, args.samples
    spp = args.ref_slices_per_plane     # 8 equally spaced slices per plane per step, ~5 s per step on 16 threads
    rates = []
    for i in range(args.warmup + args.steps):
        # all host threads, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
        v, dt, cores, n_done = cpu_reference_rate(D, N, spp, threads=os.cpu_count())
        if i >= args.warmup:
            rates.append((v, dt))
    tot_t = sum(dt for _, dt in rates)
    value = (len(rates) * 3 * spp / tot_t) / (3.0 * D)
    va, va_dt = cpu_variant_a_rate(D, N, threads=os.cpu_count())
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / max(len(rates), 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config_dict(D, N, args.interp),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = {3 * spp} of {3 * D} slices ({spp} equally spaced per plane) through the oracle "
                                       f"port of the reference CPU path (fp32, batch 1 like eval.py:105): forward once + {N} x fcomb "
                                       f"+ softmax + accumulate (variant B), extrapolated to {3 * D} slices; on the standard plane "
                                       f"grids {args.interp} resampling is exact slicing, which the port does",
                             "variant_a": {"value": va, "unit": UNIT, "seconds": va_dt,
                                           "what": f"for information: literal eval.py:148-152, {N} full predict() calls per slice, "
                                                   f"middle slice of each plane, extrapolated"}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    if not os.path.exists(os.path.join(ROOT, "probabilistic-multiplanar-unet_b200", "libpmu_b200.so")):
        # fresh checkout: the CUDA library is a git-ignored build product.  Rank 0 of the node compiles it, the others wait
        # for the file (no CPU path exists to fall back to: without the library the import below raises).
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__
            __graft_entry__.build()
        else:
            pkg, t_end = os.path.join(ROOT, "probabilistic-multiplanar-unet_b200"), time.time() + 900
            while time.time() < t_end and not (os.path.exists(os.path.join(pkg, "libpmu_b200.so"))
                                               and os.path.exists(os.path.join(pkg, "build", "stamp.txt"))):
                time.sleep(2)
            time.sleep(1)             # the stamp is written after the link step
    import pmu_b200
    from pmu_b200 import ops
    from pmu_b200.synthetic import phantom_volume, trainer_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # short collective timeout: a rank-divergence bug must cost minutes of GPU time, not ten
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    D, N, P = args.size, args.samples, 3
    peaks = load_peaks()

    sd = trainer_state_dict(seed=0)
    pred = pmu_b200.MultiPlanarPredictor(sd, dev, precision=args.precision, n_samples=N, slice_batch=args.slice_batch,
                                         interp=args.interp, rank=rank, world_size=world)
    vol_host = phantom_volume(D, seed=1234).pin_memory()
    vol = vol_host.to(dev)
    eps = torch.randn(P, D, N, 6, generator=torch.Generator(device=dev).manual_seed(4321), device=dev)
    acc = torch.zeros(2, D, 3, D, D, dtype=torch.float32, device=dev)

    # N > 1: the one exchange step is a reduce-scatter of the accumulators along x — every rank finalises its own x-slab
    # and the outputs stay slab-sharded in HBM (the path predict(output="slab") takes).
    # --resident-output rank0 sum-reduces everything to rank 0 instead (what rounds 1a-1d measured).
    slab_res = world > 1 and D % world == 0 and args.resident_output == "slab"

    # N > 1: the slice pass of a rank is replayed as ONE CUDA graph (one launch per volume instead of ~270 per rank; 8 GPUs:
    # 14.68 -> 14.53 ms resident, 16.3 -> 15.5 ms e2e); one GPU keeps eager launches (no gain measured: 120.4 vs 119.7 ms)
    use_graph = args.graph or (world > 1 and not args.no_graph)

    def step_resident():
        if use_graph:
            pred.accumulate_graphed(vol, eps, acc)          # zero + slice pass, one CUDA-graph launch
        else:
            acc.zero_()
            pred.accumulate(vol, eps, acc)
        if slab_res:
            part, _ = pmu_b200.reduce_scatter_accumulators(acc, rank, world, None)
            return ops.fuse_finalize(part[0], part[1], float(P * N))
        pmu_b200.reduce_accumulators(acc, world, None, dst=0)
        if rank == 0:
            return ops.fuse_finalize(acc[0], acc[1], float(P * N))
        return None

    # e2e leg: the public call with host buffers.  N = 1: results stream to rank 0's pinned buffers x-slab by x-slab
    # behind the last view.  N > 1: slab-sharded outputs (reduce-scatter along x) — every rank
    # finalises its x-slab and copies it to its own pinned buffers, so the 470 MB of results leave over N PCIe links.
    slab = world > 1 and D % world == 0
    pred_e2e = pred if not slab else pmu_b200.MultiPlanarPredictor(
        sd, dev, precision=args.precision, n_samples=N, slice_batch=args.slice_batch, interp=args.interp, rank=rank,
        world_size=world, output="slab", upload=args.e2e_upload, graph=True if args.graph else (False if args.no_graph else None))

    def pinned_outputs():
        if not (rank == 0 or slab):
            return {}
        Dx = D // world if slab else D
        return {"mean": torch.empty(Dx, 3, D, D, dtype=torch.float32).pin_memory(),
                "var": torch.empty(Dx, 3, D, D, dtype=torch.float32).pin_memory(),
                "entropy": torch.empty(Dx, D, D, dtype=torch.float32).pin_memory()}

    host_outs = [pinned_outputs(), pinned_outputs()]

    def step_e2e_sync():
        pred_e2e.predict(vol_host, eps=eps, host_out=host_outs[0] or None)

    # serving form of the same call: submit() returns at once, the upload of volume k+1 and the read-back of volume k
    # overlap the kernels of their neighbours (three streams, two buffer slots); wait() blocks until every result is in
    # host memory, inside the timed region
    def run_e2e(steps):
        for i in range(steps):
            pred_e2e.submit(vol_host, eps, host_outs[i % 2])
        pred_e2e.wait()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ops.LAUNCHES
    ms = timed(step_resident, args.steps)
    launches = ops.LAUNCHES - l0
    clocks = sampler.finish() if sampler else None

    if args.timed_only:
        # profiling aid (ncu launch list / captures of the timed step's kernels): skip the e2e legs and the per-kernel
        # instrumentation, print the resident number only
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                              "gpu_launches": launches, "note": "--timed-only (profiling aid): warm-up + timed resident steps only"}))
        if world > 1:
            dist.destroy_process_group()
        return
    step_e2e_sync()                                             # warm the e2e paths (pinned buffers, allocator)
    ms_e2e_sync = timed(step_e2e_sync, args.steps)
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1)

    phases = None
    if args.e2e_phases:
        # where the e2e leg's time goes when all ranks hit the host at once: each phase alone, max over ranks
        dvol = torch.empty_like(vol)
        Dx = D // world if slab else D
        dres = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host_outs[0].items()}

        def ph_h2d():
            dvol.copy_(vol_host, non_blocking=True)

        def ph_d2h():
            for k, v in dres.items():
                host_outs[0][k].copy_(v, non_blocking=True)

        def ph_slab():
            acc.zero_()
            pred_e2e.accumulate(vol, eps, acc)
            if slab:
                part, _ = pmu_b200.reduce_scatter_accumulators(acc, rank, world, None)
            else:
                part = pmu_b200.reduce_accumulators(acc, world, None, dst=0)
            if slab or rank == 0:
                ops.fuse_finalize(part[0], part[1], float(P * N))

        phases = {}
        for name, fn in (("h2d_volume", ph_h2d), ("d2h_results", ph_d2h), ("resident_step_with_exchange", ph_slab)):
            fn()
            phases[name] = timed(fn, args.steps) / args.steps
        del dvol, dres

    # ---- per-kernel roofline: one instrumented step, CUDA events around every C-ABI call ----
    roof, hbm_kernels, shares = None, {}, {}
    traffic = load_traffic()
    # EVERY rank runs the instrumented step (it contains the reduce collective); rank 0 reports
    barrier()
    ops.PROFILE = []
    step_resident()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
        tot = {}
        for name, meta, a, b in prof:
            t = a.elapsed_time(b)
            if name == "pmu_conv_gemm_pool_bf16":
                name = "pmu_conv_gemm_bf16"          # same kernel (conv_tc_kernel), pooled epilogue
            d = tot.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "tmem": 0.0})
            d["ms"] += t; d["n"] += 1
            if meta:
                d["flops"] += meta.get("flops", 0.0)
                d["bytes"] += meta.get("bytes", 0.0)
                d["tmem"] += meta.get("tmem_read_bytes", 0.0)
        step_ms = sum(d["ms"] for d in tot.values())
        shares = {k: round(d["ms"] / step_ms, 4) for k, d in sorted(tot.items(), key=lambda kv: -kv[1]["ms"])}
        dom = max(tot.items(), key=lambda kv: kv[1]["ms"])
        if dom[0] == "pmu_conv_gemm_bf16":
            c = dom[1]
            ach = c["flops"] / (c["ms"] * 1e-3) / 1e12
            tk = [traffic[k] for k in ("conv_tc_kernel", "conv_rs_kernel") if k in traffic]
            tr = (sum(t["dram_bytes"] for t in tk) / max(1, sum(t["launches_captured"] for t in tk))) if tk else None
            roof = {"kernel": "conv_tc_kernel + conv_rs_kernel (tcgen05 implicit GEMM, generic and row-shift variants)",
                    "bound": "tensor", "achieved": ach,
                    "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                    "traffic": tr,
                    "traffic_note": "dram__bytes_read+write per launch, mean over the conv launches of one ncu --set full "
                                    "capture (profiles/); algorithmic bytes per launch (in+out+weights) = "
                                    f"{c.get('bytes', 0.0) / max(c['n'], 1):.3e}",
                    "launches": c["n"], "avg_launch_ms": c["ms"] / c["n"],
                    "peak_source": peaks["source"] + ", sustained bf16"}
        else:
            c = dom[1]
            roof = {"kernel": dom[0], "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": None, "traffic": None, "launches": c["n"], "avg_launch_ms": c["ms"] / c["n"]}
        fc = tot.get("pmu_fcomb_softmax_accum_bf16")
        if fc and fc["ms"] > 0 and roof is not None:
            # the second kernel of the step, against the tensor pipe (its bound once the activations stay in tensor
            # memory); tmem_read_gbs next to the MEASURED TMEM read port (scripts/tmem_ld_bench.cu, 16 warps:
            # ~475 B/clk/SM of register bytes, profiles/r02_experiments.txt) shows that port is not the limit
            mhz = (clocks or {}).get("sm_mhz") or 1965
            tf = fc["flops"] / (fc["ms"] * 1e-3) / 1e12
            roof["second_kernel"] = {"kernel": "fcomb kernel (tcgen05 N-sample fcomb + softmax + sum / sum^2)",
                                     "ms_per_step": fc["ms"], "launches": fc["n"], "bound": "tensor",
                                     "achieved": tf, "unit": "TFLOP/s", "peak": peaks["bf16_sustained"],
                                     "frac": tf / peaks["bf16_sustained"], "tensor_frac": tf / peaks["bf16_sustained"],
                                     "tmem_read_gbs": fc["tmem"] / (fc["ms"] * 1e-3) / 1e9,
                                     "tmem_read_port_gbs_measured": 148 * 475.0 * mhz * 1e6 / 1e9,
                                     "peak_source": peaks["source"] + ", sustained bf16; TMEM port: 148 SMs x 475 B/clk (measured, "
                                                    f"16 warps) x {mhz} MHz"}
        world_frac = 1.0 / world
        V = float(D) ** 3
        if "pmu_slice_gather" in tot:
            gb = P * V * 8.0 * world_frac / 1e9          # read 4 B + write 4 B per voxel per plane
            # in-step form, measured honestly: the three whole-plane launches of one volume, eagerly launched back to back
            # behind an L2 flush (a 512 MB fill, so neither the volume nor the outputs are L2-resident), ONE event pair
            # around the three of them.  (Bracketing each ~25 us launch with its own pair mostly measures launch gaps.)
            mx = ops.plane_max(vol)
            offs = [0, D, 2 * D]
            outs3 = [torch.empty(D, 1, D, D, dtype=torch.float32, device=vol.device) for _ in range(3)]
            flush = torch.empty(128 << 20, dtype=torch.float32, device=vol.device)
            mxs = [mx[offs[p]:offs[p] + D].contiguous() for p in range(3)]
            times = []
            for _ in range(5):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for p in range(3):
                    ops.slice_gather(vol, p, 0, D, slice_max_in=mxs[p], out=outs3[p])
                e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            del flush, outs3
            times.sort()
            in_step = P * V * 8.0 / 1e9 / (times[len(times) // 2] * 1e-3)
            # Kernel throughput without launch gaps: whole-plane gathers of all three planes replayed back to back from
            # a CUDA graph, rotating over 4 volumes + 4 outputs (256 MB + 256 MB >> the 126 MB L2), one event pair
            # around the whole region.
            try:
                vols = [vol] + [vol.clone() for _ in range(3)]
                outs = [torch.empty(D, 1, D, D, dtype=torch.float32, device=vol.device) for _ in range(4)]

                def gather_round():
                    for i in range(4):
                        for p in range(3):
                            ops.slice_gather(vols[i], p, 0, D, slice_max_in=mxs[p], out=outs[(i + p) % 4])
                gather_round(); torch.cuda.synchronize()
                gs = torch.cuda.Stream()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.stream(gs):
                    gather_round()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(graph, stream=gs):
                        gather_round()
                    graph.replay(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    reps = 10
                    e0.record(gs)
                    for _ in range(reps):
                        graph.replay()
                    e1.record(gs); torch.cuda.synchronize()
                gbm = reps * 4 * 3 * V * 8.0 / 1e9
                ach = gbm / (e0.elapsed_time(e1) * 1e-3)
                del vols, outs
            except Exception as ex:  # noqa
                ach = None
            hbm_kernels["slice_gather"] = {"achieved": in_step, "unit": "GB/s", "peak": peaks["hbm_gbs"], "algorithmic_gb": gb,
                                           "how": "the three whole-plane launches of one volume, eager, behind a 512 MB L2 flush, ONE "
                                                  "event pair around the three (median of 5); normalisation fused",
                                           "graph_replay_gbs": ach,
                                           "graph_replay_how": "CUDA-graph replay of 120 whole-plane launches over 4 volumes (working set "
                                                               "512 MB > L2), one event pair: the kernel without launch gaps"}
        if "pmu_scatter_accum" in tot:
            gb = P * V * 2 * 3 * 4.0 * 3.0 * world_frac / 1e9  # read sums + RMW (read+write) accumulators, 2*C floats/voxel
            hbm_kernels["scatter_accum"] = {"achieved": gb / (tot["pmu_scatter_accum"]["ms"] * 1e-3), "unit": "GB/s",
                                            "peak": peaks["hbm_gbs"], "algorithmic_gb": gb}
        if "pmu_conv3x3_first_bf16" in tot:
            # the two 1 -> 64 first layers (U-Net and prior): 4 B in + 128 B (64 bf16 channels) out per pixel, output-bound
            gb = 2.0 * P * V * (4.0 + 128.0) * world_frac / 1e9
            hbm_kernels["first_conv"] = {"achieved": gb / (tot["pmu_conv3x3_first_bf16"]["ms"] * 1e-3), "unit": "GB/s",
                                         "peak": peaks["hbm_gbs"], "algorithmic_gb": gb}
        if "pmu_fuse_finalize" in tot:
            gb = V * 52.0 * (world_frac if slab_res else 1.0) / 1e9      # rank 0's x-slab with slab-sharded outputs
            hbm_kernels["fuse_finalize"] = {"achieved": gb / (tot["pmu_fuse_finalize"]["ms"] * 1e-3), "unit": "GB/s",
                                            "peak": peaks["hbm_gbs"], "algorithmic_gb": gb}
        # the general resampling kernel (TMA-staged brick) on an oblique grid, timed on its own: the
        # standard-plane grids of the workload take the bit-identical exact-slicing fast path
        try:
            aff = [0.5, -0.25, 0.75, 1.0, 0.01, 0.0, -0.01, 1.0, 0.02, 0.0, -0.02, 1.0]
            for _ in range(2):
                ops.slice_gather(vol, 0, 0, D, interp="trilinear", affine=aff, hw=(D, D))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.slice_gather(vol, 0, 0, D, interp="trilinear", affine=aff, hw=(D, D))
            e1.record(); torch.cuda.synchronize()
            gb = V * 8.0 / 1e9
            hbm_kernels["slice_gather_trilinear_oblique"] = {"achieved": gb / (e0.elapsed_time(e1) / 3 * 1e-3), "unit": "GB/s",
                                                             "peak": peaks["hbm_gbs"], "algorithmic_gb": gb}
        except Exception as ex:  # noqa
            hbm_kernels["slice_gather_trilinear_oblique"] = {"error": str(ex)[:200], "achieved": 0.0, "peak": peaks["hbm_gbs"]}
        for v in hbm_kernels.values():
            v["frac"] = v["achieved"] / v["peak"]
        if "slice_gather" in hbm_kernels:
            g = [traffic.get(k, {}).get("dram_bytes_per_launch") for k in ("gather_rows_kernel", "gather_plane2_kernel")]
            hbm_kernels["slice_gather"]["traffic_per_plane_launch"] = [x for x in g if x]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores, n_done = cpu_reference_rate(D, N, args.cpu_slices_per_plane, threads=os.cpu_count())
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_done} of {3 * D} slices ({args.cpu_slices_per_plane} equally spaced per plane), "
                                  f"forward once + {N} x fcomb + softmax + accumulate, {dt:.1f} s of CPU time, extrapolated"}
        va, va_dt = cpu_variant_a_rate(D, N, threads=os.cpu_count())
        cpu_baseline["variant_a"] = {"value": va, "unit": UNIT, "seconds": va_dt,
                                     "what": f"for information: literal eval.py:148-152, {N} full predict() calls per slice, "
                                             f"middle slice of each plane, extrapolated"}
    if rank == 0:
        value = args.steps / (ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": config_dict(D, N, args.interp),
                "run": {"slice_batch": args.slice_batch, **({"cuda_graph": True} if use_graph else {}),
                        "parallelism": f"slice-sharded x{world} + 1 " + ("reduce-scatter along x (x-slab outputs per rank)" if slab_res else "reduce")},
                "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                        # every rank uploads the volume (--e2e-upload broadcast: rank 0 alone, then NVLink)
                        "h2d_bytes_per_step": int(4 * D ** 3) * (1 if (slab and args.e2e_upload == "broadcast") else world),
                        "d2h_bytes_per_step": int(4 * D ** 3 * 7),              # mean + var (3 classes each) + entropy, whole job
                        "ms_per_step": ms_e2e / args.steps,
                        "how": "MultiPlanarPredictor.submit()/wait(): pinned host volume in, pinned host mean/var/entropy out, "
                               "every step; copies of neighbouring steps overlap compute (3 streams, 2 buffer slots); all "
                               "results complete in host memory before the timed region ends",
                        "sync_ms_per_step": ms_e2e_sync / args.steps,
                        "sync_note": "MultiPlanarPredictor.predict(host_out=...): one volume at a time, no overlap across steps",
                        "outputs": "x-slab per rank (reduce-scatter)" if slab else "rank 0 (streamed behind the last view)" if world == 1 else "rank 0",
                        **({"upload": args.e2e_upload} if slab else {}), **({"phases_ms": phases} if phases else {})},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "hbm_kernels": hbm_kernels,
                "kernel_time_shares": shares, "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
def run_config4(args):
    """BASELINE config 4: data-parallel training step, 8 slices of 256 x 256 per GPU (global batch 64 on 8 GPUs), GEMMs on
    tcgen05 in bf16, gradients SUM-all-reduced over NCCL in buckets (pmu_b200.dp_train_step).  One JSON line: slices/s,
    ms/step (CUDA events, max over ranks), per-kernel time shares of one instrumented step, replica consistency."""
    import torch
    import torch.distributed as dist
    import pmu_b200
    from pmu_b200 import ops
    from pmu_b200.synthetic import phantom_volume, phantom_labels
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.train_batch
    torch.manual_seed(0)
    trainer = pmu_b200.ProbUNetTrainer(dev, n_channels=1, n_classes=3, latent_dim=6, beta=10, precision=args.train_precision)
    net = trainer.net.train()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
    D = 256
    vol, lab = phantom_volume(D, seed=3), phantom_labels(D)
    s0 = 40 + rank * B                                   # every rank its own shard of x-slices (mixed content)
    mx = vol[s0:s0 + B].amax(dim=(1, 2), keepdim=True)
    imgs = (vol[s0:s0 + B] / mx)[:, None].contiguous().to(dev)
    masks = lab[s0:s0 + B, None].contiguous().to(dev)

    def step():
        return pmu_b200.dp_train_step(trainer, imgs, masks, opt, graph=not args.no_graph)

    for _ in range(max(args.warmup, 6)):                 # the caching allocator still grows during the first five steps
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.LAUNCHES
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    launches = ops.LAUNCHES - l0
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    gl = loss.clone()
    chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum()
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(gl)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ops.PROFILE = []
    pmu_b200.dp_train_step(trainer, imgs, masks, opt, graph=False); torch.cuda.synchronize()     # eager: per-call events
    tot = {}
    for name, meta, a, b in ops.PROFILE:
        tot[name] = tot.get(name, 0.0) + a.elapsed_time(b)
    launches = max(launches, len(ops.PROFILE) * args.steps)      # a graph replay launches the same kernels: count them, not the replays
    ops.PROFILE = None
    ssum = sum(tot.values())
    if rank == 0:
        fwd = 96.18 + 2 * 33.90 + 2 * 1.69            # GFLOP per 256^2 slice: U-Net + prior + posterior + two fcomb passes
        flop = B * world * (fwd + 2 * (fwd - 1.69)) * 1e9
        print(json.dumps({"metric": "slices/sec (data-parallel training step, 256x256 slices)", "value": B * world / float(ms) * 1e3,
                          "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 6),
                          "ms_per_step": float(ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": args.train_precision, "data": "synthetic",
                          "config": {"workload": f"BASELINE config 4: DP training step, {B} slices of 256x256 per GPU (global batch "
                                                 f"{B * world}), trainer model, SGD + clip, gradient all-reduce over NCCL"},
                          "cuda_graph": not args.no_graph,
                          "grad_exchange": ({"in_graph": True, "overlapped_with_backward": True, "ranges_mb": [round((hi_ - lo_) * 4 / 1e6, 1) for lo_, hi_ in trainer._graph_step[1].ar_ranges]}
                                            if (not args.no_graph and getattr(trainer, "_graph_step", None) and trainer._graph_step[1].ar_in_graph)
                                            else {"in_graph": False, "how": "after backward, flattened buckets" if world > 1 else "single GPU: none"}),
                          "useful_tflops": flop / float(ms) / 1e9, "global_loss": float(gl), "replicas_identical": bool(lo == hi),
                          "gpu_launches": launches,
                          "kernel_time_shares": {k: round(v / ssum, 4) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]},
                          "kernel_ms_instrumented_step": round(ssum, 2)}), flush=True)
    if world > 1:
        gs = getattr(trainer, "_graph_step", None)
        if gs is not None:
            gs[1].close()                 # the graph holds captured NCCL operations: release it before the communicator goes
            trainer._graph_step = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE config 5: sample-count sweep N = 1 ... 128 on a synthetic 512^3 volume with voxel entropy maps, slab-sharded
    over the ranks (reduce-scatter along x, every rank finalises its x-slab).  One JSON line per N."""
    import torch
    import torch.distributed as dist
    import pmu_b200
    from pmu_b200.synthetic import phantom_volume, trainer_state_dict
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    D, P = args.sweep_size, 3
    sd = trainer_state_dict(seed=0)
    vol = phantom_volume(D, seed=1234).to(dev)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for N in [int(x) for x in args.sweep_samples.split(",")]:
        pred = pmu_b200.MultiPlanarPredictor(sd, dev, precision=args.precision, n_samples=N, slice_batch=16, interp=args.interp,
                                             rank=rank, world_size=world, output="slab" if world > 1 else "rank0")
        eps = torch.randn(P, D, N, 6, generator=torch.Generator(device=dev).manual_seed(4321), device=dev)
        out = pred.predict(vol, eps=eps)                         # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.sweep_steps):
            out = pred.predict(vol, eps=eps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / args.sweep_steps], device=dev)
        ent = out["entropy"]
        stats = torch.stack([ent.sum(), ent.max(), out["var"].max(), torch.tensor(float(ent.numel()), device=dev)])
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            mx = stats[1:3].clone()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            stats[1:3] = mx
        if rank == 0:
            tflop = 3 * D * (130.61 + 1.099 * N) * (D / 256.0) ** 2 / 1e3
            assert int(stats[3]) == D ** 3, "entropy slabs do not tile the volume"
            tf = tflop / float(ms) * 1e3
            print(json.dumps({"metric": f"volumes/sec ({D}^3, 3 planes x N samples, entropy map)", "n_samples": N, "n_gpus": world,
                              "value": 1e3 / float(ms), "unit": "volumes/s", "ms_per_step": float(ms),
                              "config": {"workload": f"BASELINE config 5: {D}^3 x 3 planes x {N} samples, entropy map, "
                                                     f"{'x-slab outputs (reduce-scatter)' if world > 1 else 'single output'}"},
                              "dtype": args.precision, "algorithmic_tflop": round(tflop, 1), "tflops": round(tf, 1),
                              "frac_of_sustained_peak_per_gpu": round(tf / world / peaks["bf16_sustained"], 4),
                              "entropy_mean": round(float(stats[0]) / D ** 3, 5), "entropy_max": round(float(stats[1]), 5),
                              "var_max": round(float(stats[2]), 5)}), flush=True)
        del pred, eps, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--slice-batch", type=int, default=64)
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16", "fp32"],
                    help="f16: tensor-core mode, IEEE-half operands + fp32 accumulation (the inference format); bf16: the same kernels on bfloat16")
    ap.add_argument("--interp", default="trilinear", choices=["exact", "nearest", "trilinear"],
                    help="slice resampling onto the three standard plane grids (BASELINE configs[2]: trilinear)")
    ap.add_argument("--cpu-slices-per-plane", type=int, default=32,
                    help="cpu_baseline sample: equally spaced slices per plane (32 -> 96 of 768 slices, ~10 s on 16 threads)")
    ap.add_argument("--ref-slices-per-plane", type=int, default=8,
                    help="--impl reference: equally spaced slices per plane per step (8 -> 24 slices, ~5 s per step on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true", help="profiling aid: only the warm-up and the timed resident steps")
    ap.add_argument("--resident-output", default="slab", choices=["slab", "rank0"],
                    help="N > 1, resident step: reduce-scatter the accumulators along x and let every rank finalise its x-slab "
                         "(default), or sum-reduce them to rank 0")
    ap.add_argument("--graph", action="store_true",
                    help="replay the slice pass of a volume as one CUDA graph in the RESIDENT step too (the e2e leg for N > 1 does by default)")
    ap.add_argument("--no-graph", action="store_true", help="N > 1 e2e leg: eager launches instead of the CUDA-graph replay (the default there)")
    ap.add_argument("--e2e-upload", default="each", choices=["each", "broadcast"],
                    help="N > 1 e2e leg: every rank uploads the volume over its own PCIe link (default), or rank 0 uploads "
                         "once and broadcasts over NVLink (experiment)")
    ap.add_argument("--e2e-phases", action="store_true",
                    help="diagnostic: also time the e2e leg's host->device copy, device->host copy and resident slab step "
                         "on their own (all ranks at once), reported under e2e.phases_ms")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="3 (default): the headline metric (BASELINE configs[2]); 4: data-parallel training step (slices/s); "
                         "5: sample-count sweep on 512^3 (one JSON line per N)")
    ap.add_argument("--train-batch", type=int, default=8, help="--config 4: slices per GPU")
    ap.add_argument("--train-precision", default="bf16", choices=["bf16", "fp32"], help="--config 4")
    ap.add_argument("--sweep-size", type=int, default=512, help="--config 5: volume edge")
    ap.add_argument("--sweep-samples", default="1,2,4,8,16,32,64,128", help="--config 5")
    ap.add_argument("--sweep-steps", type=int, default=2, help="--config 5: timed predictions per N")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_config4(args)
    elif args.config == 5:
        run_config5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
