#!/usr/bin/env python
"""BASELINE config 5: sample-count sweep N = 1 ... 128 on a synthetic D^3 volume (default 512^3) with voxel entropy
maps, slab-sharded across the ranks it is launched on.

    python scripts/sweep_samples.py [--size 512] [--samples 1,2,4,...,128] [--steps 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep_samples.py

Per N: one warm-up + `--steps` timed predictions (volume resident in HBM, outputs left in HBM, CUDA events, max over
ranks).  With more than one rank the slices are sharded in index_map order and the accumulators are reduce-scattered
along x (every rank finalises the mean / variance / entropy of its own x-slab: MultiPlanarPredictor(output="slab")).
Prints one JSON line per N: ms / volume, volumes/s, TFLOP/s against the algorithmic work of SURVEY.md §8d
(130.61 + 1.099 N GFLOP per 256^2 slice, scaled by (D/256)^2), and the entropy / variance maps' statistics."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--samples", default="1,2,4,8,16,32,64,128")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--slice-batch", type=int, default=16)
    args = ap.parse_args()
    import torch.distributed as dist
    import pmu_b200
    from pmu_b200.synthetic import phantom_volume, trainer_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    D, P = args.size, 3
    sd = trainer_state_dict(seed=0)
    vol = phantom_volume(D, seed=1234).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for N in [int(x) for x in args.samples.split(",")]:
        pred = pmu_b200.MultiPlanarPredictor(sd, dev, precision="f16", n_samples=N, slice_batch=args.slice_batch,
                                             interp="trilinear", rank=rank, world_size=world,
                                             output="slab" if world > 1 else "rank0")
        eps = torch.randn(P, D, N, 6, generator=torch.Generator(device=dev).manual_seed(4321), device=dev)
        out = pred.predict(vol, eps=eps)                         # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = pred.predict(vol, eps=eps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
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
            print(json.dumps({"config": f"{D}^3 x 3 planes x {N} samples, entropy map, {world} GPU(s), "
                                        f"{'x-slab outputs (reduce-scatter)' if world > 1 else 'single output'}",
                              "n_samples": N, "n_gpus": world, "ms_per_volume": round(float(ms), 2),
                              "volumes_per_s": round(1e3 / float(ms), 4), "algorithmic_tflop": round(tflop, 1),
                              "tflops": round(tflop / float(ms) * 1e3, 1),
                              "entropy_mean": round(float(stats[0]) / D ** 3, 5), "entropy_max": round(float(stats[1]), 5),
                              "var_max": round(float(stats[2]), 5)}), flush=True)
        del pred, eps, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
