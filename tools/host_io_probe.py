#!/usr/bin/env python
"""What bounds the end-to-end path at N GPUs?  Every rank moves 4K frames (24.9 MB) between pinned host memory and
its GPU: H2D alone, D2H alone, both at once -- first one rank at a time, then all ranks together.  Run under torchrun.
Prints one JSON object (rank 0): GB/s per rank and aggregate, plus the PCIe / NUMA topology nvidia-smi reports."""
import json, os, subprocess, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, N = 2160, 3840, 12
host_in = torch.randint(0, 255, (N, H, W, 3), dtype=torch.uint8).pin_memory()
host_out = torch.empty((4, H, W, 3), dtype=torch.uint8).pin_memory()
dev = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(4)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
frame_gb = H * W * 3 / 1e9

def run(mode, frames=160):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(frames):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s_in):
                dev[i & 1].copy_(host_in[i % N], non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s_out):
                host_out[i & 3].copy_(dev[2 + (i & 1)], non_blocking=True)
    torch.cuda.synchronize()
    return frames * frame_gb / (time.perf_counter() - t0)     # GB/s per direction

out = {"world": world, "frame_MB": frame_gb * 1e3}
for mode in ("h2d", "d2h", "both"):
    run(mode, 20)
    # one rank at a time
    solo = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):
        dist.barrier()
        if r == rank:
            solo[r] = run(mode)
        dist.barrier()
    dist.all_reduce(solo)
    # all ranks together
    dist.barrier()
    together = torch.zeros(world, dtype=torch.float64, device="cuda")
    together[rank] = run(mode)
    dist.barrier()
    dist.all_reduce(together)
    out[mode] = {"alone_GBs_per_rank": [round(float(v), 1) for v in solo],
                 "together_GBs_per_rank": [round(float(v), 1) for v in together],
                 "together_aggregate_GBs": round(float(together.sum()), 1),
                 "frames_per_s_equivalent": round(float(together.sum()) / frame_gb)}
aff = sorted(os.sched_getaffinity(0))
allaff = [None] * world
dist.all_gather_object(allaff, (rank, len(aff), aff[0], aff[-1]))
if rank == 0:
    out["cpu_affinity(rank, n, first, last)"] = allaff
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            out[" ".join(cmd)] = [l for l in txt.splitlines() if l.strip()][:40]
        except Exception as e:
            out[" ".join(cmd)] = str(e)
    print(json.dumps(out, indent=1))
dist.destroy_process_group()
