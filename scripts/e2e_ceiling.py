"""Raw host<->device copy ceiling of the box for the e2e leg of bench.py: every rank (one per GPU, torchrun) copies the bench's
buffer sizes (default 4096 channels x 1500 blocks x 256 B = 1.57 GB each way) pinned host -> device on one stream and device ->
pinned host on another, at the same time, with NO kernel; timed with CUDA events after a barrier, max over ranks.  Also each
direction alone.  One cudaMemcpyAsync (torch copy_) per slice, as the engine does.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 scripts/e2e_ceiling.py
       (or plain `python scripts/e2e_ceiling.py` for one GPU)"""
import argparse
import json
import os

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mbytes", type=int, default=1500)
ap.add_argument("--slices", type=int, default=24)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
torch.cuda.set_device(lrank)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(lrank))
except Exception:
    pass
dev = torch.device("cuda", lrank)
n = a.mbytes * (1 << 20) // 4
h_in = torch.empty(n, dtype=torch.int32, pin_memory=True); h_in.fill_(rank + 1)
h_out = torch.empty(n, dtype=torch.int32, pin_memory=True); h_out.fill_(0)
d_in = torch.empty(n, dtype=torch.int32, device=dev)
d_out = torch.ones(n, dtype=torch.int32, device=dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
per = (n + a.slices - 1) // a.slices


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def run(h2d, d2h):
    e0, e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(torch.cuda.current_stream())
    s_in.wait_event(e0); s_out.wait_event(e0)
    for _ in range(a.reps):
        for k in range(a.slices):
            lo, hi = k * per, min(n, (k + 1) * per)
            if h2d:
                with torch.cuda.stream(s_in):
                    d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
    e1.record(s_in); e2.record(s_out)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    end = torch.cuda.Event(enable_timing=True); end.record(torch.cuda.current_stream())
    barrier()
    ms = e0.elapsed_time(end)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / a.reps


run(True, True)
res = {}
for name, (i, o) in {"h2d_only": (True, False), "d2h_only": (False, True), "full_duplex": (True, True)}.items():
    ms = run(i, o)
    gb = a.mbytes * (1 << 20) / 1e9
    res[name] = {"ms_per_pass": ms, "per_gpu_GBps_each_way": gb / (ms * 1e-3), "aggregate_GBps_each_way": world * gb / (ms * 1e-3)}
if rank == 0:
    # the bench's e2e unit: one pass moves 1.57 GB in and 1.57 GB out per GPU = 4096 x 1500 x 32 channel-samples
    fd = res["full_duplex"]
    res["n_gpus"] = world
    res["mbytes_each_way_per_gpu"] = a.mbytes
    res["e2e_ceiling_channel_samples_per_s"] = world * (a.mbytes * (1 << 20) / 8.0) / (fd["ms_per_pass"] * 1e-3)
    res["cpus"] = os.cpu_count()
    print(json.dumps(res), flush=True)
if dist is not None:
    dist.destroy_process_group()
