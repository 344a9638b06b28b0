"""Developer tool: concurrent host->device ingest bandwidth of all ranks of a node, with and without NUMA placement.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe.py
Each rank copies a pinned 1 GiB buffer to its GPU 20 times (one cudaMemcpyAsync per pass, then in 800 ragged pieces like
mesm_upload_clips); prints per-rank GB/s and the aggregate.  Run once with MESM_NO_NUMA_BIND=1 and once without."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mesm_b200  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = {"bound": False} if os.environ.get("MESM_NO_NUMA_BIND") else mesm_b200.bind_to_gpu_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
res = {}
for mode, pieces in (("one_copy", 1), ("ragged_800", 800)):
    step = n // pieces
    for it in range(2):
        for i in range(pieces):
            d[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for it in range(20):
        for i in range(pieces):
            d[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    res[mode] = 20 * n / (time.perf_counter() - t0) / 1e9
t = torch.tensor([res["one_copy"], res["ragged_800"]], device=dev)
if world > 1:
    allr = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
print(json.dumps({"rank": rank, "numa": numa, "gbs": res}), flush=True)
if rank == 0:
    a = torch.stack(allr).cpu()
    print(json.dumps({"world": world, "numa_bind": not os.environ.get("MESM_NO_NUMA_BIND"), "aggregate_one_copy_gbs": float(a[:, 0].sum()),
                      "aggregate_ragged_gbs": float(a[:, 1].sum()), "per_rank_one_copy": [round(float(x), 1) for x in a[:, 0]]}), flush=True)
if world > 1:
    dist.destroy_process_group()
