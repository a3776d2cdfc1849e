"""One timed full job of a bench workload on N GPUs (torchrun), after one warm-up job: the
multi-GPU band path of bench.py without its >= 3 warm-up steps, for the large configs.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
      tools/dist_probe.py cfg3"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import kit4b_b200 as k4b
from kit4b_b200 import hamm
from kit4b_b200.dist import CudaEngine, bands_slabwise, shard_bounds

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
k4b.gpu_init(1, [local])
concat, chroms, K, both = bench.synth_genome(wl)
L = len(concat); Nv = bench.valid_count(chroms, K)
engine = CudaEngine(dev)
if rank == 0:
    image, packed, non_acgt = engine.pack(concat, K)
    flag = torch.tensor([int(non_acgt)], dtype=torch.int64, device=dev)
else:
    image = engine.empty_image(L); flag = torch.zeros(1, dtype=torch.int64, device=dev)
if world > 1:
    dist.broadcast(flag, src=0); dist.broadcast(image, src=0)
if rank != 0:
    packed = engine.adopt(image, L, K, bool(flag.item()))
best = torch.empty(L, dtype=torch.int32, device=dev)
out = torch.empty(L, dtype=torch.int16, device=dev)
stream = torch.cuda.current_stream(dev)
qb, qe = shard_bounds(0, L, world)[rank]
def job():
    hamm.best_init_device(best.data_ptr(), L, K, stream.cuda_stream)
    hamm.diag_bootstrap_device(packed, both, qb, qe, best.data_ptr(), stream.cuda_stream)
    if world > 1:
        dist.all_reduce(best, op=dist.ReduceOp.MIN)
    bands_slabwise(engine, packed, both, rank, world, best)
    if rank == 0:
        hamm.best_finalize_device(packed, best.data_ptr(), out.data_ptr(), stream.cuda_stream)
res = {}
for name in ("warm", "timed"):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    job()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res[name + "_s"] = round(float(dt.item()), 3)
if rank == 0:
    S = 2 if both else 1
    res.update({"workload": wl, "K": K, "n_gpus": world, "kmers": int(Nv), "Gcmp_s": round(Nv * Nv * S / res["timed_s"] / 1e9, 1),
                "result_checksum": int(out.to(torch.int64).sum().item())})
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
k4b.gpu_shutdown()
