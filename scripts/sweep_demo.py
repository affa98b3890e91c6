#!/usr/bin/env python
"""BASELINE config 4: the 1024-replica scenario sweep (or its first n replicas), replica r on rank r mod G.

  python scripts/sweep_demo.py [n_replicas] [workers]                       one GPU
  python -m torch.distributed.run --nproc-per-node G ... scripts/sweep_demo.py [n] [workers]   G GPUs

Prints one JSON line (rank 0): replicas/s and aggregate agent-steps/s over all ranks, timed from a barrier
to the slowest rank's finish, host-side itinerary generation included; the per-replica statistics are
gathered on every rank at the end (`sweep.gather_results`) and a digest of them is printed so that runs
on different GPU counts can be compared."""
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from citam_b200 import sweep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))

import torch  # noqa: E402

torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
specs = sweep.make_sweep(1024)[:n]
sweep.run_sweep(specs[:2], device=local, workers=2)  # warm-up (library load, first allocations)
if dist is not None:
    dist.barrier()
t0 = time.perf_counter()
res = sweep.run_sweep(specs, device=local, workers=workers, world=world, rank=rank)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    res = sweep.gather_results(res)
if rank == 0:
    steps = sum(r["agent_steps"] for r in res)
    digest = hashlib.sha256(json.dumps([[r["replica"], r["n_pairs"], r["total_duration"], r["contact_steps"]] for r in res]).encode()).hexdigest()[:16]
    print(json.dumps({"workload": "config4 sweep", "replicas": len(res), "n_gpus": world, "workers_per_gpu": workers,
                      "seconds": dt, "replicas_per_s": len(res) / dt, "agent_steps_per_s": steps / dt,
                      "contact_steps": sum(r["contact_steps"] for r in res), "results_digest": digest}))
if dist is not None:
    dist.destroy_process_group()
