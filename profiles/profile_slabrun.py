"""Kernel timeline of a REAL multi-rank resident slab run (ncu cannot wrap a multi-rank job): every rank runs
SlabRun, rank 0 records its own kernels with the CUPTI-based torch profiler and prints per-kernel statistics.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
        profiles/profile_slabrun.py [steps] [graph|eager]
"""
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

import torch
import torch.distributed as dist

import bench
from turtlemd_b200.parallel import SlabRun

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
cfg = dict(rep=63, method="nlist", dt=0.005)
pos, vel, size, n, describe, make = bench.build_lj(cfg)
system, integ, pot = make()
system.evaluate(7)
run = SlabRun(system, integ)
run.prepare()
if graph:
    run.capture()
run.run(6)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        run.run(steps)
        torch.cuda.synchronize()
    events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    events.sort(key=lambda e: e.time_range.start)
    t0, t1 = events[0].time_range.start, events[-1].time_range.end
    busy = sum(e.time_range.end - e.time_range.start for e in events)
    print(f"world={world} steps={steps} graph={graph}: span {(t1 - t0) / steps:.1f} us per step, kernels busy "
          f"{busy / steps:.1f} us per step, {len(events) / steps:.1f} kernels per step")
    table = {}
    for e in events:
        name = e.name.split("(")[0][-48:]
        row = table.setdefault(name, [0, 0.0, 0.0])
        row[0] += 1
        d = e.time_range.end - e.time_range.start
        row[1] += d
        row[2] = max(row[2], d)
    for name, (cnt, tot, mx) in sorted(table.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:50s} n={cnt:5d} total/step={tot / steps:8.2f} us mean={tot / cnt:8.2f} max={mx:8.1f}")
    # gaps between consecutive kernels
    gaps = [events[i + 1].time_range.start - events[i].time_range.end for i in range(len(events) - 1)]
    print(f"  gaps: total/step={sum(g for g in gaps if g > 0) / steps:.2f} us, max={max(gaps):.1f}")
else:
    run.run(steps)
    torch.cuda.synchronize()
dist.barrier()
print(f"rank {rank}", run.status()) if rank == 0 else None
run.release()
dist.destroy_process_group()
