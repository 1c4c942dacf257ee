"""Multi-GPU parity check, launched by torch.distributed.run (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py [slab|atoms] [rep] [steps] [graph] [vv|langevin|dimer]

Every rank steps its share of one LJ fluid with the chosen decomposition; rank 0 also runs the
plain single-GPU integrator on the same input and compares positions, velocities, forces, energy
and virial.  Prints one line `MULTI_GPU_CHECK ok ...` (or `FAILED ...`) on rank 0; exit code 1 on
failure.  tests/test_gpu_parity.py runs it when the box has two or more GPUs."""
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

import numpy as np
import torch
import torch.distributed as dist

import bench
from turtlemd_b200.parallel import AtomDecomposition, SlabDecomposition, SlabRun


def main():
    import faulthandler

    faulthandler.dump_traceback_later(int(os.environ.get("TMD_CHECK_WATCHDOG", "240")), exit=True)  # never hang a GPU box
    kind = sys.argv[1] if len(sys.argv) > 1 else "slab"
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    graph = len(sys.argv) > 4 and sys.argv[4] == "graph"
    physics = sys.argv[5] if len(sys.argv) > 5 else "vv"  # langevin: config #4 in small; dimer: config #5
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = dict(rep=rep, method="nlist", dt=0.005, dimer=physics == "dimer",
               integrator="langevin" if physics == "langevin" else None)
    pos, vel, size, _, _, build = bench.build_lj(cfg)

    def make():
        return build()[0]

    system, integrator, _ = build()
    if kind == "slabrun":  # resident slab run: peer stores + device-side barriers, no NCCL inside a step
        system.potential_and_force()
        dec = SlabRun(system, integrator)
        dec.prepare()
        done = 0
        if graph:
            dec.capture()
        first = (steps - done) // 2  # two runs: a second start() after a completed run
        dec.run(first)
        dec.run(steps - done - first)
    else:
        dec = (SlabDecomposition if kind == "slab" else AtomDecomposition)(system, integrator)
        dec.prepare()
        done = 0
        if graph:
            if not dec.capture(warm_steps=3):
                raise RuntimeError("CUDA graph capture failed")
            done = 3
        for _ in range(steps - done):
            dec.step()
    x, v, f = dec.gather_state()
    v_pot, virial = dec.reduce_observables()
    ok, text = True, ""
    if rank == 0:
        ref, integ, _ = build()
        ref.potential_and_force()
        for _ in range(steps):
            integ(ref)
        scale = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))  # noqa: E731
        errs = {"pos": scale(x, ref.particles.pos), "vel": scale(v, ref.particles.vel),
                "force": scale(f, ref.particles.force),
                "v_pot": abs(v_pot - ref.particles.v_pot) / abs(ref.particles.v_pot),
                "virial": scale(virial, ref.particles.virial)}
        ok = errs["pos"] < 1e-11 and errs["vel"] < 1e-9 and errs["force"] < 1e-9 and errs["v_pot"] < 1e-10 \
            and errs["virial"] < 1e-9
        extra = f" rebuilds={dec.rebuilds} halo={dec.halo}" if kind == "slab" else ""
        if kind == "slabrun":
            extra = f" status={dec.status()}"
        text = (f"MULTI_GPU_CHECK {'ok' if ok else 'FAILED'} kind={kind} physics={physics} world={world} n={len(pos)} steps={steps} "
                f"graph={graph}{extra} " + " ".join(f"{k}={e:.2e}" for k, e in errs.items()))
        print(text, flush=True)
    if graph or kind == "slabrun":
        dec.release()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
