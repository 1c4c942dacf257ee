"""Diagnosis: host enqueue time vs device time of the fused resident run (N = 1 000 188)."""
import os, sys, time, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import lj_inputs, SINGLE
from turtlemd_b200.integrators import VelocityVerlet
from turtlemd_b200.potentials import LennardJonesCut
from turtlemd_b200.system import Box, Particles, System

rep = int(os.environ.get("REP", "63"))
pos, vel, size = lj_inputs(rep)
pot = LennardJonesCut(dim=3, shift=True, method="nlist"); pot.set_parameters(SINGLE)
system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel), [pot])
system.evaluate(7)
integ = VelocityVerlet(0.005)
integ.run(system, 10)
torch.cuda.synchronize()
for steps in (20, 100, 100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    integ.run(system, steps)
    t1 = time.perf_counter(); e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"steps {steps}: host enqueue {1e3*(t1-t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms ({e0.elapsed_time(e1)/steps:.4f} per step), wall {1e3*(t2-t0):.2f} ms", pot.nlist_stats(), flush=True)
