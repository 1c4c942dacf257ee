#!/usr/bin/env python
"""Benchmark of the hot path: atom-steps/s of FP64 Lennard-Jones force + integrate.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

Prints ONE JSON line (rank 0).  Workloads (BASELINE.json configs):
  lj1m      (default; the size BASELINE.json's metric is quoted at) LJ fluid N=1 000 188, rc=2.5,
            VelocityVerlet(0.005), Verlet neighbour list over a cell grid
  lj32k     configs[1]: LJ fluid N=32 000, all-pairs kernel
  lj1m_cells  same as lj1m with the stateless cell-list kernel
  lj1m_langevin  configs[3]: the same fluid under LangevinInertia(gamma=0.3, T=0.8), in-kernel Philox noise
  lj1m_langevin_default_rng  the same as a stock script writes it, LangevinInertia(seed=0): the reference's
            numpy default_rng(0) continued on the device (bit-exact PCG64 + ziggurat, csrc/npy_normal.cu)
  dimer256k   configs[4]: N=256 000 fluid, two bonded atoms of type 1 (DoubleWellPair) in LJ solvent
  replicas  configs[2]: 1M DoubleWell LangevinInertia replicas (replica-steps/s)
For --gpus N > 1 (launched by torch.distributed.run) the LJ workloads are atom-decomposed with an
all-gather of positions each step (TMD_BENCH_DECOMP=slab: slab decomposition with a halo exchange
instead -- measured slower at 2 and 8 GPUs in round 1), the replicas are
sharded with no communication.

value     device-resident throughput (state in HBM when the timed region starts)
e2e       the same step through the public Python API with HOST arrays: every step re-binds
          pos/vel/force from page-locked host memory (H2D inside the timed region) and reads
          pos/vel/force/v_pot back (D2H)
roofline  the dominant kernel against the FP64 pipe (all-pairs, cells) or HBM (replicas)
cpu_baseline  the CPU oracle (NumPy restatement of the reference, 1 core) on a bounded sample
configs   (N = 1, default workload only) one short device-resident measurement of every other BASELINE.json
          configuration -- lj32k, lj256k, dimer256k, lj1m_langevin (+ _default_rng), replicas -- so that none is a
          builder-run claim: value, ms_per_step, roofline.frac, clocks each
parity    (N > 1) after the timed loop rank 0 repeats the same number of steps on ONE GPU from the same
          initial state and compares positions, velocities, forces, V and W with the decomposed run

--impl reference: the UNMODIFIED reference package (baseline/_ref, see baseline/install_ref.py) on the host
CPU through its own public API (System / LennardJonesCut / VelocityVerlet, or DoubleWell / LangevinInertia
for `replicas`), each step fully measured on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

SINGLE = {0: {"sigma": 1.0, "epsilon": 1.0, "rcut": 2.5}}
RC, RHO = 2.5, 0.8
WORKLOADS = {
    "lj32k": dict(rep=20, method="allpairs", dt=0.005, unit="atom-steps/s"),
    "lj256k": dict(rep=40, method="nlist", dt=0.005, unit="atom-steps/s"),
    "lj1m": dict(rep=63, method="nlist", dt=0.005, unit="atom-steps/s"),
    "lj1m_cells": dict(rep=63, method="cells", dt=0.005, unit="atom-steps/s"),
    "lj1m_langevin": dict(rep=63, method="nlist", dt=0.005, unit="atom-steps/s", integrator="langevin"),
    # the same under LangevinInertia(seed=0) as a stock TurtleMD script writes it: the reference's default_rng(0),
    # continued on the device (tmd_npstream_*, csrc/npy_normal.cu) -- the reference's own noise, bit for bit
    "lj1m_langevin_default_rng": dict(rep=63, method="nlist", dt=0.005, unit="atom-steps/s", integrator="langevin",
                                      rgen="default_rng"),
    "dimer256k": dict(rep=40, method="nlist", dt=0.005, unit="atom-steps/s", dimer=True),
    "replicas": dict(n=1_000_000, dt=0.002, unit="replica-steps/s"),
}


def lj_inputs(rep):
    """Jittered FCC fluid of BASELINE.md section 3 (density 0.8, jitter 0.05, seeds 3 and 1)."""
    from turtlemd_b200.tools import generate_lattice

    pos, size = generate_lattice("fcc", [rep] * 3, density=RHO)
    pos = pos + 0.05 * (2.0 * np.random.default_rng(3).random(pos.shape) - 1.0)
    vel = np.random.default_rng(1).standard_normal(pos.shape)
    vel -= vel.mean(axis=0)
    vel *= math.sqrt(0.8 / (np.sum(vel * vel) / (3 * len(pos) - 3)))  # T = 0.8, as generate_maxwell_velocities
    return pos, vel, size


def algorithmic_flops(n, method):
    """21 flop per pair check + 36 per pair inside the cut-off (FMA = 2).  The pair
    checks are those of the reference algorithm (all pairs) for lj32k and of a half-stencil cell list
    with edge rc for the O(N) methods, whatever the kernel actually does."""
    hits = n * (4.0 / 3.0) * math.pi * RC**3 * RHO / 2.0
    checks = n * (n - 1) / 2.0 if method == "allpairs" else n * 27.0 * RC**3 * RHO / 2.0
    return 21.0 * checks + 36.0 * hits


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPUs run the benchmark's steps (B200_PROFILING.md).
    One sampler (rank 0) for all GPUs of the job.  nvidia-smi needs a second or more to start on an
    eight-GPU box and the timed region is milliseconds long, so the sampler is started before the system is
    built, `mark()` is called when the timed region starts, and the caller keeps issuing the same steps for
    ~1.5 s after it: `sm_mhz` is the median over the readings of that window."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices, enabled=True):
        self.indices = [int(i) for i in (indices if isinstance(indices, (list, tuple)) else [indices])]
        self.rows, self.proc, self.enabled, self.first = [], None, enabled, 0

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Readings from here on are the ones summarised (call when the GPUs start running the steps)."""
        self.first = len(self.rows)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows[self.first:] if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in rows)]
        return {"sm_mhz": statistics.median(float(r[0]) for r in rows), "sm_max_mhz": float(rows[0][1]),
                "reasons": reasons, "samples": len(rows), "gpus": self.indices,
                "window": "timed region + the same step loop continued for >= 1.5 s right after it (a timed region "
                          "of a few ms cannot be sampled at nvidia-smi's 100 ms period)"}


def dist_setup(gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def timed(fn, steps, world):
    """Barrier + sync, `steps` calls of fn timed with CUDA events, max over ranks -> seconds."""
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


# ---- the measured arms ---------------------------------------------------------------------------------


KERNEL_NAMES = {  # the kernels that run, as ncu prints them (profiles/)
    "allpairs": "lj_allpairs_sym_kernel<3,7,1>", "cells": "lj_cells_kernel<3>",
    "nlist": "nl_force_kernel<3,1,4,1,1,5,0>",
}


def build_lj(cfg, skin=None):
    """The benchmark system of a workload + a factory for an identical fresh copy (parity re-run)."""
    from turtlemd_b200.integrators import LangevinInertia, VelocityVerlet
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWellPair, LennardJonesCut
    from turtlemd_b200.system import Box, Particles, System

    pos, vel, size = lj_inputs(cfg["rep"])
    n = len(pos)
    extra = {} if skin is None else {"skin": skin}
    ptype, describe = None, ""
    if cfg.get("dimer"):
        # examples/movie-doublewell-pair.py at scale: two neighbouring atoms become type 1 and are bonded
        # by a DoubleWellPair; LJ parameters identical for both types (rc = 2.5)
        ptype = np.zeros(n, dtype=int)
        first = n // 2
        near = int(np.argsort(np.sum((pos - pos[first]) ** 2, axis=1))[1])
        ptype[[first, near]] = 1
        describe = " + DoubleWellPair dimer (types 0/1)"
    describe += (", LangevinInertia gamma=0.3 T=0.8" if cfg.get("integrator") == "langevin"
                 else f", VelocityVerlet dt={cfg['dt']}")
    if cfg.get("rgen") == "default_rng":
        describe += ", rgen = numpy default_rng(0) continued on the device (PCG64 + ziggurat, bit-exact)"

    def make():
        pot = LennardJonesCut(dim=3, shift=True, method=cfg["method"], **extra)
        pots = [pot]
        if cfg.get("dimer"):
            pot.set_parameters({0: dict(SINGLE[0]), 1: dict(SINGLE[0])})
            bond = DoubleWellPair(types=(1, 1), dim=3)
            bond.set_parameters({"rzero": 2.0 ** (1.0 / 6.0), "height": 6.0, "width": 0.25})
            pots.append(bond)
        else:
            pot.set_parameters(SINGLE)
        system = System(Box(low=size[:, 0], high=size[:, 1]), Particles.from_arrays(pos, vel=vel, ptype=ptype), pots)
        if cfg.get("rgen") == "default_rng":
            integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.8, seed=0)  # integrators.py:216
        elif cfg.get("integrator") == "langevin":
            integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.8, rgen=PhiloxNormal(key=0))
        else:
            integ = VelocityVerlet(cfg["dt"])
        return system, integ, pot

    return pos, vel, size, n, describe, make


def parity_lj(stepper, make, steps_done, rank):
    """Multi-rank parity on the record: the decomposed state after `steps_done` steps against ONE GPU taking
    the same steps from the same initial state (rank 0).  Tolerances as tests/multi_gpu_check.py: the two
    runs differ only in summation order, which a chaotic fluid amplifies slowly."""
    x, v, f = stepper.gather_state()
    v_pot, virial = stepper.reduce_observables()
    if rank != 0:
        return None
    ref, integ, _ = make()
    ref.potential_and_force()
    integ.run(ref, steps_done) if hasattr(integ, "run") else [integ(ref) for _ in range(steps_done)]
    q = ref.particles
    scale = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))  # noqa: E731
    errs = {"pos": scale(x, q.pos), "vel": scale(v, q.vel), "force": scale(f, q.force),
            "v_pot": float(abs(v_pot - q.v_pot) / abs(q.v_pot)), "virial": scale(virial, q.virial)}
    tol = {"pos": 1e-11, "vel": 1e-9, "force": 1e-9, "v_pot": 1e-10, "virial": 1e-9}
    return {"against": f"single-GPU {type(integ).__name__}.run of the same {steps_done} steps on rank 0",
            "steps": steps_done, "max_rel_err": errs, "tolerance": tol,
            "ok": bool(all(errs[k] < tol[k] for k in tol))}


def bench_lj(args, cfg, rank, world, extras=True):
    import ctypes as C

    import torch

    from turtlemd_b200 import _device, _lib

    pos, vel, size, n, describe, make = build_lj(cfg, args.skin)
    system, integ, pot = make()
    graphed, parallelism, stepper, decomp = False, "single GPU", None, None
    resident = False  # SlabRun: run(k) instead of step()
    if world > 1 or os.environ.get("TMD_BENCH_DECOMP") in ("slab1", "atoms1"):
        from turtlemd_b200.parallel import AtomDecomposition, SlabDecomposition, SlabRun

        # one LennardJonesCut under VelocityVerlet or LangevinInertia, or LennardJonesCut + DoubleWellPair under VelocityVerlet
        simple = cfg["method"] == "nlist" and not (cfg.get("dimer") and cfg.get("integrator") == "langevin")
        decomp = os.environ.get("TMD_BENCH_DECOMP", "slabrun" if simple and world > 1 else "atoms")
        if decomp == "slabrun" and simple and world > 1:
            try:
                system.evaluate(7)  # the run starts from valid forces, like VelocityVerlet.run
                stepper = SlabRun(system, integ)
                stepper.prepare()
                resident = True
                st = stepper.status()
                parallelism = (f"resident slab run x{world}: {st['layers_per_rank']} of {st['layers']} cell layers per rank, "
                               "boundary records stored into the neighbours' buffers by the traversal epilogue (peer "
                               "memory over NVLink), one device-side barrier per step, rebuilds / migration gated on "
                               "the device; no NCCL call inside a step")
            except NotImplementedError:
                stepper = None  # fewer than two cell layers per rank: replicate positions instead
                system, integ, pot = make()
        if decomp in ("slab", "slab1") and simple and not cfg.get("dimer"):
            try:
                stepper = SlabDecomposition(system, integ)
                stepper.prepare()
                parallelism = (f"slab decomposition x{world} (cell-order slot ranges), halo exchange of "
                               f"{stepper.halo} boundary records with the two neighbours per step (NCCL send/recv), "
                               "full exchange + re-binning at list rebuilds")
            except NotImplementedError:
                stepper = None  # slabs thinner than a halo: replicate positions instead
                system, integ, pot = make()
        if stepper is None:
            stepper = AtomDecomposition(system, integ)
            stepper.prepare()
            parallelism = f"atom-decomposition x{world}, NCCL all-gather of positions"
        step = stepper.step if not resident else None
    else:
        system.evaluate(7)
        step = lambda: integ(system)  # noqa: E731
    clocks = args.clocks
    if stepper is not None:
        graphed = stepper.capture() if os.environ.get("TMD_BENCH_GRAPH", "1") == "1" else False
        parallelism += ", steps replayed from CUDA graphs" if graphed else ", steps launched eagerly"
        if resident:
            stepper.run(max(args.warmup, 2))
        else:
            for _ in range(args.warmup):
                step()
    elif args.warmup > 0:
        integ.run(system, args.warmup)  # the entry point that is timed below (allocations, first list build, module load)
    clocks.mark()
    launches0 = _lib.launch_count()
    if stepper is None:
        # the multi-step entry point of the integrator: the closing velocity update of a step and the opening
        # update of the next one are a single pass over the state (same roundings, same trajectory)
        sec = timed(lambda: integ.run(system, args.steps), 1, world)
        parallelism += f", {type(integ).__name__}.run(system, {args.steps})"
    elif resident:
        sec = timed(lambda: stepper.run(args.steps), 1, world)
        parallelism += f", SlabRun.run({args.steps})"
        # safety net: a slab run that did not complete cleanly on EVERY rank (a barrier timed out, a capacity was
        # exceeded) is void -- say so and measure the atom decomposition instead of printing a wrong number
        bad = torch.tensor([stepper.status()["err"]], device="cuda", dtype=torch.int32)
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()) != 0:
            failed_status = stepper.status()
            stepper.release()
            system, integ, pot = make()
            stepper = AtomDecomposition(system, integ)
            stepper.prepare()
            graphed = stepper.capture()
            resident, step = False, stepper.step
            for _ in range(args.warmup):
                step()
            launches0 = _lib.launch_count()
            sec = timed(step, args.steps, world)
            parallelism = (f"atom-decomposition x{world}, NCCL all-gather of positions (the resident slab run FAILED on "
                           f"this box: {failed_status['error']})")
    else:
        sec = timed(step, args.steps, world)
    launches = _lib.launch_count() - launches0
    value = n * args.steps / sec

    out = {
        "metric": "atom-steps/sec (FP64 LJ force+integrate)", "value": value, "unit": cfg["unit"],
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"LJ fluid N={n} jittered FCC rho=0.8, rc=2.5 sigma, shifted{describe}, "
                               f"{cfg['method']} kernel",
                   "n_atoms": n, "parallelism": parallelism,
                   "scaling_note": "one system of fixed size whatever the number of GPUs (strong scaling)",
                   "l2": (f"positions {pos.nbytes / 1e6:.2f} MB are L2-resident by nature of the workload and the kernel "
                          "is FP64-pipe bound: no L2 flush between steps") if cfg["method"] == "allpairs" else
                         (f"a step streams the state ({6 * pos.nbytes / 1e6:.0f} MB) and, for the list method, "
                          "~300 MB of list entries: larger than L2, no flush needed")},
        "gpu_launches": launches,
    }

    def load_for_clocks():
        """The same steps for ~1.5 s more, so that nvidia-smi's 100 ms readings see the benchmark's load (the
        count follows from the max-over-ranks step time: identical on every rank)."""
        more = max(50, min(20000, int((1.5 if extras is True else 0.5) / (sec / args.steps))))
        if stepper is None:
            integ.run(system, more)
        elif resident:
            stepper.run(more)
        else:
            for _ in range(more):
                step()
        torch.cuda.synchronize()

    # ---- multi-GPU extras: parity against one GPU, e2e with host buffers --------------------------------
    if stepper is not None:
        if world > 1 and extras and not args.quick:
            out["parity"] = parity_lj(stepper, make, stepper.steps_done, rank)
            if resident:
                out["slab_run"] = stepper.status()
        load_for_clocks()
        out["clocks"] = clocks.summary()
        if resident and world > 1 and extras is True and not args.quick:
            # end to end with host buffers: through the atom decomposition, whose state is the plain atom-ordered
            # arrays a host shard can be copied into (the resident run keeps its state in cell order on the device)
            stepper.release()
            system2, integ2, _ = make()
            stepper = AtomDecomposition(system2, integ2)
            stepper.prepare()
            stepper.capture()
            out["e2e"] = e2e_decomposed(stepper, system2, n, cfg, world, args)
        elif world > 1 and extras is True and not args.quick:
            out["e2e"] = e2e_decomposed(stepper, system, n, cfg, world, args)
        stepper.release()
        return out if rank == 0 else None

    # ---- single-GPU extras: roofline of the dominant kernel, e2e, CPU baseline ---------------------
    lib = _lib.load()
    tf, ms = C.c_double(0), C.c_double(0)
    _lib.check(lib.tmd_bench_dfma(5, C.byref(tf), C.byref(ms)))
    peak = tf.value
    # the force evaluation alone, positions unchanged (for nlist: gather/check + traversal, no rebuild)
    evaluate = lambda: system.evaluate(7)  # noqa: E731
    for _ in range(3):
        evaluate()
    ksec = timed(evaluate, args.steps, 1) / args.steps
    flops = algorithmic_flops(n, cfg["method"])
    achieved = flops / ksec / 1e12
    kernel = KERNEL_NAMES[cfg["method"]]
    traffic = None
    try:
        table = json.loads((REPO / "profiles" / "traffic.json").read_text())
        traffic = table.get(f"{kernel}@{n}", {}).get("bytes")
    except (OSError, ValueError):
        pass
    step_bytes = 184.0 * n  # x, v, F read + written once, force-kernel reads, bookkeeping
    step_sec = sec / args.steps
    out["roofline"] = {
        "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": kernel,
        "hbm": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_bytes / step_sec / 1e9,
                "peak_gbs": peak_hbm(), "frac": step_bytes / step_sec / 1e9 / peak_hbm()},
        "algorithmic_flops_per_launch": flops, "launch_ms": ksec * 1e3,
        "step_frac": flops / step_sec / 1e12 / peak,
        "step_roofline_frac": (flops / (peak * 1e12) + step_bytes / (peak_hbm() * 1e9)) / step_sec,
        "note": "algorithmic flops = SURVEY.md 8(d): 21 per pair check of a half-stencil cell list (all pairs "
                "for lj32k) + 36 per pair inside the cut-off; launch_ms = one whole force evaluation "
                "(all kernels of System.evaluate), step_frac = the same flops over the full timed step, "
                "step_roofline_frac = SURVEY's T_roof = F/P + B N/BW over the full timed step",
        "peak_source": "DFMA-only micro-benchmark (tmd_bench_dfma) run in this process; "
                       "MEASURED_PEAKS.json has no FP64 entry (nominal 37.2 TFLOP/s)",
    }
    if cfg["method"] == "nlist":
        out["config"]["nlist"] = dict(pot.nlist_stats(), skin=pot.skin)
        if type(integ).__name__ == "VelocityVerlet":
            # the kernel that dominates the timed step is the FUSED traversal (forces + kick + kick + drift epilogue):
            # its own launches, timed live with CUDA events on the launching stream over another args.steps steps
            pot.time_traversals(True)
            integ.run(system, args.steps)
            tms, tcount = pot.traversal_time()
            pot.time_traversals(False)
            if tcount:
                fused = tms / tcount * 1e-3
                out["roofline"].update({
                    "kernel": "nl_force_kernel<3,1,4,1,1,5,1>", "launch_ms": fused * 1e3, "achieved": flops / fused / 1e12,
                    "frac": flops / fused / 1e12 / peak, "launches_timed": tcount,
                    "traffic": table.get(f"nl_force_kernel<3,1,4,1,1,5,1>@{n}", {}).get("bytes") if traffic is not None else None,
                    "force_evaluation": {"kernel": kernel, "launch_ms": ksec * 1e3, "frac": achieved / peak,
                                         "what": "System.evaluate: gather/check pass + bare traversal + reduction"},
                    "note": out["roofline"]["note"].replace(
                        "launch_ms = one whole force evaluation (all kernels of System.evaluate)",
                        "launch_ms = average duration of the fused traversal launches of a run (CUDA events around each "
                        "launch; the kernel also integrates), force_evaluation = one whole System.evaluate")})

    if args.quick or not extras:
        load_for_clocks()
        out["clocks"] = clocks.summary()
        return out
    # e2e: host arrays every step through the public API
    hp, hv, hf = (_device.pinned_array(pos.shape) for _ in range(3))
    part = system.particles
    np.copyto(hp, part.pos)
    np.copyto(hv, part.vel)
    np.copyto(hf, part.force)

    def e2e_step():
        part.pos, part.vel, part.force = hp, hv, hf  # re-bind host arrays: device copies are stale
        integ(system)
        _ = part.pos, part.vel, part.force, part.v_pot  # read back into the same page-locked arrays

    for _ in range(2):
        e2e_step()
    esteps = max(3, min(args.steps, 20))
    esec = timed(e2e_step, esteps, 1)
    out["e2e"] = {"value": n * esteps / esec, "unit": cfg["unit"],
                  "h2d_bytes_per_step": 3 * pos.nbytes, "d2h_bytes_per_step": 3 * pos.nbytes + 80,
                  "ms_per_step": esec / esteps * 1e3,
                  "api": f"{type(integ).__name__}.integration_step(system) with particles.pos/vel/force re-bound "
                         "from page-locked host arrays and read back every step"}
    integ.run(system, 3)  # the device copies are current again
    load_for_clocks()
    out["clocks"] = clocks.summary()
    out["cpu_baseline"] = cpu_baseline(cfg, pos, size, n)
    return out


def cpu_baseline(cfg, pos, size, n):
    """The reference itself on a bounded sample when baseline/_ref travelled with the repo (kind "reference":
    three fully measured steps, ~20 s), else the NumPy oracle's row loop (kind "port")."""
    port = cpu_baseline_lj(pos, size, n, cfg["method"], budget_s=6.0)
    try:
        if _reference_package() is not None:
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            base, _, _ = reference_lj_sample(cfg, 3, 1, 24.0, cores)
            base["port"] = port
            return base
    except Exception as exc:  # noqa: BLE001 - the port always exists
        port["reference_error"] = f"{type(exc).__name__}: {exc}"
    return port


def e2e_decomposed(stepper, system, n, cfg, world, args):
    """End to end at N > 1: every step each rank uploads its shard of pos / vel / force from page-locked host
    memory into the decomposed state, all ranks take the step, and each rank reads its shard (+ V, W) back."""
    import torch

    from turtlemd_b200.parallel import AtomDecomposition

    if not isinstance(stepper, AtomDecomposition):
        return {"value": None, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "not measured for this decomposition"}
    part, plan, dim = system.particles, stepper.plan, stepper.dim
    lo, cnt = plan.first * dim, plan.count * dim
    dev = [part._pos.device(), part._vel.device(), part._force.device()]
    host = [torch.empty(cnt, dtype=torch.float64, pin_memory=True) for _ in range(3)]
    vw_host = torch.empty(10, dtype=torch.float64, pin_memory=True)
    for h, d in zip(host, dev):
        h.copy_(d[lo: lo + cnt])
    torch.cuda.synchronize()

    def e2e_step():
        for h, d in zip(host, dev):
            d[lo: lo + cnt].copy_(h, non_blocking=True)
        stepper.step()  # (the step's own all-gather carries the uploaded positions to the other ranks)
        for h, d in zip(host, dev):
            h.copy_(d[lo: lo + cnt], non_blocking=True)
        vw_host.copy_(part._vw_buffer(), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    esteps = max(3, min(args.steps, 20))
    esec = timed(e2e_step, esteps, world)
    return {"value": n * esteps / esec, "unit": cfg["unit"], "h2d_bytes_per_step": 3 * 8 * n * dim,
            "d2h_bytes_per_step": 3 * 8 * n * dim + 80 * world, "ms_per_step": esec / esteps * 1e3,
            "api": "AtomDecomposition.step() with every rank's shard of pos/vel/force uploaded from page-locked host "
                   "memory before the step and read back after it (bytes summed over the ranks)"}


def _oracle_rows(job):
    """Worker: the body of the oracle's lj_potential_and_force row loop for a list of rows, for at most
    budget_s seconds.  Returns (pair checks done, seconds)."""
    pos, size, rows, budget_s = job
    from oracle import turtle_oracle as to

    n = len(pos)
    length = size[:, 1] - size[:, 0]
    table, _ = to.lj_tables(SINGLE)
    tidx = np.zeros(n, dtype=int)
    lj1, lj2, lj3, lj4, offset, rcut2 = table
    checks, t0, pot = 0, time.perf_counter(), 0.0
    for i in rows:
        delta, rsq, k = to._row(pos, tidx, int(i), length, 1.0 / length, [True] * 3, rcut2)
        checks += n - int(i) - 1
        if len(k):
            r2inv = 1.0 / rsq[k]
            r6inv = r2inv**3
            pot += np.sum(r6inv * (lj3[0, 0] * r6inv - lj4[0, 0]) - offset[0, 0])
            flj = r2inv * r6inv * (lj1[0, 0] * r6inv - lj2[0, 0])
            fij = np.einsum("i,ij->ij", flj, delta[k])
            np.einsum("ij,ik->jk", fij, delta[k])
        if time.perf_counter() - t0 > budget_s:
            break
    return checks, time.perf_counter() - t0


def cpu_baseline_lj(pos, size, n, method, budget_s=15.0, cores=1):
    """The NumPy oracle (restated reference row loop) on a strided sample of rows.  cores > 1: the rows
    are dealt to that many worker processes (rows are independent; the reference itself is one thread)."""
    stride = max(1, n // (640 * cores))
    rows = np.arange(0, n - 1, stride)
    if cores == 1:
        results = [_oracle_rows((pos, size, rows, budget_s))]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(cores) as pool:
            results = pool.map(_oracle_rows, [(pos, size, rows[w::cores], budget_s) for w in range(cores)])
    checks = sum(c for c, _ in results)
    rate = sum(c / s for c, s in results)  # pair checks per second of all workers running side by side
    sec = max(s for _, s in results)
    full = (n * (n - 1) / 2.0) / rate  # the reference is all-pairs whatever the GPU method
    return {"value": n / full, "unit": "atom-steps/s", "cores": cores, "kind": "port",
            "sample": f"NumPy oracle row loop on {checks:.3g} of {n * (n - 1) / 2:.3g} pair checks "
                      f"(every {stride}th row) in {sec:.1f} s on {cores} core(s), scaled to one full step; "
                      f"host has {os.cpu_count()} cores",
            "pair_checks_per_s": rate}


def bench_replicas(args, cfg, rank, world, extras=True):
    import ctypes as C

    import torch

    from turtlemd_b200 import _device, _lib
    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    n_global = cfg["n"] * world  # weak scaling: 1M replicas per GPU
    n, first = cfg["n"], rank * cfg["n"]
    x0 = np.full((n, 1), -1.0)
    system = System(Box(periodic=[False]), Particles.from_arrays(x0), [DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.evaluate(7)
    integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.07, rgen=PhiloxNormal(key=0))
    inner = 10  # integrator steps per launch of the fused kernel
    step = lambda: integ.run(system, inner, first=first, n_global=n_global)  # noqa: E731
    for _ in range(args.warmup):
        step()
    clocks = args.clocks
    clocks.mark()
    sec = timed(step, args.steps, world)
    value = n_global * inner * args.steps / sec
    out = {
        "metric": "replica-steps/sec (FP64 DoubleWell LangevinInertia)", "value": value, "unit": cfg["unit"],
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['n']} independent 1-D DoubleWell(1,2,0) LangevinInertia replicas per GPU, "
                               f"dt={cfg['dt']}, gamma=0.3, beta=1/0.07; one bench step = {inner} integrator steps "
                               "in one fused kernel", "replicas_per_gpu": n, "parallelism": f"replica sharding x{world}"},
        "gpu_launches": 2 * args.steps,
    }
    if world > 1 and extras and not args.quick:
        out["parity"] = parity_replicas(system, cfg, inner, args.warmup + args.steps, rank, world)
    for _ in range(max(10, min(20000, int((1.5 if extras is True else 0.5) / (sec / args.steps))))):  # load for the clock readings
        step()
    torch.cuda.synchronize()
    out["clocks"] = clocks.summary()
    if rank != 0:
        return None
    # 32 B per replica-step (x, v read + written), whatever the kernel does; the fused kernel
    # keeps a replica in registers for `inner` steps, so what it really moves is 48 B per replica per launch
    bytes_per_launch = 32.0 * n * inner
    out["roofline"] = {"bound": "hbm", "achieved": bytes_per_launch / (sec / args.steps) / 1e9,
                       "peak": peak_hbm(), "unit": "GB/s", "traffic": None,
                       "kernel": "inertia_doublewell_pair_kernel<1,true>",
                       "algorithmic_bytes_per_launch": bytes_per_launch,
                       "moved_bytes_per_launch_estimate": 48.0 * n,
                       "note": "algorithmic bytes = 32 B per replica-step x replica-steps per launch (SURVEY.md 8d); "
                               f"with {inner} steps fused per launch the kernel is bound by the FP64 / integer work "
                               "of the noise (Philox4x64 + log / sincos), not by HBM; see DESIGN.md"}
    out["roofline"]["frac"] = out["roofline"]["achieved"] / out["roofline"]["peak"]
    try:  # the bound ncu shows (profiles/r02p_replicas_full.txt): the instruction stream, not HBM
        prof = json.loads((REPO / "profiles" / "traffic.json").read_text())[f"inertia_doublewell_pair_kernel<1,true>@{n}x{inner}"]
        sm_hz = 1e6 * float((out["clocks"].get("sm_mhz") or 1965.0))
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        ideal = prof["warp_instructions"] / (sms * 4.0 * sm_hz)
        out["roofline"]["traffic"] = prof["bytes"]
        out["roofline"]["issue"] = {
            "warp_instructions_per_launch": prof["warp_instructions"], "ideal_ms": ideal * 1e3,
            "frac": ideal / (sec / args.steps),
            "note": "instruction-issue roofline: warp instructions of one launch (ncu) / (SMs x 4 per cycle x SM clock) over the "
                    "measured launch time; the kernel is bound by Philox4x64 integer products + FP64 log / sincos"}
    except (OSError, ValueError, KeyError):
        pass
    if world == 1 and not args.quick and extras:
        # e2e: host arrays in and out around every launch through the public API
        part = system.particles
        hx, hv, hf = (_device.pinned_array(x0.shape) for _ in range(3))
        np.copyto(hx, part.pos)
        np.copyto(hv, part.vel)
        np.copyto(hf, part.force)

        def e2e_step():
            part.pos, part.vel, part.force = hx, hv, hf
            integ.run(system, inner, first=first, n_global=n_global)
            _ = part.pos, part.vel, part.force, part.v_pot

        for _ in range(2):
            e2e_step()
        esteps = max(3, min(args.steps, 20))
        esec = timed(e2e_step, esteps, 1)
        out["e2e"] = {"value": n * inner * esteps / esec, "unit": cfg["unit"],
                      "h2d_bytes_per_step": 3 * x0.nbytes, "d2h_bytes_per_step": 3 * x0.nbytes + 80,
                      "ms_per_step": esec / esteps * 1e3,
                      "api": f"LangevinInertia.run(system, {inner}) with particles.pos/vel/force re-bound from "
                             "page-locked host arrays and read back after every call"}
        out["cpu_baseline"] = cpu_baseline_replicas(cfg)
    del C, _lib
    return out


def parity_replicas(system, cfg, inner, calls, rank, world, sample=4096):
    """Sharding must not change a replica's trajectory: rank 0 recomputes the first `sample` replicas of EVERY
    rank's block on its own GPU (same global indices, same number of steps) and compares bit for bit."""
    import torch
    import torch.distributed as dist

    from turtlemd_b200.integrators import LangevinInertia
    from turtlemd_b200.philox import PhiloxNormal
    from turtlemd_b200.potentials import DoubleWell
    from turtlemd_b200.system import Box, Particles, System

    part = system.particles
    mine = torch.stack([part._pos.device()[:sample], part._vel.device()[:sample]]).contiguous()
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)
    if rank != 0:
        return None
    n, bad = cfg["n"], 0
    for r in range(world):
        ref = System(Box(periodic=[False]), Particles.from_arrays(np.full((sample, 1), -1.0)),
                     [DoubleWell(a=1.0, b=2.0, c=0.0)])
        ref.evaluate(7)
        integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.07, rgen=PhiloxNormal(key=0))
        for _ in range(calls):
            integ.run(ref, inner, first=r * n, n_global=n * world)
        got = everyone[r].cpu().numpy()
        bad += int(np.sum(got[0] != ref.particles.pos[:, 0]) + np.sum(got[1] != ref.particles.vel[:, 0]))
    return {"against": f"rank 0 recomputing replicas [r*{n}, r*{n}+{sample}) of every rank r on one GPU, "
                       f"{calls * inner} steps", "mismatching_values": bad, "compared": 2 * sample * world,
            "ok": bad == 0, "tolerance": "bit-exact"}


def cpu_baseline_replicas(cfg, sample=20000, budget_s=10.0):
    """The NumPy oracle's LangevinInertia step (restated reference: one Python-level draw per particle,
    integrators.py:284-315) over a sample of the replicas, one core."""
    from oracle import turtle_oracle as to

    ff = to.ForceField(np.array([np.inf]), np.array([0.0]), [False], [("dw", (1.0, 2.0, 0.0))])
    st = to.State(np.full((sample, 1), -1.0), np.zeros((sample, 1)), np.ones(sample))
    st.force, st.virial = ff.force(st.pos)
    stepper = to.InertiaStepper(cfg["dt"], 0.3, 1.0 / 0.07, np.random.default_rng(0))
    stepper.step(st, ff)  # builds the per-particle constants
    steps, t0 = 0, time.perf_counter()
    while steps < 1 or time.perf_counter() - t0 < budget_s:
        stepper.step(st, ff)
        steps += 1
    sec = time.perf_counter() - t0
    return {"value": sample * steps / sec, "unit": "replica-steps/s", "cores": 1, "kind": "port",
            "sample": f"NumPy oracle LangevinInertia + DoubleWell on {sample} of {cfg['n']} replicas, {steps} steps in "
                      f"{sec:.1f} s (throughput does not depend on the ensemble size); host has {os.cpu_count()} cores"}


def peak_hbm():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        return json.loads(path.read_text())["hbm_gbs"]
    return 6650.0


def _reference_package():
    """The unmodified reference from baseline/_ref (baseline/install_ref.py); None when it is not there."""
    ref = REPO / "baseline" / "_ref"
    if not (ref / "turtlemd").is_dir():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import turtlemd  # noqa: F401  (its version.py finds the dist-info next to it)
    return ref


def _reference_particles(pos, vel):
    """A reference Particles filled by array assignment (add_particle is O(N^2): it vstacks per atom)."""
    from turtlemd.system import Particles

    n, dim = pos.shape
    part = Particles(dim=dim)
    part.pos, part.vel, part.force = np.array(pos), np.array(vel), np.zeros_like(pos)
    part.mass = np.ones((n, 1))
    part.imass = 1.0 / part.mass
    part.ptype = np.zeros(n, dtype=int)
    part.name = np.array(["Ar"] * n)
    part.npart = n
    return part


def bench_reference(args, cfg, rank, world):
    """--impl reference: the UNMODIFIED reference (baseline/_ref) through its own public API on the host CPU,
    rank 0 only, nothing here touches CUDA or this repo's kernels.  Every step is fully measured.  The
    reference evaluates all N (N - 1) / 2 pairs in Python (lennardjones.py:242-272: 180 s per step at
    N = 32 000, two days at N = 1M), so each step runs on a bounded sample of the workload: the same fluid
    (density, cut-off, temperature, time step) at the largest N whose warm-up + K steps fit in ~2 minutes.
    `value` is what was measured on that sample; the reference's cost per atom-step grows linearly with N, so
    at the b200 arm's N it is lower still by N / N_sample (`extrapolated_to_workload`)."""
    if rank != 0:
        return None
    if _reference_package() is None:
        return {"impl": "reference", "unavailable": "baseline/_ref is missing: run python baseline/install_ref.py "
                                                    "where /root/reference is mounted"}
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    total_steps = args.warmup + args.steps
    if "rep" not in cfg:
        return reference_replicas(args, cfg, world, cores, total_steps)
    base, n, sec_per_step = reference_lj_sample(cfg, args.steps, args.warmup,
                                                float(os.environ.get("TMD_REFERENCE_BUDGET_S", "120")), cores)
    value = base["value"]
    n_work = 4 * cfg["rep"] ** 3
    return {
        "impl": "reference", "metric": "atom-steps/sec (FP64 LJ force+integrate)", "value": value,
        "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"LJ fluid N={n_work} jittered FCC rho=0.8, rc=2.5 sigma, shifted, VelocityVerlet "
                               f"dt={cfg['dt']}", "n_atoms": n_work, "sample_atoms": n,
                   "note": "the reference's own all-pairs Python loop (lennardjones.py:242-272) cannot take a step of "
                           f"the workload's size in any budget; every step here is a full reference step at N = {n}"},
        "extrapolated_to_workload": base["extrapolated_to_workload"],
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def reference_lj_sample(cfg, steps, warmup, budget_s, cores):
    """`warmup` + `steps` VelocityVerlet steps of the unmodified reference on the largest sample of the
    workload's fluid that fits `budget_s`; returns (cpu_baseline dict, N of the sample, seconds per step)."""
    from turtlemd.integrators import VelocityVerlet
    from turtlemd.potentials import LennardJonesCut
    from turtlemd.system import Box, System

    def make(rep):
        pos, vel, size = lj_inputs(rep)
        pot = LennardJonesCut(dim=3, shift=True)
        pot.set_parameters(SINGLE)
        system = System(box=Box(low=size[:, 0], high=size[:, 1]), particles=_reference_particles(pos, vel),
                        potentials=[pot])
        system.potential_and_force()
        return system, len(pos)

    # calibrate on N = 864 (~0.4 s per step), then pick the sample size: step time ~ N^2
    probe, n0 = make(6)
    integ = VelocityVerlet(cfg["dt"])
    t0 = time.perf_counter()
    integ(probe)
    per_pair = (time.perf_counter() - t0) / (n0 * (n0 - 1) / 2)
    rep = 6
    for r in (7, 8, 9, 10, 11, 12, 13, 14, 16):
        n_r = 4 * r**3
        # (+2: the initial force evaluation and slack; larger N run at a better pair rate than the probe)
        if r <= cfg["rep"] and per_pair * n_r * (n_r - 1) / 2 * (warmup + steps + 2) <= budget_s:
            rep = r
    system, n = make(rep)
    integ = VelocityVerlet(cfg["dt"])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        integ(system)  # MDIntegrator.__call__ = integration_step (integrators.py:32-33)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times)
    value = n * len(times) / sec
    n_work = 4 * cfg["rep"] ** 3
    base = {"value": value, "unit": "atom-steps/s", "cores": 1, "kind": "reference",
            "sample": f"unmodified reference VelocityVerlet + LennardJonesCut at N = {n} (same fluid as the workload, "
                      f"N = {n_work}), {len(times)} fully measured steps of {sec / len(times):.2f} s on 1 core "
                      f"(the reference is single-threaded Python; host has {cores} cores)",
            "pair_checks_per_s": n * (n - 1) / 2 * len(times) / sec,
            "extrapolated_to_workload": {"value": value * n / n_work, "unit": "atom-steps/s",
                                         "rule": "pair checks per second as measured, all N (N - 1) / 2 pairs per step"}}
    return base, n, sec / len(times)


def reference_replicas(args, cfg, world, cores, total_steps):
    """The unmodified reference LangevinInertia + DoubleWell on a sample of the replicas (its throughput does not
    depend on the ensemble size: one Python-level draw per particle, integrators.py:284-297)."""
    from turtlemd.integrators import LangevinInertia
    from turtlemd.potentials.well import DoubleWell
    from turtlemd.system import Box, System

    sample = 20000
    per_step = 0.12  # s at 20 000 replicas
    while sample > 1000 and per_step * (sample / 20000) * 10 * total_steps > 110:
        sample //= 2
    system = System(box=Box(periodic=[False]), particles=_reference_particles(np.full((sample, 1), -1.0),
                                                                              np.zeros((sample, 1))),
                    potentials=[DoubleWell(a=1.0, b=2.0, c=0.0)])
    system.potential_and_force()
    integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.07, seed=0)
    inner, times = 10, []  # one bench step of the b200 arm = 10 integrator steps
    for it in range(total_steps):
        t0 = time.perf_counter()
        for _ in range(inner):
            integ(system)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times)
    value = sample * inner * len(times) / sec
    base = {"value": value, "unit": "replica-steps/s", "cores": 1, "kind": "reference",
            "sample": f"unmodified reference LangevinInertia + DoubleWell on {sample} of {cfg['n']} replicas, "
                      f"{inner * len(times)} fully measured steps on 1 core; host has {cores} cores"}
    return {
        "impl": "reference", "metric": "replica-steps/sec (FP64 DoubleWell LangevinInertia)", "value": value,
        "unit": "replica-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['n']} independent 1-D DoubleWell(1,2,0) LangevinInertia replicas per GPU, "
                               f"dt={cfg['dt']}, gamma=0.3, beta=1/0.07; one bench step = {inner} integrator steps",
                   "sample_replicas": sample},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "replica-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


OTHER_CONFIGS = ("lj32k", "lj256k", "dimer256k", "lj1m_langevin", "lj1m_langevin_default_rng", "replicas")
OTHER_CONFIGS_MULTI = ("lj1m_langevin", "dimer256k", "replicas")  # BASELINE.json configs[3], [4], [2] on N > 1 GPUs


def measure_other_configs(args, rank, world):
    """One short device-resident measurement of every other BASELINE.json configuration: all five on one GPU; on
    N > 1 GPUs the three that are defined there (config #4 LJ + LangevinInertia, config #5 dimer in solvent, the
    replica ensemble), each with its `parity` object.  Every rank takes part; rank 0 keeps the rows."""
    import copy
    import gc

    import torch

    rows = []
    for name in (OTHER_CONFIGS if world == 1 else OTHER_CONFIGS_MULTI):
        sub = copy.copy(args)
        sub.steps, sub.warmup, sub.quick = min(args.steps, 20), 3, world == 1
        cfg = WORKLOADS[name]
        try:
            line = (bench_replicas if name == "replicas" else bench_lj)(sub, cfg, rank, world,
                                                                            extras="parity" if world > 1 else False)
            if line is not None:
                row = {"workload": name, "config": line["config"]["workload"], "value": line["value"],
                       "unit": line["unit"], "ms_per_step": line["ms_per_step"], "steps": sub.steps,
                       "clocks": line["clocks"]}
                if "roofline" in line:
                    row["roofline"] = {k: line["roofline"].get(k) for k in ("bound", "frac", "kernel", "achieved", "peak", "unit")}
                if world > 1:
                    row["parallelism"] = line["config"].get("parallelism")
                    row["parity"] = line.get("parity")
                rows.append(row)
        except Exception as exc:  # noqa: BLE001 - one failing configuration must not cost the headline line
            rows.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})
        gc.collect()
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lj1m", choices=sorted(WORKLOADS))
    ap.add_argument("--skin", type=float, default=None, help="neighbour-list skin (default: the package default)")
    ap.add_argument("--quick", action="store_true", help="tuning runs: skip the e2e and cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        out = bench_reference(args, cfg, rank, int(os.environ.get("WORLD_SIZE", "1")))
    else:
        rank, world, _ = dist_setup(args.gpus)
        with ClockSampler(list(range(world)), enabled=rank == 0) as clocks:
            args.clocks = clocks
            out = (bench_replicas if args.workload == "replicas" else bench_lj)(args, cfg, rank, world)
            if args.workload == "lj1m" and not args.quick and os.environ.get("TMD_BENCH_CONFIGS", "1") == "1":
                rows = measure_other_configs(args, rank, world)  # (every rank runs them; rank 0 holds the line)
                if out is not None:
                    out["configs"] = rows
        if out is not None:
            print(json.dumps(out), flush=True)
            out = None
        if world > 1:
            import torch.distributed as dist

            # the result is out; a communicator that will not shut down must not hang the box
            watchdog = threading.Timer(60.0, lambda: os._exit(0))
            watchdog.daemon = True
            watchdog.start()
            dist.barrier()
            dist.destroy_process_group()
            watchdog.cancel()
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
