"""Round 2: what the device continuation of NumPy's generator (tmd_npstream_*) costs and buys.

    python profiles/r02ba_npstream.py > gpurun_out/r02ba_npstream.json

One JSON object per line:
  * generator: tmd_npstream_normals for the 6 000 000 normals one LangevinInertia step of the N = 1M fluid consumes
    (CUDA events, 20 calls) beside rgen.standard_normal on the host + the upload it replaces;
  * lj1m_langevin under LangevinInertia(seed=0) -- the reference's default_rng(0) -- with the stream on the device,
    with host draws (the round-2 default until now), and under the in-kernel Philox stream (the bench configuration);
  * 1M DoubleWell replicas under LangevinInertia(seed=0): device stream against host draws.
"""
from __future__ import annotations

import json
import os
import pathlib
import sys
import time

import numpy as np

REPO = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import torch  # noqa: E402

import bench  # noqa: E402
from turtlemd_b200 import npstream  # noqa: E402
from turtlemd_b200.integrators import LangevinInertia  # noqa: E402
from turtlemd_b200.potentials import DoubleWell  # noqa: E402
from turtlemd_b200.system import Box, Particles, System  # noqa: E402


def device_ms(fn, reps):
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(reps):
        fn()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / reps


def main():
    count = 6_001_128  # 2 * 3 * 1 000 188
    rgen = np.random.default_rng(0)
    twin = np.random.default_rng(0)
    stream = npstream.NumpyStream(rgen)
    stream.begin()
    first = stream.normals(count).cpu().numpy()
    exact = bool(np.array_equal(first.view(np.uint64), twin.standard_normal(count).view(np.uint64)))
    for _ in range(3):
        stream.normals(count)
    ms = device_ms(lambda: stream.normals(count), 20)
    stream.end()
    t0 = time.perf_counter()
    host = twin.standard_normal(count)
    t_host = (time.perf_counter() - t0) * 1e3
    pinned = torch.from_numpy(host).pin_memory()
    dev = torch.empty(count, dtype=torch.float64, device="cuda")
    t_up = device_ms(lambda: dev.copy_(pinned, non_blocking=True), 5)
    words, rewalks, off_chain = stream.stats()
    print(json.dumps({"what": "generator", "normals_per_call": count, "bit_exact_vs_numpy": exact,
                      "device_ms_per_call": ms, "device_normals_per_s": count / ms * 1e3,
                      "out_bytes_per_s": 8 * count / ms * 1e3, "host_numpy_ms": t_host, "h2d_pinned_ms": t_up,
                      "words_per_normal": words / (24.0 * count), "rewalks": rewalks, "chunks_placed_off_chain": off_chain}), flush=True)

    if os.environ.get("TMD_PROF_ONLY_GENERATOR") == "1":
        return
    cfg = bench.WORKLOADS["lj1m_langevin"]
    pos, vel, size, n, describe, make = bench.build_lj(cfg)
    rows = {}
    skip_host = os.environ.get("TMD_PROF_SKIP_HOST") == "1"  # (measured once: 435 ms per step)
    for label, steps in (("device_stream", 20), ("philox", 20), ("host_draw", 3)):
        if skip_host and label == "host_draw":
            continue
        system, integ, pot = make()
        if label != "philox":
            integ = LangevinInertia(timestep=cfg["dt"], gamma=0.3, beta=1.0 / 0.8, seed=0)
        npstream.ENABLED = label != "host_draw"
        system.potential_and_force()
        integ.run(system, 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms_step = device_ms(lambda: integ.run(system, steps), 1) / steps
        wall = (time.perf_counter() - t0) * 1e3 / steps
        npstream.ENABLED = True
        rows[label] = {"ms_per_step_device": ms_step, "ms_per_step_wall": wall, "atom_steps_per_s": n / wall * 1e3}
        del system, integ, pot
    rows["what"] = "lj1m_langevin, LangevinInertia(seed=0) = default_rng(0) [device_stream, host_draw] vs PhiloxNormal"
    print(json.dumps(rows), flush=True)

    nrep = 1_000_000
    reps = {}
    for label, steps in (("device_stream", 20), ("host_draw", 3)):
        if skip_host and label == "host_draw":
            continue
        x0 = np.full((nrep, 1), -1.0)
        system = System(Box(periodic=[False]), Particles.from_arrays(x0), [DoubleWell(a=1.0, b=2.0, c=0.0)])
        system.potential_and_force()
        integ = LangevinInertia(timestep=0.002, gamma=0.3, beta=1.0 / 0.07, seed=0)
        npstream.ENABLED = label != "host_draw"
        integ.run(system, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        integ.run(system, steps)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / steps
        npstream.ENABLED = True
        reps[label] = {"ms_per_step_wall": wall, "replica_steps_per_s": nrep / wall * 1e3}
    reps["what"] = "1M DoubleWell replicas, LangevinInertia(seed=0) = default_rng(0)"
    print(json.dumps(reps), flush=True)


if __name__ == "__main__":
    main()
