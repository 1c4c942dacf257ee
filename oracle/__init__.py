"""CPU oracle for the TurtleMD hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference
(infretis/turtlemd 2023.3.1) for the one path the CUDA library replaces:
pair-potential force/energy/virial evaluation and the per-particle
integrator updates.  Every function cites the reference ``file:line`` it
follows.

It exists to CHECK the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``turtlemd_b200/``
imports it, and the product path never falls back to it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every
function here against (i) the known-answer vectors in the reference's own
tests and (ii) fixtures under ``tests/golden/`` produced by importing the
reference itself (``tests/golden/make_golden.py``).
"""
