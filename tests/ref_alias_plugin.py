"""pytest plugin: make `import turtlemd...` resolve to turtlemd_b200 (same module layout)."""
import importlib
import sys

import turtlemd_b200

MODULES = ["", ".integrators", ".potentials", ".potentials.lennardjones", ".potentials.potential", ".potentials.well",
           ".simulation", ".simulation.mdsimulation", ".simulation.stopping", ".system", ".system.box",
           ".system.particles", ".system.system", ".tools", ".tools.tools", ".tools.xyz", ".version"]
for name in MODULES:
    try:
        sys.modules["turtlemd" + name] = importlib.import_module("turtlemd_b200" + name)
    except ImportError as exc:
        print("alias: no module turtlemd_b200" + name, exc)
