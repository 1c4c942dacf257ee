"""ncu launch list of the five passes of tmd_npstream_normals at the size one N = 1M Langevin step needs."""
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from turtlemd_b200 import npstream  # noqa: E402

stream = npstream.NumpyStream(np.random.default_rng(0))
stream.begin()
for _ in range(3):
    stream.normals(6_001_128)
stream.end()
