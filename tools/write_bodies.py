"""Writes a side-loader body file (solaris_b200/host/side_loader.h) for the drop-in program's SOLARIS_B200_BODIES.

    from tools.write_bodies import write_bodies
    write_bodies(path, btype, state[n,6], kind=0|1, mass=None, radius=None, density=None, cD=None)
"""
import struct

import numpy as np


def write_bodies(path, btype, state, kind=0, mass=None, radius=None, density=None, cD=None):
    state = np.ascontiguousarray(state, dtype="<f8").reshape(-1, 6)
    n = len(state)
    assert btype in (6, 7) and kind in (0, 1)
    z = np.zeros(n, dtype="<f8")
    with open(path, "wb") as f:
        f.write(b"SOLB200B")
        f.write(struct.pack("<4i", 1, n, btype, kind))
        f.write(state.tobytes())
        for arr in (mass, radius, density, cD):
            f.write((z if arr is None else np.ascontiguousarray(arr, dtype="<f8")).tobytes())
