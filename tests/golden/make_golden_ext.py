"""Generates tests/golden/golden_ext.npz: the EXTENSION interpolation modes (tetrahedral, nearest —
no reference counterpart) on small seeded inputs, frozen so that their definition in
oracle/vf_oracle.c cannot drift silently.  Every vector is what the oracle AND the independent numpy
formulation of tests/test_interpolation_ext.py agree on.
Run from the repo root:  python tests/golden/make_golden_ext.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import test_interpolation_ext as T  # noqa: E402
from gst_plugins_rs_b200 import frames  # noqa: E402

W, H = 96, 24
LUTS = {"lut5dom": frames.cube_text_3d(5, domain_min=(0.0, 0.1, 0.0), domain_max=(1.0, 0.9, 2.0)),
        "lut17": frames.cube_text_3d(17)}


def main():
    oracle.build()
    out = {}
    px = frames.frame_rand(W, H, 4, 142).reshape(-1)
    px.reshape(-1, 4)[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]   # the grey axis: all ties
    out["in_rgba"] = px
    px64 = frames.random_bytes(W * H * 8, 144)
    out["in_rgba64"] = px64
    p4 = px.reshape(-1, 4)
    p16 = np.frombuffer(px64.tobytes(), "<u2").reshape(-1, 4)
    for lname, text in LUTS.items():
        lut = oracle.Lut(text=text)
        tet = oracle.colorlut(lut, px, W, H, interpolation="tetrahedral").reshape(-1, 4)
        want, _ = T._np_tetrahedral(lut, p4[:, :3], 255)
        assert np.array_equal(tet[:, :3], want) and np.array_equal(tet[:, 3], p4[:, 3]), lname
        near = oracle.colorlut(lut, px, W, H, interpolation="nearest").reshape(-1, 4)
        assert np.array_equal(near[:, :3], T._np_nearest(lut, p4[:, :3], 255)), lname
        out[f"tetrahedral_RGBA_{lname}"] = tet.reshape(-1)
        out[f"nearest_RGBA_{lname}"] = near.reshape(-1)
        tet16 = oracle.colorlut(lut, px64, W, H, "RGBA64_LE", interpolation="tetrahedral")
        got16 = np.frombuffer(tet16.tobytes(), "<u2").reshape(-1, 4)
        want16, _ = T._np_tetrahedral(lut, p16[:, :3], 65535)
        assert np.array_equal(got16[:, :3], want16) and np.array_equal(got16[:, 3], p16[:, 3]), lname
        out[f"tetrahedral_RGBA64_LE_{lname}"] = tet16
        out[f"nearest_RGBA64_BE_{lname}"] = oracle.colorlut(lut, px64, W, H, "RGBA64_BE",
                                                          interpolation="nearest")
    np.savez_compressed(os.path.join(HERE, "golden_ext.npz"), **out)
    print("wrote golden_ext.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
