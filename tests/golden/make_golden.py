"""Generates tests/golden/golden_small.npz: small seeded inputs and the outputs the two
independent CPU restatements (oracle/ C port and tests/np_emulation.py) AGREE on.

The reference itself cannot be executed here (Rust, no toolchain), so these vectors are not
reference outputs; they freeze the agreed restatement so that (a) the oracle cannot drift
silently and (b) the CUDA path is checked against fixed files on the GPU box.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import np_emulation as npe  # noqa: E402
import oracle  # noqa: E402
from gst_plugins_rs_b200 import frames  # noqa: E402

W, H = 96, 24
FILTER = {"identity": (0.0, 1.0, 0.0, 1.0, 0.0), "cfg2": (37.5, 1.2, 0.05, 0.9, 0.02),
          "wide": (-725.5, 0.6, 0.3, 1.4, -0.2)}
DETECT = {"default": (0.0, 10.0, 0.0, 0.15, 0.0, 0.3), "cfg4": (120.0, 30.0, 0.6, 0.4, 0.6, 0.4)}
LUT3 = frames.cube_text_3d(5, domain_min=(0.0, 0.1, 0.0), domain_max=(1.0, 0.9, 2.0))
LUT33 = frames.cube_text_3d(33)
LUT1 = frames.cube_text_1d(17)


def main():
    oracle.build()
    out = {}
    px = frames.frame_rand(W, H, 4, 42).reshape(-1)
    px.reshape(-1, 4)[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]
    out["in_rgba"] = px
    p4 = px.reshape(-1, 4)
    for name, s in FILTER.items():
        got = oracle.hsvfilter(px, W, H, "RGBA", s).reshape(-1, 4)
        r, g, b = npe.hsvfilter_rgb(p4[:, 0], p4[:, 1], p4[:, 2], s)
        assert np.array_equal(got[:, 0], r) and np.array_equal(got[:, 1], g) and \
            np.array_equal(got[:, 2], b), name
        out[f"hsvfilter_RGBA_{name}"] = got.reshape(-1)
        out[f"hsvfilter_xBGR_{name}"] = oracle.hsvfilter(px, W, H, "xBGR", s)
    rgb3 = frames.random_bytes(W * H * 3, 43)
    out["in_rgb"] = rgb3
    out["hsvfilter_BGR_cfg2"] = oracle.hsvfilter(rgb3, W, H, "BGR", FILTER["cfg2"])
    for name, s in DETECT.items():
        got = oracle.hsvdetector(px, W, H, "BGRx", "RGBA", s).reshape(-1, 4)
        assert np.array_equal(got[:, 3], npe.hsvdetector_mask(p4[:, 2], p4[:, 1], p4[:, 0], s))
        out[f"hsvdetector_BGRx_RGBA_{name}"] = got.reshape(-1)
        out[f"hsvdetector_RGB_ABGR_{name}"] = oracle.hsvdetector(rgb3, W, H, "RGB", "ABGR", s)
    for lname, text in (("lut5dom", LUT3), ("lut33", LUT33)):
        lut = oracle.Lut(text=text)
        got = oracle.colorlut(lut, px, W, H).reshape(-1, 4)
        want = npe.colorlut_3d(p4[:, :3], lut.data.reshape(-1, 4), lut.size, lut.scale, lut.offset)
        assert np.array_equal(got[:, :3], want), lname
        out[f"colorlut_RGBA_{lname}"] = got.reshape(-1)
    px64 = frames.random_bytes(W * H * 8, 44)
    out["in_rgba64"] = px64
    lut = oracle.Lut(text=LUT33)
    out["colorlut_RGBA64_LE_lut33"] = oracle.colorlut(lut, px64, W, H, "RGBA64_LE")
    out["colorlut_RGBA64_BE_lut33"] = oracle.colorlut(lut, px64, W, H, "RGBA64_BE")
    l1 = oracle.Lut(text=LUT1)
    got = oracle.colorlut(l1, px, W, H).reshape(-1, 4)
    assert np.array_equal(got[:, :3], npe.colorlut_1d(p4[:, :3], l1.data.reshape(3, 17), 17,
                                                      l1.scale, l1.offset))
    out["colorlut_RGBA_lut1d17"] = got.reshape(-1)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("wrote golden_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
