"""Committed golden vectors (tests/golden/golden_small.npz, made by tests/golden/make_golden.py):
the oracle must still reproduce them (CPU), and so must the CUDA path through the C ABI (GPU)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_small.npz"))
W, H = mg.W, mg.H


def _cases():
    for name, s in mg.FILTER.items():
        yield ("hsvfilter", "RGBA", None, s, "in_rgba", f"hsvfilter_RGBA_{name}")
        yield ("hsvfilter", "xBGR", None, s, "in_rgba", f"hsvfilter_xBGR_{name}")
    yield ("hsvfilter", "BGR", None, mg.FILTER["cfg2"], "in_rgb", "hsvfilter_BGR_cfg2")
    for name, s in mg.DETECT.items():
        yield ("hsvdetector", "BGRx", "RGBA", s, "in_rgba", f"hsvdetector_BGRx_RGBA_{name}")
        yield ("hsvdetector", "RGB", "ABGR", s, "in_rgb", f"hsvdetector_RGB_ABGR_{name}")
    yield ("colorlut", "RGBA", mg.LUT3, None, "in_rgba", "colorlut_RGBA_lut5dom")
    yield ("colorlut", "RGBA", mg.LUT33, None, "in_rgba", "colorlut_RGBA_lut33")
    yield ("colorlut", "RGBA", mg.LUT1, None, "in_rgba", "colorlut_RGBA_lut1d17")
    yield ("colorlut", "RGBA64_LE", mg.LUT33, None, "in_rgba64", "colorlut_RGBA64_LE_lut33")
    yield ("colorlut", "RGBA64_BE", mg.LUT33, None, "in_rgba64", "colorlut_RGBA64_BE_lut33")


CASES = list(_cases())
IDS = [c[5] for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_reproduces_golden(orc, case):
    elem, fmt, extra, settings, kin, kout = case
    src = G[kin]
    if elem == "hsvfilter":
        got = orc.hsvfilter(src, W, H, fmt, settings)
    elif elem == "hsvdetector":
        got = orc.hsvdetector(src, W, H, fmt, extra, settings)
    else:
        got = orc.colorlut(orc.Lut(text=extra), src, W, H, fmt)
    assert np.array_equal(got, G[kout])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_gpu_reproduces_golden(ctx, vf, case):
    import util
    elem, fmt, extra, settings, kin, kout = case
    src = G[kin]
    for memory in ("device", "host"):
        if elem == "hsvfilter":
            got = util.gpu_hsvfilter(ctx, src, W, H, fmt, settings, memory=memory)
        elif elem == "hsvdetector":
            got = util.gpu_hsvdetector(ctx, src, W, H, fmt, extra, settings, memory=memory)
        else:
            ctx.set_lut_from_cube(vf.parse_cube(extra))
            got = util.gpu_colorlut(ctx, src, W, H, fmt, memory=memory)
        assert np.array_equal(got, G[kout]), memory


# ---- extension interpolation modes: frozen definition --------------------------
GX = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ext.npz"))


def _ext_cases():
    import make_golden_ext as mx
    for lname, text in mx.LUTS.items():
        yield (text, "RGBA", "tetrahedral", "in_rgba", f"tetrahedral_RGBA_{lname}")
        yield (text, "RGBA", "nearest", "in_rgba", f"nearest_RGBA_{lname}")
        yield (text, "RGBA64_LE", "tetrahedral", "in_rgba64", f"tetrahedral_RGBA64_LE_{lname}")
        yield (text, "RGBA64_BE", "nearest", "in_rgba64", f"nearest_RGBA64_BE_{lname}")


EXT_CASES = list(_ext_cases())


@pytest.mark.parametrize("case", EXT_CASES, ids=[c[4] for c in EXT_CASES])
def test_oracle_reproduces_extension_golden(orc, case):
    text, fmt, mode, kin, kout = case
    got = orc.colorlut(orc.Lut(text=text), GX[kin], W, H, fmt, interpolation=mode)
    assert np.array_equal(got, GX[kout])
