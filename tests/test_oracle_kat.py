"""Pin the oracle (and the product .cube parser) against every known-answer test the
reference itself holds for this path:
  parser.rs:377-474  — 5 parser tests
  hsvutils.rs:200-280 — 4 HSV conversion tests
plus the probe facts of SURVEY.md §8c and the oracle self-check invariants.
"""
import math

import numpy as np
import pytest

from gst_plugins_rs_b200 import frames
from gst_plugins_rs_b200.api import B200VFError, ERR_PARSE, parse_cube

# ---- the reference's parser tests, verbatim inputs (parser.rs:381-473) ---------------------
PARSE_3D = """
            LUT_3D_SIZE 2

            0.0 0.0 0.0
            1.0 0.0 0.0
            0.0 1.0 0.0
            1.0 1.0 0.0
            0.0 0.0 1.0
            1.0 0.0 1.0
            0.0 1.0 1.0
            1.0 1.0 1.0
        """
KEYWORD_AFTER_SIZE = """
            LUT_1D_SIZE 2

            TITLE "test"
            DOMAIN_MIN 0.0 0.0 0.0
            DOMAIN_MAX 1.0 1.0 1.0

            0.0 0.0 0.0
            1.0 0.5 0.7
        """
KEYWORD_AFTER_DATA = """
            LUT_1D_SIZE 2

            0.0 0.0 0.0
            1.0 0.0 0.0
            TITLE "invalid"
        """
KEYWORD_BETWEEN_DATA = """
            LUT_1D_SIZE 2

            0.0 0.0 0.0
            TITLE "invalid"
            1.0 0.0 0.0
        """
MULTIPLE_SIZES = """
            LUT_1D_SIZE 2
            LUT_3D_SIZE 2

            0.0 0.0 0.0
            1.0 1.0 1.0
        """


def _parsers(orc):
    """Both parsers with one calling convention → dict or raises."""
    def oracle_parse(text):
        lut = orc.Lut(text=text)
        return {"kind": lut.kind, "size": lut.size, "data": lut.data, "scale": lut.scale,
                "offset": lut.offset}
    return [("oracle", oracle_parse, orc.CubeError), ("product", parse_cube, B200VFError)]


def test_parse_3d_lut(orc):
    """parser.rs:381-408"""
    for name, parse, _ in _parsers(orc):
        c = parse(PARSE_3D)
        assert c["kind"] == 3 and c["size"] == 2, name
        flat = c["data"].reshape(-1, 4)
        assert len(flat) == 8
        assert list(flat[0]) == [0.0, 0.0, 0.0, 1.0]          # at(0,0,0)
        assert list(flat[1 + 1 * 2 + 1 * 4]) == [1.0, 1.0, 1.0, 1.0]  # at(1,1,1)
        assert (flat[:, 3] == 1.0).all()


def test_keyword_after_lut_size(orc):
    """parser.rs:410-434"""
    for name, parse, _ in _parsers(orc):
        c = parse(KEYWORD_AFTER_SIZE)
        assert c["kind"] == 1 and c["size"] == 2, name
        r, g, b = c["data"].reshape(3, 2)
        assert list(r) == [0.0, 1.0]
        assert list(g) == [np.float32(0.0), np.float32(0.5)]
        assert list(b) == [np.float32(0.0), np.float32(0.7)]


@pytest.mark.parametrize("text", [KEYWORD_AFTER_DATA, KEYWORD_BETWEEN_DATA, MULTIPLE_SIZES])
def test_parser_error_cases(orc, text):
    """parser.rs:436-447, 449-460, 462-473"""
    for name, parse, exc in _parsers(orc):
        with pytest.raises(exc):
            parse(text)


# ---- the reference's HSV tests (hsvutils.rs:203-279) ----------------------------------------
EPS = 0.00001
COLOURS = {  # name: (rgb, bgr, hsv)
    "white": ((255, 255, 255), (255, 255, 255), (0.0, 0.0, 1.0)),
    "black": ((0, 0, 0), (0, 0, 0), (0.0, 0.0, 0.0)),
    "red": ((255, 0, 0), (0, 0, 255), (0.0, 1.0, 1.0)),
    "green": ((0, 255, 0), (0, 255, 0), (120.0, 1.0, 1.0)),
    "blue": ((0, 0, 255), (255, 0, 0), (240.0, 1.0, 1.0)),
}


def is_equivalent(hsv, expected, eps):
    """hsvutils.rs:203-217 (hue compared on the circle)"""
    shifted = np.float32(hsv[0]) + (np.float32(180.0) - np.float32(expected[0]))
    if shifted < 0.0:
        shifted += np.float32(360.0)
    shifted = np.float32(math.fmod(shifted, 360.0))
    return (abs(shifted - 180.0) < eps and abs(hsv[1] - expected[1]) < eps and
            abs(hsv[2] - expected[2]) < eps)


@pytest.mark.parametrize("colour", sorted(COLOURS))
def test_from_rgb_from_bgr(orc, colour):
    """hsvutils.rs:237-257"""
    rgb, bgr, hsv = COLOURS[colour]
    assert is_equivalent(orc.from_rgb(rgb), hsv, EPS)
    assert is_equivalent(orc.from_bgr(bgr), hsv, EPS)


@pytest.mark.parametrize("colour", sorted(COLOURS))
def test_to_rgb_to_bgr(orc, colour):
    """hsvutils.rs:259-279 — exact bytes"""
    rgb, bgr, hsv = COLOURS[colour]
    assert tuple(orc.to_rgb(hsv)) == rgb
    assert tuple(orc.to_bgr(hsv)) == bgr


# ---- parser behaviour table, oracle and product must agree -----------
GOOD = [
    "LUT_3D_SIZE 2\n" + "0 0 0\n" * 8,
    "# c\n\nTITLE x y z\nLUT_1D_SIZE 3\nDOMAIN_MIN -1 -1 -1\nDOMAIN_MAX 2 2 2\n0 0 0\n.5 5. 1e-3\n1 1 1",
    "LUT_1D_SIZE +2\r\n0 0 0\r\n1 1 1\r\n",
    "LUT_1D_SIZE 2\n+1 -0.0 1E+2\ninf -INF NaN\n",
    "LUT_1D_SIZE 2\n0\t0 0\n1　1 1\n",               # Unicode White_Space separators
    "LUT_1D_SIZE 2\nDOMAIN_MIN nan 0 0\n0 0 0\n1 1 1\n",        # NaN bound passes `min >= max`
    "LUT_1D_SIZE 2\n1e50 1e-50 0.1\n16777217 0.30000001192092896 3.4028236e38\n",
]
BAD = [
    "",                                                   # Missing LUT size
    "0 0 0\n",                                            # data before size
    "LUT_1D_SIZE 1\n0 0 0\n",                             # size range
    "LUT_1D_SIZE 65537\n",
    "LUT_3D_SIZE 257\n",
    "LUT_3D_SIZE 2 2\n",
    "LUT_3D_SIZE\n",
    "LUT_3D_SIZE -2\n",
    "LUT_3D_SIZE 2.0\n",
    "LUT_3D_SIZE 2\n" + "0 0 0\n" * 7,                    # value count
    "LUT_1D_SIZE 2\n0 0 0\n1 1 1\n1 1 1\n",
    "LUT_1D_SIZE 2\n0 0\n1 1 1\n",                        # too few components
    "LUT_1D_SIZE 2\n0 0 0 0\n1 1 1\n",                    # too many
    "LUT_1D_SIZE 2\n0 0 x\n1 1 1\n",                      # Invalid float
    "LUT_1D_SIZE 2\n0x10 0 0\n1 1 1\n",                   # hex floats are not Rust floats
    "LUT_1D_SIZE 2\n1e 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\n. 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\n1_0 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\nLUT_3D_INPUT_RANGE 0 1\n0 0 0\n1 1 1\n",   # unknown keyword = data line
    "LUT_1D_SIZE 2\nDOMAIN_MIN 0 0\n0 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\nDOMAIN_MIN 1 1 1\nDOMAIN_MAX 1 2 2\n0 0 0\n1 1 1\n",   # min >= max
    "LUT_1D_SIZE 2\n0 0 0\n1 1 1\nDOMAIN_MAX 1 1 1\n",    # header after data
    "lut_1d_size 2\n0 0 0\n1 1 1\n",                      # keywords are case-sensitive
]


@pytest.mark.parametrize("text", GOOD)
def test_parser_accepts_same(orc, text):
    a = None
    for name, parse, _ in _parsers(orc):
        c = parse(text)
        key = (c["kind"], c["size"], c["data"].tobytes(), np.asarray(c["scale"]).tobytes(),
               np.asarray(c["offset"]).tobytes())
        if a is None:
            a = key
        assert key == a, f"{name} parser disagrees"


@pytest.mark.parametrize("text", BAD)
def test_parser_rejects_same(orc, text):
    msgs = []
    for name, parse, exc in _parsers(orc):
        with pytest.raises(exc) as e:
            parse(text)
        msgs.append(str(getattr(e.value, "message", e.value)))
        if name == "product":
            assert e.value.status == ERR_PARSE
    assert msgs[0] == msgs[1], "error texts differ: %r" % (msgs,)


def test_parser_invalid_utf8_is_io_error(orc):
    """fs::read_to_string rejects invalid UTF-8 → CubeParseError::Io (parser.rs:105-108)."""
    from gst_plugins_rs_b200.api import ERR_IO
    raw = b"LUT_1D_SIZE 2\n0 0 0\n1 1 \xff1\n"
    with pytest.raises(orc.CubeError) as e:
        orc.Lut(text=raw)
    assert e.value.code == 2
    with pytest.raises(B200VFError) as e2:
        parse_cube(raw)
    assert e2.value.status == ERR_IO


def test_domain_scale_offset(orc):
    """parser.rs:264-274"""
    lut = orc.Lut(text="LUT_1D_SIZE 2\nDOMAIN_MIN 0.25 -1 0\nDOMAIN_MAX 0.75 1 4\n0 0 0\n1 1 1\n")
    assert list(lut.scale) == [2.0, 0.5, 0.25]
    assert list(lut.offset) == [-0.5, 0.5, -0.0]
    d = orc.Lut(text="LUT_1D_SIZE 2\n0 0 0\n1 1 1\n")
    assert list(d.scale) == [1.0, 1.0, 1.0]
    assert all(v == 0.0 and math.copysign(1, v) < 0 for v in d.offset)  # -0.0


# ---- oracle self-check invariants -------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 17, 33, 64, 65])
def test_identity_lut_reproduces_input(orc, n):
    lut = orc.Lut(text=frames.cube_text_3d(n, frames.identity_lut_values(n)))
    w, h = 4096, 64
    src = frames.frame_rand(w, h, 4, n)
    src.reshape(-1, 4)[:256, :3] = np.arange(256, dtype=np.uint8)[:, None]
    assert np.array_equal(orc.colorlut(lut, src, w, h), src.reshape(-1))


def test_grid_points_return_lut_entries(orc):
    """A 256^3-style property at small scale: with N = 256 every t is 0 (§8c probe v) — use the
    1D analogue and a 3D LUT of size 2 whose 8 grid inputs are 0/255."""
    vals = np.random.default_rng(0).uniform(0, 1, (8, 3))
    lut = orc.Lut(text=frames.cube_text_3d(2, vals))
    px = np.array([[255 * (i & 1), 255 * ((i >> 1) & 1), 255 * ((i >> 2) & 1), i] for i in range(8)],
                  np.uint8)
    out = orc.colorlut(lut, px, 8, 1).reshape(8, 4)
    stored = np.array([[float("%.6f" % v) for v in row] for row in vals], np.float32)
    want = np.floor(np.clip(stored, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(out[:, :3], want) and np.array_equal(out[:, 3], px[:, 3])


def test_hsvfilter_identity_probe_fact(orc):
    """§8c probe (i): identity settings change 11,093,274 of 2^24 triples, all by exactly 1."""
    src = frames.all_rgb_frame()
    out = orc.hsvfilter(src, 4096, 4096, "RGBA", (0.0, 1.0, 0.0, 1.0, 0.0)).reshape(-1, 4)
    d = np.abs(out[:, :3].astype(np.int16) - src.reshape(-1, 4)[:, :3].astype(np.int16))
    assert int((d != 0).any(1).sum()) == 11093274 and int(d.max()) == 1
    assert (out[:, 3] == 77).all()


def test_hsvdetector_copies_colour_and_keeps_padding(orc):
    w, h, stride_in, stride_out = 33, 5, 33 * 4 + 8, 33 * 4 + 12
    src = frames.random_bytes(stride_in * h, 3)
    dst = np.full(stride_out * h, 0xEE, np.uint8)
    out = orc.hsvdetector(src, w, h, "xBGR", "ARGB", (0, 180, 0.5, 0.5, 0.5, 0.5), stride_in,
                          stride_out, dst=dst)
    s = src.reshape(h, stride_in)[:, :w * 4].reshape(h, w, 4)
    o = out.reshape(h, stride_out)
    assert (o[:, w * 4:] == 0xEE).all()
    o = o[:, :w * 4].reshape(h, w, 4)
    assert np.array_equal(o[..., 1], s[..., 3]) and np.array_equal(o[..., 2], s[..., 2])
    assert np.array_equal(o[..., 3], s[..., 1]) and set(np.unique(o[..., 0])) <= {0, 255}


def test_oracle_is_free_of_undefined_behaviour(tmp_path):
    """oracle/vf_oracle.c under ASan + UBSan on extreme settings / LUT values (tests/cpp/oracle_sanitize.c)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "oracle_sanitize"
    subprocess.run(["gcc", "-std=c11", "-O1", "-g", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=all", "-ffp-contract=off", "-pthread", "-o", str(exe),
                    os.path.join(root, "tests", "cpp", "oracle_sanitize.c"),
                    os.path.join(root, "oracle", "vf_oracle.c"), "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-4000:]
    assert "oracle_sanitize: ok" in out.stdout
