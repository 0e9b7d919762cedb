"""The invalid-domain error text and the float grammar must not depend on C-library formatting or
on the process locale.

Expected strings below are written BY HAND from the reference text, not taken from the oracle:
`format!("Invalid domain min {domain_min:?}, max {domain_max:?}")` (parser.rs:209-211) wrapped by
`CubeParseError::InvalidLut` → "Invalid LUT: …" (parser.rs:76-95).  `{:?}` of `[f32; 3]` prints each
element with f32's Debug: shortest round-trip digits, always a fractional part (`1.0`), exponential
form below 1e-4 and from 1e16 (`1e-7`, `1e16`), `NaN`, `inf`, `-inf`.
"""
import os
import shutil
import subprocess
import sys

import pytest

import oracle
from gst_plugins_rs_b200.api import B200VFError, parse_cube

BODY = "LUT_1D_SIZE 2\n0 0 0\n1 1 1\n"

CASES = [
    # DOMAIN_MIN tokens, DOMAIN_MAX tokens, expected message (hand-written)
    ("1 1 1", "1 1 1", "Invalid LUT: Invalid domain min [1.0, 1.0, 1.0], max [1.0, 1.0, 1.0]"),
    ("0 0 0", "0 0 0", "Invalid LUT: Invalid domain min [0.0, 0.0, 0.0], max [0.0, 0.0, 0.0]"),
    ("0.5 0 0", "0.25 1 1", "Invalid LUT: Invalid domain min [0.5, 0.0, 0.0], max [0.25, 1.0, 1.0]"),
    ("1e-7 0 0", "1e-7 1 1", "Invalid LUT: Invalid domain min [1e-7, 0.0, 0.0], max [1e-7, 1.0, 1.0]"),
    ("0.0001 0 0", "0.0001 1 1",
     "Invalid LUT: Invalid domain min [0.0001, 0.0, 0.0], max [0.0001, 1.0, 1.0]"),
    ("0.00009 0 0", "0.00009 1 1",
     "Invalid LUT: Invalid domain min [9e-5, 0.0, 0.0], max [9e-5, 1.0, 1.0]"),
    ("inf 0 0", "inf 1 1", "Invalid LUT: Invalid domain min [inf, 0.0, 0.0], max [inf, 1.0, 1.0]"),
    ("0 0 0", "1 -inf 1", "Invalid LUT: Invalid domain min [0.0, 0.0, 0.0], max [1.0, -inf, 1.0]"),
    ("1e16 0 0", "1e16 1 1", "Invalid LUT: Invalid domain min [1e16, 0.0, 0.0], max [1e16, 1.0, 1.0]"),
    ("1e15 0 0", "1e15 1 1",
     "Invalid LUT: Invalid domain min [1000000000000000.0, 0.0, 0.0], max [1000000000000000.0, 1.0, 1.0]"),
    ("-0 0 0", "-0.0 1 1", "Invalid LUT: Invalid domain min [-0.0, 0.0, 0.0], max [-0.0, 1.0, 1.0]"),
    ("0.1 2.5 100", "0.1 2.5 100",
     "Invalid LUT: Invalid domain min [0.1, 2.5, 100.0], max [0.1, 2.5, 100.0]"),
    ("16777217 0 0", "3 1 1",  # 2^24 + 1 rounds to 16777216 in f32
     "Invalid LUT: Invalid domain min [16777216.0, 0.0, 0.0], max [3.0, 1.0, 1.0]"),
    ("1e39 0 0", "3.4028235e38 1 1",  # 1e39 overflows to inf in dec2flt; f32::MAX prints in e-form
     "Invalid LUT: Invalid domain min [inf, 0.0, 0.0], max [3.4028235e38, 1.0, 1.0]"),
    ("0.1 0 0", "0.1 1 1", "Invalid LUT: Invalid domain min [0.1, 0.0, 0.0], max [0.1, 1.0, 1.0]"),
    ("1.5e-5 0 0", "1e-46 1 1",  # 1e-46 underflows to 0
     "Invalid LUT: Invalid domain min [1.5e-5, 0.0, 0.0], max [0.0, 1.0, 1.0]"),
    ("123456.79 0 0", "2 1 1", "Invalid LUT: Invalid domain min [123456.79, 0.0, 0.0], max [2.0, 1.0, 1.0]"),
]


def _text(dmin, dmax):
    return f"DOMAIN_MIN {dmin}\nDOMAIN_MAX {dmax}\n" + BODY


@pytest.mark.parametrize("dmin,dmax,expected", CASES)
def test_domain_message_product(dmin, dmax, expected):
    with pytest.raises(B200VFError) as e:
        parse_cube(_text(dmin, dmax))
    assert str(e.value).endswith(expected), str(e.value)


@pytest.mark.parametrize("dmin,dmax,expected", CASES)
def test_domain_message_oracle(dmin, dmax, expected):
    with pytest.raises(oracle.CubeError) as e:
        oracle.Lut(text=_text(dmin, dmax))
    assert str(e.value) == expected


def test_nan_domain_is_accepted():
    """NaN bounds pass `min >= max` (every comparison with NaN is false), as in Rust."""
    cube = parse_cube(_text("nan 0 0", "1 1 1"))
    assert cube["size"] == 2


CHILD = r"""
import ctypes, locale, sys
sys.path.insert(0, %r)
locale.setlocale(locale.LC_NUMERIC, "xx_XX")
assert locale.localeconv()["decimal_point"] == ","
libc = ctypes.CDLL(None)
libc.strtod.restype = ctypes.c_double
assert libc.strtod(b"0.5", None) == 0.0   # the hazard: plain strtof/strtod now stop at the '.'
import numpy as np
import oracle
from gst_plugins_rs_b200.api import parse_cube
text = "LUT_1D_SIZE 2\nDOMAIN_MIN 0.25 0.25 0.25\nDOMAIN_MAX 1.5 1.5 1.5\n0.5 0.25 0.125\n1.0 0.75 0.625\n"
c = parse_cube(text)
assert list(c["data"]) == [0.5, 1.0, 0.25, 0.75, 0.125, 0.625], list(c["data"])
assert abs(float(c["scale"][0]) - 0.8) < 1e-6, c["scale"]
o = oracle.Lut(text=text)
assert np.array_equal(o.data.reshape(-1)[:6], np.asarray(c["data"], np.float32)), o.data
bad = "DOMAIN_MIN 1.5 0 0\nDOMAIN_MAX 1.5 1 1\nLUT_1D_SIZE 2\n0 0 0\n1 1 1\n"
want = "Invalid LUT: Invalid domain min [1.5, 0.0, 0.0], max [1.5, 1.0, 1.0]"
for parse in (parse_cube, lambda t: oracle.Lut(text=t)):
    try:
        parse(bad)
    except Exception as e:
        assert str(e).endswith(want), str(e)
    else:
        raise AssertionError("no error")
print("OK")
"""

CHARMAP = "<code_set_name> ASCII-MIN\n<mb_cur_min> 1\n<mb_cur_max> 1\nCHARMAP\n<U0000>..<U007F> /x00\nEND CHARMAP\n"
LOCALE_SRC = 'LC_NUMERIC\ndecimal_point ","\nthousands_sep "."\ngrouping 3;3\nEND LC_NUMERIC\n'


def test_parser_ignores_comma_decimal_locale(tmp_path):
    """Rust's f32::from_str never looks at LC_NUMERIC, while gst_init / GTK apps call
    setlocale(LC_ALL, "") and plain strtof would then read "0.5" as 0.  The image ships no
    comma-decimal locale, so one is compiled with localedef (LC_NUMERIC only) and selected through
    LOCPATH in a child process; the child first proves that libc's strtod IS affected."""
    if not shutil.which("localedef"):
        pytest.skip("localedef not available")
    (tmp_path / "charmap").write_text(CHARMAP)
    (tmp_path / "xx_XX.src").write_text(LOCALE_SRC)
    out = tmp_path / "out"
    out.mkdir()
    subprocess.run(["localedef", "-c", "-f", str(tmp_path / "charmap"), "-i", str(tmp_path / "xx_XX.src"),
                    str(out / "xx_XX")], capture_output=True)
    if not (out / "xx_XX" / "LC_NUMERIC").exists():
        pytest.skip("localedef could not build the test locale")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, LOCPATH=str(out))
    r = subprocess.run([sys.executable, "-c", CHILD % root], capture_output=True, text=True, cwd=root, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().endswith("OK"), r.stdout
