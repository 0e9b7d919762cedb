"""Differential fuzzing of the two .cube parsers (product C++ vs oracle C): for thousands of
mutated files both must agree on accept/reject, on the error text, and on every parsed float."""
import random

import numpy as np

from gst_plugins_rs_b200.api import B200VFError, parse_cube

TOKENS = ["0", "1", "0.5", "-0.25", "1e-3", "1E+2", ".5", "5.", "+1", "-0", "inf", "-inf", "nan", "NaN",
          "Infinity", "1e", "e5", ".", "0x10", "1_0", "1,5", "abc", "TITLE", "DOMAIN_MIN", "DOMAIN_MAX",
          "LUT_1D_SIZE", "LUT_3D_SIZE", "LUT_3D_INPUT_RANGE", "#", "\"x\"", "2", "3", "4", "65537", "256",
          "257", "-2", "+2", "2.0", "1e400", "1e-400", "0.30000001192092896", "340282356779733661637539395458142568448"]
SPACES = [" ", "  ", "\t", " ", "　", " ", "\x0b", "\x0c"]
EOLS = ["\n", "\r\n", "\n\n", "\r\n\r\n"]


def base_text(rng):
    kind = rng.choice(["1d", "3d"])
    n = rng.choice([2, 2, 3])
    lines = []
    if rng.random() < 0.3:
        lines.append("# comment")
    if rng.random() < 0.3:
        lines.append('TITLE "fuzz"')
    lines.append(f"LUT_{'1' if kind == '1d' else '3'}D_SIZE {n}")
    if rng.random() < 0.4:
        lines.append("DOMAIN_MIN 0 0 0")
    if rng.random() < 0.4:
        lines.append("DOMAIN_MAX 1 2 4")
    count = n if kind == "1d" else n ** 3
    for _ in range(count):
        lines.append(" ".join("%.4f" % rng.random() for _ in range(3)))
    return lines


def mutate(lines, rng):
    lines = list(lines)
    for _ in range(rng.randint(0, 3)):
        op = rng.randint(0, 7)
        i = rng.randrange(len(lines)) if lines else 0
        if op == 0 and lines:
            del lines[i]
        elif op == 1:
            lines.insert(i, " ".join(rng.choice(TOKENS) for _ in range(rng.randint(1, 4))))
        elif op == 2 and lines:
            toks = lines[i].split()
            if toks:
                toks[rng.randrange(len(toks))] = rng.choice(TOKENS)
            lines[i] = " ".join(toks)
        elif op == 3 and lines:
            lines[i] = lines[i] + " " + rng.choice(TOKENS)
        elif op == 4 and lines:
            lines[i] = rng.choice(SPACES) + lines[i].replace(" ", rng.choice(SPACES)) + rng.choice(SPACES)
        elif op == 5 and lines:
            j = rng.randrange(len(lines))
            lines[i], lines[j] = lines[j], lines[i]
        elif op == 6 and lines:
            lines.insert(i, lines[i])
        elif op == 7:
            lines.insert(i, rng.choice(["", "   ", "#x", " "]))
    eol = rng.choice(EOLS)
    text = eol.join(lines)
    if rng.random() < 0.7:
        text += eol
    return text


def test_parsers_agree_on_mutated_files(orc):
    rng = random.Random(20261017)
    accepted = rejected = 0
    for case in range(4000):
        text = mutate(base_text(rng), rng)
        try:
            lut = orc.Lut(text=text)
            want = ("ok", lut.kind, lut.size, lut.data.tobytes(), lut.scale.tobytes(), lut.offset.tobytes())
        except orc.CubeError as e:
            want = ("err", e.code, str(e))
        try:
            c = parse_cube(text)
            got = ("ok", c["kind"], c["size"], c["data"].tobytes(), np.asarray(c["scale"]).tobytes(),
                   np.asarray(c["offset"]).tobytes())
        except B200VFError as e:
            got = ("err", {-5: 1, -6: 2}[e.status], e.message)
        assert got == want, f"case {case}: {text!r}\nproduct {got[:3]}\noracle  {want[:3]}"
        accepted += want[0] == "ok"
        rejected += want[0] == "err"
    assert accepted > 300 and rejected > 300, (accepted, rejected)


def test_parsers_agree_on_raw_bytes(orc):
    """Arbitrary bytes (invalid UTF-8 included): same accept / Io / InvalidLut classification."""
    rng = random.Random(7)
    seeds = [b"LUT_1D_SIZE 2\n0 0 0\n1 1 1\n", b"LUT_3D_SIZE 2\n" + b"0.5 0.25 1\n" * 8]
    for case in range(1500):
        b = bytearray(rng.choice(seeds))
        for _ in range(rng.randint(1, 4)):
            i = rng.randrange(len(b))
            r = rng.random()
            if r < 0.4:
                b[i] = rng.randrange(256)
            elif r < 0.7:
                b.insert(i, rng.choice([0xC2, 0xA0, 0xE3, 0x80, 0xFF, 0x20, 0x0A, 0x31, 0x2E, 0x65]))
            else:
                del b[i]
        raw = bytes(b)
        try:
            lut = orc.Lut(text=raw)
            want = ("ok", lut.data.tobytes())
        except orc.CubeError as e:
            want = ("err", e.code, str(e))
        try:
            c = parse_cube(raw)
            got = ("ok", c["data"].tobytes())
        except B200VFError as e:
            got = ("err", {-5: 1, -6: 2}[e.status], e.message)
        if want[0] == "err" and got[0] == "err":
            # error texts embed the offending line; compare after the same lossy decoding
            assert got[1] == want[1], f"case {case}: {raw!r}: {got} vs {want}"
        else:
            assert got == want, f"case {case}: {raw!r}: {got[:2]} vs {want[:2]}"


def test_product_parser_under_sanitizers(tmp_path):
    """csrc/vf_cube_parser.cpp compiled with AddressSanitizer + UBSan and fed 20,000 mutated /
    random files (tests/cpp/parser_sanitize.cpp): no memory error, no undefined behaviour, every
    outcome a clean rejection or a well-formed LUT."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "parser_sanitize"
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=all", "-I/usr/local/cuda/include", "-o", str(exe),
                    os.path.join(root, "tests", "cpp", "parser_sanitize.cpp"),
                    os.path.join(root, "gst-plugins-rs_b200", "csrc", "vf_cube_parser.cpp")], check=True)
    out = subprocess.run([str(exe), "20000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "parser_sanitize: ok" in out.stdout
