"""The colour-blocked order of the 2^24-entry function tables (vf_ops.cuh blk_index),
checked on the CPU through b200vf_debug_table_indices — the host build of the function the kernels
use: it is a bijection of [0, 2^24), ignores the alpha byte, and every 128-byte line (32 entries)
holds a 4 x 4 x 2 block of neighbouring colours, every 32-byte sector a 4 x 1 x 2 sub-block."""
import ctypes as C

import numpy as np

import gst_plugins_rs_b200 as g


def indices(colours):
    lib = g._lib.load()
    colours = np.ascontiguousarray(colours, np.uint32)
    out = np.empty_like(colours)
    rc = lib.b200vf_debug_table_indices(colours.ctypes.data_as(C.POINTER(C.c_uint32)), colours.size,
                                        out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    return out


def test_index_is_a_bijection_and_ignores_alpha():
    col = np.arange(1 << 24, dtype=np.uint32)
    idx = indices(col)
    assert idx.max() == (1 << 24) - 1
    seen = np.zeros(1 << 24, np.bool_)
    seen[idx] = True
    assert seen.all()
    sample = col[:: 4099]
    assert np.array_equal(indices(sample | np.uint32(0xA5000000)), indices(sample))


def test_lines_are_4x4x2_colour_blocks():
    col = np.arange(1 << 24, dtype=np.uint32)
    idx = indices(col)
    c0, c1, c2 = col & 255, (col >> 8) & 255, col >> 16
    block = (c0 >> 2) | (c1 >> 2) << 6 | (c2 >> 1) << 12       # which 4x4x2 block a colour is in
    line = idx >> 5
    # colours of one block share a line, colours of different blocks never do
    order = np.argsort(line, kind="stable")
    l, b = line[order].reshape(-1, 32), block[order].reshape(-1, 32)
    assert (l == l[:, :1]).all() and (b == b[:, :1]).all()
    assert np.unique(b[:, 0]).size == (1 << 24) // 32
    # a sector (8 entries) = 4 values of c0 x 2 values of c2 at one c1
    sector = idx >> 3
    sub = (c0 >> 2) | c1 << 6 | (c2 >> 1) << 14
    order = np.argsort(sector, kind="stable")
    s, u = sector[order].reshape(-1, 8), sub[order].reshape(-1, 8)
    assert (s == s[:, :1]).all() and (u == u[:, :1]).all()


def test_noise_neighbourhood_spans_few_lines():
    """What the layout is for: the +-2-code neighbourhood of a colour (125 colours) lies in ~12 lines
    here against 25-30 lines of a natural [c2][c1][c0] table."""
    rng = np.random.default_rng(1)
    centres = rng.integers(2, 254, size=(200, 3))
    d = np.arange(-2, 3)
    dd = np.stack(np.meshgrid(d, d, d, indexing="ij"), -1).reshape(-1, 3)
    blocked, natural = [], []
    for c in centres:
        n = c + dd
        col = (n[:, 0] | n[:, 1] << 8 | n[:, 2] << 16).astype(np.uint32)
        blocked.append(np.unique(indices(col) >> 5).size)
        natural.append(np.unique(col >> 5).size)
    assert np.mean(blocked) < 14 and np.mean(natural) > 25
