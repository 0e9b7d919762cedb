"""Synthetic inputs of SURVEY.md §8(d): frame content classes and .cube files.

Pure data generation (numpy); no pixel transform lives here.
"""
import math

import numpy as np

_M64 = (1 << 64) - 1


def splitmix64(seed, n):
    """n successive SplitMix64 outputs (uint64) for `seed`."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_bytes(n, frame_index=0):
    """C-rand bytes: seed 0x5EED0000 + frame_index, little-endian bytes of successive outputs."""
    words = splitmix64(0x5EED0000 + frame_index, (n + 7) // 8)
    return words.view(np.uint8)[:n].copy()


def frame_rand(width, height, bpp=4, frame_index=0):
    return random_bytes(width * height * bpp, frame_index).reshape(height, width * bpp)


def frame_grad(width, height):
    """C-grad RGBA: r = x*255/(W-1), g = y*255/(H-1), b = (x+y)*255/(W+H-2), a = 255."""
    x = np.arange(width, dtype=np.int64)[None, :]
    y = np.arange(height, dtype=np.int64)[:, None]
    out = np.empty((height, width, 4), np.uint8)
    out[..., 0] = (x * 255 // max(width - 1, 1)).astype(np.uint8)
    out[..., 1] = (y * 255 // max(height - 1, 1)).astype(np.uint8)
    out[..., 2] = ((x + y) * 255 // max(width + height - 2, 1)).astype(np.uint8)
    out[..., 3] = 255
    return out.reshape(height, width * 4)


_BARS = [(191, 191, 191), (191, 191, 0), (0, 191, 191), (0, 191, 0), (191, 0, 191), (191, 0, 0),
         (0, 0, 191)]


def frame_bars(width, height):
    """C-bars RGBA: 7 vertical 75 % bars over the top 2/3, a grey ramp below, a = 255
    (stand-in for videotestsrc's default SMPTE pattern)."""
    out = np.empty((height, width, 4), np.uint8)
    idx = np.minimum(np.arange(width) * 7 // max(width, 1), 6)
    bars = np.array(_BARS, np.uint8)[idx]  # (W,3)
    split = (height * 2) // 3
    out[:split, :, :3] = bars[None, :, :]
    ramp = (np.arange(width) * 255 // max(width - 1, 1)).astype(np.uint8)
    out[split:, :, :3] = ramp[None, :, None]
    out[..., 3] = 255
    return out.reshape(height, width * 4)


def frame_noise(width, height, frame_index=0, amplitude=2):
    """C-noise RGBA: C-grad plus independent uniform noise in [-amplitude, +amplitude] on R, G, B
    (a camera-like stand-in: smooth content, noisy low bits), a = 255."""
    base = frame_grad(width, height).reshape(height, width, 4).astype(np.int16)
    span = 2 * amplitude + 1
    n = (random_bytes(width * height * 3, 7000 + frame_index).reshape(height, width, 3) % span)
    base[..., :3] += n.astype(np.int16) - amplitude
    return np.clip(base, 0, 255).astype(np.uint8).reshape(height, width * 4)


def frame_of_class(content, width, height, frame_index=0):
    if content == "noise":
        return frame_noise(width, height, frame_index)
    if content == "bars":
        return frame_bars(width, height)
    if content == "grad":
        return frame_grad(width, height)
    if content == "rand":
        return frame_rand(width, height, 4, frame_index)
    raise ValueError(content)


def all_rgb_frame(fmt_r=0, fmt_g=1, fmt_b=2, other=3, other_value=77):
    """4096x4096 4-byte frame holding every 8-bit RGB triple once, channels at the given
    byte positions, the remaining byte constant."""
    a = np.arange(1 << 24, dtype=np.uint32)
    px = np.empty((1 << 24, 4), np.uint8)
    px[:, fmt_r] = a & 255
    px[:, fmt_g] = (a >> 8) & 255
    px[:, fmt_b] = (a >> 16) & 255
    px[:, other] = other_value
    return px.reshape(4096, 4096 * 4)


def synthetic_lut_values(n):
    """§8(d) synthetic LUT, (n^3, 3) float64, R fastest."""
    g = np.arange(n, dtype=np.float64) / (n - 1)
    r_, g_, b_ = np.meshgrid(g, g, g, indexing="ij")  # r_[r,g,b]
    # order entries with R fastest: index = r + g*n + b*n^2 → transpose to [b,g,r]
    r = r_.transpose(2, 1, 0).ravel()
    gg = g_.transpose(2, 1, 0).ravel()
    b = b_.transpose(2, 1, 0).ravel()
    o0 = r ** 0.8 * 0.9 + 0.1 * gg
    o1 = 0.5 - 0.45 * np.cos(math.pi * gg) + 0.05 * b
    o2 = b ** 1.2 * 0.85 + 0.15 * r
    return np.clip(np.stack([o0, o1, o2], 1), 0.0, 1.0)


def identity_lut_values(n):
    g = np.arange(n, dtype=np.float64) / (n - 1)
    r_, g_, b_ = np.meshgrid(g, g, g, indexing="ij")
    return np.stack([r_.transpose(2, 1, 0).ravel(), g_.transpose(2, 1, 0).ravel(),
                     b_.transpose(2, 1, 0).ravel()], 1)


def cube_text_3d(n, values=None, domain_min=None, domain_max=None, title=None):
    """Adobe .cube text for an n^3 LUT (values default to the §8(d) synthetic LUT), '%.6f'."""
    if values is None:
        values = synthetic_lut_values(n)
    head = []
    if title:
        head.append(f'TITLE "{title}"')
    head.append(f"LUT_3D_SIZE {n}")
    if domain_min is not None:
        head.append("DOMAIN_MIN %s" % " ".join(repr(float(v)) for v in domain_min))
    if domain_max is not None:
        head.append("DOMAIN_MAX %s" % " ".join(repr(float(v)) for v in domain_max))
    body = "\n".join("%.6f %.6f %.6f" % tuple(v) for v in values)
    return "\n".join(head) + "\n" + body + "\n"


def cube_text_1d(n, values=None, domain_min=None, domain_max=None):
    if values is None:
        x = np.arange(n, dtype=np.float64) / (n - 1)
        values = np.clip(np.stack([x ** 0.8, 0.5 - 0.5 * np.cos(math.pi * x), x ** 1.2], 1), 0, 1)
    head = [f"LUT_1D_SIZE {n}"]
    if domain_min is not None:
        head.append("DOMAIN_MIN %s" % " ".join(repr(float(v)) for v in domain_min))
    if domain_max is not None:
        head.append("DOMAIN_MAX %s" % " ".join(repr(float(v)) for v in domain_max))
    body = "\n".join("%.6f %.6f %.6f" % tuple(v) for v in values)
    return "\n".join(head) + "\n" + body + "\n"
