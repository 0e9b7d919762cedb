"""Second, independent restatement of the reference arithmetic in numpy float32 (vectorised).

numpy never fuses multiply-add and np.fmod / division on float32 arrays are IEEE single
operations, so this follows the same rounding sequence as the Rust.
It exists to cross-check the C oracle — two restatements written separately (this one array-
at-a-time from the appendix, the C one loop-by-loop from the sources) must agree bit for bit.
Test infrastructure only.
"""
import numpy as np

F = np.float32


def _u8_trunc(v):
    """Rust `as u8` after clamp(0,255): NaN → 0."""
    v = np.where(np.isnan(v), F(0), v)
    return np.clip(v, 0, 255).astype(np.uint8)  # astype truncates toward zero


def from_rgb(r8, g8, b8):
    """A.2 — r8,g8,b8 uint8 arrays → (h,s,v) float32."""
    r, g, b = (x.astype(F) / F(255.0) for x in (r8, g8, b8))
    mx = np.maximum(np.maximum(r8, g8), b8)
    mn = np.minimum(np.minimum(r8, g8), b8)
    value = mx.astype(F) / F(255.0)
    chroma = value - mn.astype(F) / F(255.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        h_r = F(60.0) * ((g - b) / chroma)
        h_g = F(60.0) * (F(2.0) + ((b - r) / chroma))
        h_b = F(60.0) * (F(4.0) + ((r - g) / chroma))
        sat = np.where(value == 0, F(0), chroma / value)
    eps = F(0.00001)
    hue = np.where(chroma == 0, F(0),
                   np.where(np.abs(value - r) < eps, h_r,
                            np.where(np.abs(value - g) < eps, h_g,
                                     np.where(np.abs(value - b) < eps, h_b, F(0)))))
    hue = np.where(hue < 0, hue + F(360.0), hue).astype(F)
    return np.fmod(hue, F(360.0)).astype(F), np.clip(sat, 0, 1).astype(F), np.clip(value, 0, 1)


def to_rgb(h, s, v):
    """A.4 → (r,g,b) uint8."""
    c = (v * s).astype(F)
    hp = (h / F(60.0)).astype(F)
    with np.errstate(invalid="ignore"):
        x = (c * (F(1.0) - np.abs(np.fmod(hp, F(2.0)).astype(F) - F(1.0)))).astype(F)
    z = np.zeros_like(c)
    conds = [hp < 0, hp <= 1, hp <= 2, hp <= 3, hp <= 4, hp <= 5, hp <= 6]
    p0 = np.select(conds, [z, c, x, z, z, x, c], z)
    p1 = np.select(conds, [z, x, c, c, x, z, z], z)
    p2 = np.select(conds, [z, z, z, x, c, c, x], z)
    m = (v - c).astype(F)
    return tuple(_u8_trunc(((p + m).astype(F) * F(255.0)).astype(F)) for p in (p0, p1, p2))


def _trait_clamp01(v):
    """hsvutils.rs:23-37: self.max(0).min(1) with f32::max/min NaN rules (NaN → 0)."""
    return np.fmin(np.fmax(v, F(0)), F(1)).astype(F)


def hsvfilter_rgb(r8, g8, b8, settings):
    """A.2 + A.3 + A.4 on channel arrays."""
    hs, sm, so, vm, vo = (F(x) for x in settings)
    h, s, v = from_rgb(r8, g8, b8)
    with np.errstate(invalid="ignore", over="ignore"):
        h = np.fmod((h + hs).astype(F), F(360.0)).astype(F)
        h = np.where(h < 0, h + F(360.0), h).astype(F)
        s = _trait_clamp01((sm * s).astype(F) + so)
        v = _trait_clamp01((vm * v).astype(F) + vo)
    return to_rgb(h, s, v)


def hsvdetector_mask(r8, g8, b8, settings):
    """A.5 → uint8 alpha (255 / 0)."""
    href, hvar, sref, svar, vref, vvar = (F(x) for x in settings)
    h, s, v = from_rgb(r8, g8, b8)
    with np.errstate(invalid="ignore"):
        sh = (h + (F(180.0) - href)).astype(F)
        sh = np.where(sh < 0, sh + F(360.0), sh).astype(F)
        sh = np.fmod(sh, F(360.0)).astype(F)
        ok = (np.abs(sh - F(180.0)) <= hvar) & (np.abs(s - sref) <= svar) & (np.abs(v - vref) <= vvar)
    return np.where(ok, 255, 0).astype(np.uint8)


def _round_half_away(y):
    return np.where(y >= 0, np.floor(y + 0.5), np.ceil(y - 0.5))


def colorlut_3d(codes, lut, n, scale, offset, maxcode=255):
    """A.1 — codes (P,3) uint8/uint16, lut (n^3,4) float32 → (P,3) codes."""
    mc = F(maxcode)
    sm1 = F(n) - F(1.0)
    idx, t = [], []
    for c in range(3):
        v = codes[:, c].astype(F) / mc
        nrm = np.clip((v * F(scale[c])).astype(F) + F(offset[c]), F(0), F(1)).astype(F)
        p = (nrm * sm1).astype(F)
        i0 = np.minimum(np.where(np.isnan(p), 0, np.floor(p)).astype(np.int64), n - 1)
        idx.append((i0, np.minimum(i0 + 1, n - 1)))
        t.append((p - i0.astype(F)).astype(F))

    def at(xi, yi, zi):
        return lut[xi + yi * n + zi * n * n]

    def lerp(a, b, tt):
        with np.errstate(invalid="ignore", over="ignore"):
            return (a + ((b - a).astype(F) * tt[:, None]).astype(F)).astype(F)

    (x0, x1), (y0, y1), (z0, z1) = idx
    c00 = lerp(at(x0, y0, z0), at(x1, y0, z0), t[0])
    c10 = lerp(at(x0, y1, z0), at(x1, y1, z0), t[0])
    c01 = lerp(at(x0, y0, z1), at(x1, y0, z1), t[0])
    c11 = lerp(at(x0, y1, z1), at(x1, y1, z1), t[0])
    c0, c1 = lerp(c00, c10, t[1]), lerp(c01, c11, t[1])
    o = lerp(c0, c1, t[2])[:, :3]
    with np.errstate(invalid="ignore"):
        y = (np.clip(o, F(0), F(1)).astype(F) * mc).astype(F)
    y = np.where(np.isnan(y), F(0), y)
    return _round_half_away(y.astype(np.float64)).astype(np.uint16 if maxcode > 255 else np.uint8)


def colorlut_1d(codes, planes, n, scale, offset, maxcode=255):
    """A.1 1D variant — planes (3,n) float32."""
    mc = F(maxcode)
    out = []
    for c in range(3):
        v = codes[:, c].astype(F) / mc
        nrm = np.clip((v * F(scale[c])).astype(F) + F(offset[c]), F(0), F(1)).astype(F)
        p = (nrm * (F(n) - F(1.0))).astype(F)
        i0 = np.minimum(np.where(np.isnan(p), 0, np.floor(p)).astype(np.int64), n - 1)
        i1 = np.minimum(i0 + 1, n - 1)
        tt = (p - i0.astype(F)).astype(F)
        a, b = planes[c][i0], planes[c][i1]
        o = (a + ((b - a).astype(F) * tt).astype(F)).astype(F)
        y = (np.clip(o, F(0), F(1)).astype(F) * mc).astype(F)
        y = np.where(np.isnan(y), F(0), y)
        out.append(_round_half_away(y.astype(np.float64)))
    return np.stack(out, 1).astype(np.uint16 if maxcode > 255 else np.uint8)
