"""Deterministic synthetic elevation maps: gradient-noise fBm evaluated at the
grid_map cell positions, plus NaN holes, cliffs and exactly-flat plateaus.  numpy only."""
from __future__ import annotations

import numpy as np


def cell_positions(n, resolution, position=0.0):
    """grid_map::getPositionFromIndex operand order."""
    length = n * resolution
    offset = 0.5 * length - 0.5 * resolution
    return (position + offset) + resolution * (-np.arange(n, dtype=np.float64))


def _perlin(x, y, perm, grads):
    xi = np.floor(x).astype(np.int64)
    yi = np.floor(y).astype(np.int64)
    xf = x - xi
    yf = y - yi
    u = xf * xf * xf * (xf * (xf * 6 - 15) + 10)
    v = yf * yf * yf * (yf * (yf * 6 - 15) + 10)

    def g(ix, iy, dx, dy):
        h = perm[(perm[ix & 255] + iy) & 255]
        gr = grads[h & 7]
        return gr[..., 0] * dx + gr[..., 1] * dy

    n00 = g(xi, yi, xf, yf)
    n10 = g(xi + 1, yi, xf - 1, yf)
    n01 = g(xi, yi + 1, xf, yf - 1)
    n11 = g(xi + 1, yi + 1, xf - 1, yf - 1)
    return (n00 * (1 - u) + n10 * u) * (1 - v) + (n01 * (1 - u) + n11 * u) * v


def fbm(rows, cols, resolution, seed, amplitude=0.15, octaves=5, wavelength=2.0, position=(0.0, 0.0)):
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.permutation(256)
    ang = np.arange(8) * (np.pi / 4)
    grads = np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    X = cell_positions(rows, resolution, position[0])[:, None]
    Y = cell_positions(cols, resolution, position[1])[None, :]
    z = np.zeros((rows, cols), dtype=np.float64)
    amp, freq, tot = 1.0, 1.0 / wavelength, 0.0
    for o in range(octaves):
        z += amp * _perlin(X * freq + 17.3 * o, Y * freq - 5.1 * o, perm, grads)
        tot += amp
        amp *= 0.5
        freq *= 2.0
    return (amplitude * z / tot * 2.0).astype(np.float32)


def add_features(z, seed, hole_fraction=0.01, cliffs=4, flats=2, white_noise=0.0, stripe=True, lone_valid=True):
    """NaN blobs (~8 cells wide), raised rectangles (cliffs), exactly flat plateaus, optional NaN stripe."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    rows, cols = z.shape
    z = z.copy()
    if white_noise > 0:
        z += (white_noise * rng.standard_normal(z.shape)).astype(np.float32)
    for _ in range(cliffs):
        h, w = rng.integers(max(4, rows // 16), max(6, rows // 6)), rng.integers(max(4, cols // 16), max(6, cols // 6))
        i0, j0 = rng.integers(0, rows - h), rng.integers(0, cols - w)
        z[i0:i0 + h, j0:j0 + w] += np.float32(0.2)
    for _ in range(flats):
        h, w = rng.integers(max(4, rows // 20), max(6, rows // 8)), rng.integers(max(4, cols // 20), max(6, cols // 8))
        i0, j0 = rng.integers(0, rows - h), rng.integers(0, cols - w)
        z[i0:i0 + h, j0:j0 + w] = np.float32(0.125)
    if hole_fraction > 0:
        nblobs = max(1, int(hole_fraction * rows * cols / 50.0))
        ci = rng.integers(0, rows, nblobs)
        cj = rng.integers(0, cols, nblobs)
        ii, jj = np.arange(rows)[:, None], np.arange(cols)[None, :]
        mask = np.zeros(z.shape, dtype=bool)
        for a, b in zip(ci, cj):
            i0, i1, j0, j1 = max(0, a - 5), min(rows, a + 6), max(0, b - 5), min(cols, b + 6)
            sub = (ii[i0:i1] - a) ** 2 + (jj[:, j0:j1] - b) ** 2 <= 16
            mask[i0:i1, j0:j1] |= sub
            if lone_valid and (a + b) % 3 == 0:
                mask[a, b] = False  # isolated valid cell inside a hole (nPoints == 1 roughness case)
        z[mask] = np.nan
    if stripe and cols > 40:
        j0 = cols // 3
        z[: rows // 4, j0:j0 + 3] = np.nan  # NaN stripe touching the map edge
    return z


def terrain(rows, cols, resolution=0.02, seed=1, preset="mixed", position=(0.0, 0.0), holes=0.01):
    if preset == "gentle":
        return fbm(rows, cols, resolution, seed, 0.15, position=position)
    if preset == "rough":
        z = fbm(rows, cols, resolution, seed, 0.6, position=position)
        return add_features(z, seed, 0.0, 0, 0, white_noise=0.02, stripe=False)
    z = fbm(rows, cols, resolution, seed, 0.15, position=position)
    return add_features(z, seed, holes)
