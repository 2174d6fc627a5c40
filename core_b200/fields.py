"""Analytic vertex size fields of the benchmark configurations.

Synthetic inputs only: each returns per-vertex arrays in the layout `Part.set_size_field_*`
takes (h [nv,3], R [nv,9] row-major with the frame vectors in the columns)."""
import numpy as np


def iso_linear(xyz, hbar):
    """config 1: IsotropicFunction h = hbar (1 + 2 x)."""
    return hbar * (1.0 + 2.0 * xyz[:, 0])


def shock_planar(xyz, hbar):
    """config 2: planar shock layer, H = (hbar (0.1 + 2|x-0.5|), hbar, hbar), R = I."""
    nv = len(xyz)
    h = np.empty((nv, 3))
    h[:, 0] = hbar * (0.1 + 2.0 * np.abs(xyz[:, 0] - 0.5))
    h[:, 1] = hbar
    h[:, 2] = hbar
    R = np.zeros((nv, 9))
    R[:, 0] = R[:, 4] = R[:, 8] = 1.0
    return h, R


def shock_rotating(xyz, hbar):
    """config 3: rotating shock layer, R = Rz(theta), theta = (pi/3) y,
    H = (hbar (0.1 + 2|x'-0.5|), hbar, 2 hbar) with x' = p . R[:,0]."""
    nv = len(xyz)
    th = (np.pi / 3.0) * xyz[:, 1]
    c, s = np.cos(th), np.sin(th)
    R = np.zeros((nv, 9))
    R[:, 0] = c
    R[:, 1] = -s
    R[:, 3] = s
    R[:, 4] = c
    R[:, 8] = 1.0
    xp = xyz[:, 0] * c + xyz[:, 1] * s
    h = np.empty((nv, 3))
    h[:, 0] = hbar * (0.1 + 2.0 * np.abs(xp - 0.5))
    h[:, 1] = hbar
    h[:, 2] = 2.0 * hbar
    return h, R


def jitter(xyz, amplitude, seed=12345):
    """Moves interior vertices of the unit box by amplitude*(u-0.5) per component."""
    rng = np.random.default_rng(seed)
    out = xyz.copy()
    lo, hi = xyz.min(axis=0), xyz.max(axis=0)
    interior = np.all((xyz > lo + 1e-12) & (xyz < hi - 1e-12), axis=1)
    out[interior] += amplitude * (rng.random((int(interior.sum()), 3)) - 0.5)
    return out
