"""Synthetic clouds (TEST / BENCH INPUT GENERATORS).

* bunny_like : points on a bumpy closed surface with the extent of data/bun000.ply (~0.15 m)
* lidar_sweep: 64-beam spinning LiDAR ray-cast against a ground plane, boxes and walls
               (C3: N=100 000 seed 2024; C5: N=1 000 000 seed 2025)
"""
import numpy as np


def bunny_like(n, seed=0):
    rng = np.random.default_rng(seed)
    u = rng.uniform(0, 2 * np.pi, n)
    v = np.arccos(rng.uniform(-1, 1, n))
    r = 0.06 * (1 + 0.25 * np.sin(3 * u) * np.sin(2 * v) + 0.15 * np.cos(5 * v))
    p = np.stack([r * np.sin(v) * np.cos(u) - 0.017, 1.25 * r * np.cos(v) + 0.11, 0.8 * r * np.sin(v) * np.sin(u)], 1)
    p += rng.normal(0, 2e-4, p.shape)
    return p.astype(np.float32)


def lidar_sweep(n, seed=2024, n_beams=64):
    """elevations linspace(-25,+3 deg), uniform azimuth, ground z=-1.8 m, 40 boxes, 4 walls at +-70 m,
    range noise N(0, 0.02 m), max range 75 m; misses dropped, first n kept."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-60, 60, (40, 2))
    sizes = rng.uniform(1, 8, (40, 2))
    heights = rng.uniform(1.5, 4, 40)
    lo = np.c_[centres - sizes / 2, np.full(40, -1.8)]
    hi = np.c_[centres + sizes / 2, -1.8 + heights]
    out = []
    have = 0
    elev = np.deg2rad(np.linspace(-25.0, 3.0, n_beams))
    while have < n:
        m = max(4 * (n - have), 65536)
        az = rng.uniform(0, 2 * np.pi, m)
        el = elev[rng.integers(0, n_beams, m)]
        d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], 1)
        t = np.full(m, np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = -1.8 / d[:, 2]
            t = np.where((d[:, 2] < 0) & (tg > 0), np.minimum(t, tg), t)
            for ax in (0, 1):
                for s in (-70.0, 70.0):
                    tw = s / d[:, ax]
                    t = np.where(tw > 0, np.minimum(t, tw), t)
            inv = 1.0 / d
            for b in range(40):
                t1 = lo[b] * inv
                t2 = hi[b] * inv
                tn = np.nanmax(np.minimum(t1, t2), axis=1)
                tf = np.nanmin(np.maximum(t1, t2), axis=1)
                hit = (tn <= tf) & (tn > 0)
                t = np.where(hit, np.minimum(t, tn), t)
        t = t + rng.normal(0, 0.02, m)
        ok = np.isfinite(t) & (t > 0.5) & (t < 75.0)
        out.append((d[ok] * t[ok, None]))
        have += int(ok.sum())
    return np.concatenate(out)[:n].astype(np.float32)
