"""
Synthetic-input helpers shared by bench.py and the tools (not part of the product API).
"""
import numpy as np


def hill(x, y, width_u, width_v, centre=(0.0, 0.0), azimuth=0.0):
    """
    A smooth unit-height hill: exp(-(u/width_u)^2 - (v/width_v)^2) in axes (u, v) turned by
    ``azimuth`` degrees around ``centre``.  Only used to make the synthetic topography of the
    relief workloads ("sum of 8 seeded bumps").
    """
    phi = np.deg2rad(azimuth)
    du, dv = np.asarray(x) - centre[0], np.asarray(y) - centre[1]
    u = (np.cos(phi) * du + np.sin(phi) * dv) / width_u
    v = (np.cos(phi) * dv - np.sin(phi) * du) / width_v
    return np.exp(-(u * u + v * v))


def noisy(values, sigma, seed):
    """values + zero-mean pseudorandom normal noise of standard deviation sigma (seeded)."""
    noise = np.random.RandomState(seed).normal(0.0, sigma, np.shape(values))
    return np.asarray(values) + (noise - noise.mean())
