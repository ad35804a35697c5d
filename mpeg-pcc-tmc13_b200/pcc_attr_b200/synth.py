"""Synthetic voxelised point clouds of SURVEY.md 8(d) / BASELINE.json configs
(seeded, numpy only).  Shared by bench.py and the tests; no reference or
oracle code in here."""
import numpy as np


def np_morton(xyz):
    """Vectorised Morton code (x->bit2, y->bit1, z->bit0), 21 bits/axis."""
    xyz = np.asarray(xyz, dtype=np.int64)
    out = np.zeros(xyz.shape[0], dtype=np.int64)
    for b in range(21):
        out |= ((xyz[:, 0] >> b) & 1) << (3 * b + 2)
        out |= ((xyz[:, 1] >> b) & 1) << (3 * b + 1)
        out |= ((xyz[:, 2] >> b) & 1) << (3 * b)
    return out


def sort_cloud(xyz, attrs):
    """Stable Morton sort of a cloud; returns (morton, attrs_sorted, order)."""
    keys = np_morton(xyz)
    order = np.argsort(keys, kind="stable")
    return keys[order], np.ascontiguousarray(attrs[order]), order



def _smooth_attr(xyz, rng, a, noise=8, bitdepth=8):
    x = xyz.astype(np.float64)
    span = max(1.0, float(x.max()))
    base = np.stack([
        128 + 90 * np.sin(6.0 * x[:, 0] / span + 0.3) * np.cos(4.0 * x[:, 1] / span),
        128 + 80 * np.cos(5.0 * x[:, 1] / span + 1.1) * np.sin(3.0 * x[:, 2] / span),
        128 + 70 * np.sin(7.0 * (x[:, 0] + x[:, 2]) / span),
    ], axis=1)[:, :a]
    v = base * ((1 << bitdepth) / 256.0) + rng.integers(-noise, noise + 1, size=(xyz.shape[0], a))
    return np.clip(np.rint(v), 0, (1 << bitdepth) - 1).astype(np.int32)


def cloud_cube(n=100000, side=47, offset=8, seed=1, a=3):
    """Config 1: first n voxels (Morton order) of a filled side^3 cube."""
    rng = np.random.default_rng(seed)
    g = np.arange(side, dtype=np.int32) + offset
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    keys = np_morton(xyz)
    xyz = xyz[np.argsort(keys, kind="stable")][:n]
    xyz = xyz[rng.permutation(xyz.shape[0])]
    return np.ascontiguousarray(xyz), _smooth_attr(xyz, rng, a)


def cloud_shell(n=100000, bits=10, seed=3, a=3, dups=False):
    """Sphere-shell surface voxelised to `bits` bits (configs 3-5 shape)."""
    rng = np.random.default_rng(seed)
    m = int(n * (1.6 if not dups else 1.0))
    v = rng.normal(size=(m, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = (1 << bits) * 0.45 * (1 + 0.1 * np.sin(5 * v[:, 0]) * np.cos(3 * v[:, 1]))
    xyz = np.clip(np.rint(v * r[:, None] + (1 << bits) / 2), 0, (1 << bits) - 1).astype(np.int32)
    if not dups:
        xyz = np.unique(xyz, axis=0)
        xyz = xyz[rng.permutation(xyz.shape[0])]
    xyz = xyz[:n]
    return np.ascontiguousarray(xyz), _smooth_attr(xyz, rng, a)


def texture(attrs, amplitude, seed, bitdepth=8):
    """Per-point texture on top of a smooth attribute field: uniform noise of
    +-amplitude (the survey's clouds have 12-27 % of their coefficient positions
    quantise to 1 or 2 at the CTC rate points; the smooth field alone leaves
    0.1 % of them non-zero)."""
    rng = np.random.default_rng(seed)
    v = attrs + rng.integers(-amplitude, amplitude + 1, size=attrs.shape)
    return np.clip(v, 0, (1 << bitdepth) - 1).astype(np.int32)


def cloud_lidar(n=1000000, seed=2, a=3, scale=0.25, lasers=64):
    """Config 2: Ford-shaped spinning-LiDAR ring cloud, 1 mm grid scaled by
    `scale` (positionQuantizationScale), duplicates merged."""
    rng = np.random.default_rng(seed)
    az_steps = (n * 115 // 100) // lasers + 1
    theta = np.deg2rad(np.linspace(-24.8, 2.0, lasers))
    az = np.linspace(0, 2 * np.pi, az_steps, endpoint=False)
    T, AZ = np.meshgrid(theta, az, indexing="ij")
    T = T.ravel()
    AZ = AZ.ravel()
    # range to ground plane z = -1.8 m, capped at 80 m
    with np.errstate(divide="ignore"):
        rng_ground = np.where(np.sin(T) < -1e-3, -1.8 / np.sin(T), 80.0)
    r = np.minimum(rng_ground, 80.0)
    # four box obstacles (azimuth sector, distance)
    for a0, a1, d in ((0.3, 0.6, 12.0), (1.8, 2.3, 20.0), (3.5, 3.7, 7.0), (5.0, 5.6, 30.0)):
        hit = (AZ > a0) & (AZ < a1) & (r * np.cos(T) > d)
        r = np.where(hit, d / np.maximum(np.cos(T), 1e-3), r)
    r = r + rng.normal(0, 0.01, size=r.shape)
    x = r * np.cos(T) * np.cos(AZ)
    y = r * np.cos(T) * np.sin(AZ)
    z = r * np.sin(T)
    p = np.stack([x, y, z], axis=1) * 1000.0  # mm
    p -= p.min(axis=0)
    xyz = np.rint(p * scale).astype(np.int32)
    xyz = np.unique(xyz, axis=0)
    xyz = xyz[rng.permutation(xyz.shape[0])][:n]
    return np.ascontiguousarray(xyz), _smooth_attr(xyz, rng, a)


def cloud_terrain(n, seed=7, a=3):
    """Dense voxelised surface with millions of points (configs[2]-[4] sizes):
    a height field over a square grid, one voxel per (x, y), so the points are
    unique by construction (no 30M-point np.unique)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(n)))
    x, y = np.meshgrid(np.arange(side, dtype=np.int32), np.arange(side, dtype=np.int32), indexing="ij")
    x, y = x.ravel()[:n], y.ravel()[:n]
    fx, fy = x / float(side), y / float(side)
    h = (0.25 * np.sin(7.0 * fx + 0.5) * np.cos(5.0 * fy) + 0.12 * np.sin(23.0 * fx * fy + 1.0)
         + 0.05 * np.cos(61.0 * fy)) * side
    z = np.rint(h - h.min()).astype(np.int32)
    xyz = np.stack([x, y, z], axis=1).astype(np.int32)
    xyz = xyz[rng.permutation(n)]
    return np.ascontiguousarray(xyz), _smooth_attr(xyz, rng, a)


def morton_slices(xyz, attrs_list, max_points):
    """Partition a cloud into contiguous Morton ranges of at most max_points
    (octree-aligned slices, tmc3/TMC3.cpp:781-810 level limit): returns the
    arrays reordered slice by slice and the slice offsets."""
    order = np.argsort(np_morton(xyz), kind="stable")
    n = xyz.shape[0]
    k = (n + max_points - 1) // max_points
    offs = np.linspace(0, n, k + 1).astype(np.int64)
    return (np.ascontiguousarray(xyz[order]), [np.ascontiguousarray(a[order]) for a in attrs_list], offs)


def cloud_random(n, bits, seed, a=3, dup_frac=0.0, bitdepth=8):
    """Uniform random voxels (worst case for neighbourhood structure)."""
    rng = np.random.default_rng(seed)
    xyz = rng.integers(0, 1 << bits, size=(n, 3), dtype=np.int32)
    if dup_frac > 0:
        k = int(n * dup_frac)
        src = rng.integers(0, n, size=k)
        dst = rng.integers(0, n, size=k)
        xyz[dst] = xyz[src]
    attrs = rng.integers(0, 1 << bitdepth, size=(n, a), dtype=np.int32)
    return xyz, attrs
