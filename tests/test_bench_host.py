"""CPU tests of the host-side logic around the hot path: synthetic workloads,
slice partitioning, the bench's bookkeeping helpers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))


def test_terrain_cloud_and_slices():
    from pcc_attr_b200.synth import cloud_terrain, morton_slices, np_morton

    xyz, rgb = cloud_terrain(50000, seed=3)
    assert xyz.shape == (50000, 3) and rgb.shape == (50000, 3)
    assert len(np.unique(xyz, axis=0)) == 50000              # one voxel per point
    refl = rgb[:, :1].copy()
    sx, (srgb, srefl), offs = morton_slices(xyz, [rgb, refl], 12000)
    assert offs[0] == 0 and offs[-1] == 50000 and np.all(np.diff(offs) > 0)
    assert np.max(np.diff(offs)) <= 12000                    # level limit per slice
    keys = np_morton(sx)
    assert np.all(np.diff(keys) >= 0)                        # slices are contiguous Morton ranges
    # the attributes travelled with their points
    order = np.argsort(np_morton(xyz), kind="stable")
    assert np.array_equal(sx, xyz[order]) and np.array_equal(srgb, rgb[order])
    assert np.array_equal(srefl, refl[order])


def test_texture_is_seeded_and_clipped():
    from pcc_attr_b200.synth import texture

    a = np.full((1000, 3), 250, dtype=np.int32)
    t1, t2 = texture(a, 16, 5), texture(a, 16, 5)
    assert np.array_equal(t1, t2) and t1.max() <= 255 and t1.min() >= 234
    assert not np.array_equal(texture(a, 16, 6), t1)


def test_bench_bookkeeping():
    import bench

    cfg_a, cfg_b = bench.workload_config(16), bench.workload_config(16)
    assert cfg_a == cfg_b and "attribute_model" in cfg_a     # both arms print the same config
    assert bench.physical_cores() >= 1
    c = np.array([[0, 1, -2, 5, 0], [0, 0, 0, 0, 0], [0, 1, 0, -1, 0]], dtype=np.int32)
    h = bench.coefficient_histogram([("x", c)])["x"]
    assert h == {"zero": 0.4, "soft": 0.4, "hard": 0.2}
    seeds = [set(bench.frame_seeds(r, 16)) for r in range(8)]
    assert all(not (seeds[i] & seeds[j]) for i in range(8) for j in range(i))


def test_extra_workloads_are_described():
    import bench_workloads as bw

    for name, (npts, per_slice, desc, alg) in bw.WORKLOADS.items():
        assert per_slice <= 1_100_000 and npts >= 3 * per_slice - per_slice and alg > 0 and "configs[" in desc
