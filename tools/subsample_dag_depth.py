"""Developer analysis (CPU): depth of the dependency DAG of the distance
subsampling, per level of detail.  A cell waits for the decisions of those of
its 19 neighbour cells (tmc3/PCCTMC3Common.h:2010-2050) that precede it in
Morton order and share its atlas; depth = longest chain of such waits.  With
the kernel times of a launch list this gives the time per hop.

  subsample_dag_depth.py [n] [bits]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from pcc_attr_b200.synth import cloud_shell  # noqa: E402
import pcc_testlib as tl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 11
xyz, _ = cloud_shell(n, bits=bits, seed=40)
lp = tl.make_lod_params(levels=12)
preds, idx, npl = tl.oracle_lod_build(lp, xyz)  # numPointsInLod, coarse -> fine (cumulative)
print(f"# {n} points, {bits}-bit shell; cumulative points per LoD (coarse -> fine): {list(npl)}")

# the input of level l is what the levels before it retained: replay with the
# oracle's own retained sets (indexes are in coding order, coarse first)
order = np.asarray(idx, dtype=np.int64)  # point index by coding position
npl = list(npl)


def depth_of_level(points, shift0):
    """points: positions entering the level; cells of edge 2^(shift0+1)"""
    c = points >> (shift0 + 1)
    # Morton order of cells == lexicographic order of interleaved bits; use the oracle's morton
    mort = np.array([tl.load_oracle().oracle_morton_addr(int(a), int(b), int(d)) for a, b, d in
                     np.unique(c, axis=0)], dtype=np.int64)
    cells = np.unique(c, axis=0)
    o = np.argsort(mort, kind="stable")
    cells = cells[o]
    index = {tuple(v): i for i, v in enumerate(cells.tolist())}
    level = np.zeros(len(cells), dtype=np.int32)
    offs = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)
            if (dx, dy, dz) != (0, 0, 0) and min(dx, dy, dz) < 0]  # the 19 that can precede in Morton order
    for i, v in enumerate(cells.tolist()):
        m = 0
        for dx, dy, dz in offs:
            q = index.get((v[0] + dx, v[1] + dy, v[2] + dz))
            if q is not None and q < i and level[q] > m:
                m = level[q]
        level[i] = m + 1
    return len(cells), int(level.max())


total = 0
cur = np.arange(xyz.shape[0])
# level l (fine first): the input is all points not yet refined; replay using the
# retained counts of the oracle: the points of the coarser LoDs are the retained ones
L = len(npl)
for l in range(L - 1):
    retained_count = npl[L - 2 - l]           # points that survive level l
    inp = order[:npl[L - 1 - l]]              # points entering level l (coding order, coarse first)
    ncells, depth = depth_of_level(xyz[inp], l)
    total += depth
    print(f"level {l:2d}: {len(inp):8d} points in {ncells:8d} cells, depth {depth:6d}")
    if len(inp) < 64:
        break
print(f"# sum of depths over the levels: {total}")
