// lod_subsample_warp.cuh — distance subsampling (subsampleByDistance,
// tmc3/PCCTMC3Common.h:1984-2085) as a warp-cooperative dataflow over cells
// (device only).  Same decisions as SubsampleDistanceFn in lod_core.cuh (the
// host-testable definition):
//   * k_cell_neighbours: one thread per (cell, neighbour offset) resolves the
//     19 neighbour cells by binary search — geometry only, fully parallel;
//   * k_subsample_cells: one warp per cell in Morton ticket order; 19 lanes
//     poll the decision words of the earlier neighbour cells and fetch their
//     retained points, then the cell's points are tested one after the other,
//     all neighbours at once (ballot).
#pragma once

#include "lod_core.cuh"

namespace pccb200 {

struct SubsampleCellsArgs {
  Voxels v;
  const uint32_t* input;
  const int32_t* cellFirst;
  int nCells;
  int shiftBits0;
  int* decision;
  uint8_t* keep;
  int32_t* nb;  // nCells * 19 neighbour cell indices (or -1)
};

__global__ void __launch_bounds__(256)
k_cell_neighbours(const SubsampleCellsArgs a)
{
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (tid >= int64_t(a.nCells) * 19)
    return;
  const int c = int(tid / 19);
  const int n = int(tid - int64_t(c) * 19) + 1;
  const uint8_t kOff[20] = {7,  3,  5,  6,  12, 10, 17, 20, 34, 33,
                            4,  2,  1,  24, 40, 48, 32, 16, 8,  0};
  const int sb3 = 3 * (a.shiftBits0 + 1);
  const int atlasBit = sb3 + 21 < 63 ? sb3 + 21 : 63;
  const int64_t code0 = a.v.code[a.input[a.cellFirst[c]]];
  const int64_t cell = code0 >> sb3;
  const int64_t atlasId = code0 >> atlasBit;
  const uint64_t base = morton3d_add(uint64_t(cell), ~uint64_t(0));
  const int64_t nc = int64_t(morton3d_add(base, kOff[n]));
  int q = -1;
  if ((nc >> 21) == atlasId) {
    q = find_cell(a.v.code, a.input, a.cellFirst, a.nCells, sb3, nc);
    if (q >= c || (q >= 0 && (a.v.code[a.input[a.cellFirst[q]]] >> atlasBit) != atlasId))
      q = -1;
  }
  a.nb[tid] = q;
}

__global__ void __launch_bounds__(256)
k_subsample_cells(const SubsampleCellsArgs a, unsigned long long* ticket)
{
  const int lane = threadIdx.x & 31;
  const int64_t radius2 = int64_t(3) << (a.shiftBits0 << 1);
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0)
      t = atomicAdd(ticket, 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= (unsigned long long)a.nCells)
      return;
    const int c = int(t);
    // retained point of this lane's neighbour cell (if any)
    bool have = false;
    int32_t np[3] = {0, 0, 0};
    if (lane < 19) {
      const int q = a.nb[size_t(c) * 19 + lane];
      if (q >= 0) {
        int d;
        while ((d = ld_acquire(&a.decision[q])) == kCellUndecided)
          __nanosleep(32);
        if (d >= 0) {
          const int32_t* p = &a.v.pos[size_t(a.input[d]) * 3];
          np[0] = p[0];
          np[1] = p[1];
          np[2] = p[2];
          have = true;
        }
      }
    }
    const int i0 = a.cellFirst[c], i1 = a.cellFirst[c + 1];
    int chosen = kCellNone;
    for (int i = i0; i < i1; i++) {
      const int32_t* p = &a.v.pos[size_t(a.input[i]) * 3];
      const int32_t pp[3] = {p[0], p[1], p[2]};
      const bool hit = have && norm2_3(np, pp) <= radius2;
      const bool found = __ballot_sync(0xffffffffu, hit) != 0;
      if (lane == 0)
        a.keep[i] = found ? 0 : 1;
      if (!found) {
        chosen = i;
        for (int r = i + 1 + lane; r < i1; r += 32)
          a.keep[r] = 0;
        break;
      }
    }
    if (lane == 0)
      st_release(&a.decision[c], chosen);
  }
}

}  // namespace pccb200
