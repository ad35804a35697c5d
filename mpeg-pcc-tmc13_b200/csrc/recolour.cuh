// recolour.cuh — attribute transfer from a source cloud to the (re)coded
// geometry: recolourColour / recolourReflectance of the reference
// (tmc3/pointset_processing.cpp:253-923, driver :925-958, call site
// tmc3/encoder.cpp:1031-1037).  The per-item bodies are host/device functors
// (the CPU unit tests run them through tests/emu's HostExec); the schedule is
// written against the executor concept of raht_pipeline.cuh.
//
// The reference searches its neighbours with nanoflann kd-trees
// (dependencies/nanoflann; the one place in the attribute path that does).
// Here both searches are exact k-nearest-neighbour queries over a grid hash:
// the indexed points are sorted by the Morton code of their cell, a cell's
// points are one contiguous range found by binary search, and a query walks
// the cells around it ring by ring until no unvisited cell can hold a closer
// point.  Distances are the reference's (double, squared Euclidean, summed in
// axis order, no fused multiply-add); ties are broken by the lower point
// index, which nanoflann does not promise -- see DESIGN.md for what that
// means for parity (bit-exact against the oracle, which restates the same
// rule; within a stated tolerance of the compiled reference on clouds where
// distance ties reach the k-th neighbour).
#pragma once

#include <math.h>

#include "raht_core.cuh"

namespace pccb200 {

constexpr int kRecolourMaxK = 16;      // neighbours per query (reference default: 8 / 1)
constexpr int kRecolourMaxRing = 24;   // rings walked before a query scans everything

struct RecolourConfig {
  double distOffsetFwd, distOffsetBwd;
  double maxGeomFwd, maxGeomBwd;   // already mapped: >= 512 -> DBL_MAX
  double maxAttrFwd, maxAttrBwd;
  int searchRange;
  int kFwd, kBwd;
  int weightedFwd, weightedBwd;
  int skipFwd, skipBwd;
  int A;             // 1 or 3
  double clipMax;    // (1 << bitdepth) - 1
  double scale;      // sourceToTargetScaleFactor
  double invScale;   // 1.0 / scale
  int off[3];        // targetToSourceOffset
  int nSrc, nTgt;
};

// exact products and sums (the compiler must not contract them into FMAs:
// the reference is built without)
PCC_HD double
dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
PCC_HD double
dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
PCC_HD double
dsub(double a, double b)
{
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}

PCC_HD int
atomic_fetch_add_i32(int* p, int v)
{
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  int old = *p;
  *p += v;
  return old;
#endif
}

//----------------------------------------------------------------------------
// the grid hash over one point set

struct PointGrid {
  const int64_t* code;   // Morton code of the point's cell, ascending
  const int32_t* order;  // original index of sorted entry i
  const int32_t* spos;   // positions in sorted order (n x 3)
  int n;
  int shift;             // cell edge = 1 << shift
  const int32_t* bbox;   // [6]: min / max cell coordinate per axis
};

// cell coordinates of every point (input of the Morton sort)
struct CellCoordFn {
  const int32_t* xyz;
  int shift;
  int32_t* cell;   // n x 3
  int* flag;       // bit 0: a coordinate is negative or too large for a 63-bit Morton code
  PCC_HD void operator()(int64_t i) const
  {
    for (int k = 0; k < 3; k++) {
      const int32_t v = xyz[3 * i + k];
      if (v < 0 || v >= (1 << 21))
        atomic_or_i32(flag, 1);
      cell[3 * i + k] = v < 0 ? 0 : v >> shift;
    }
  }
};

struct GatherPosFn {
  const int32_t* xyz;
  const int32_t* order;
  int shift;
  int32_t* spos;
  int32_t* bbox;   // [6], initialised to INT_MAX x3, INT_MIN x3
  PCC_HD void operator()(int64_t i) const
  {
    const int32_t o = order[i];
    for (int k = 0; k < 3; k++) {
      const int32_t v = xyz[3 * size_t(o) + k];
      spos[3 * i + k] = v;
#if defined(__CUDA_ARCH__)
      // one atomic per warp and bound instead of one per point
      const unsigned act = __activemask();
      const int lo = __reduce_min_sync(act, v >> shift);
      const int hi = __reduce_max_sync(act, v >> shift);
      if ((threadIdx.x & 31) == __ffs(act) - 1) {
        atomicMin(&bbox[k], lo);
        atomicMax(&bbox[3 + k], hi);
      }
#else
      atomic_min_i32(&bbox[k], v >> shift);
      atomic_max_i32(&bbox[3 + k], v >> shift);
#endif
    }
  }
};

// number of distinct cells (adjacent-difference count over the sorted codes)
struct DistinctCodeFn {
  const int64_t* code;
  int* count;
  PCC_HD void operator()(int64_t i) const
  {
    const bool head = i == 0 || code[i] != code[i - 1];
#if defined(__CUDA_ARCH__)
    const unsigned act = __activemask();
    const unsigned heads = __ballot_sync(act, head);
    if (heads && (threadIdx.x & 31) == __ffs(act) - 1)
      atomicAdd(count, __popc(heads));
#else
    if (head)
      atomic_add_i32(count, 1);
#endif
  }
};

// the (dist, index)-ordered result list of one query
struct KnnList {
  double d[kRecolourMaxK];
  int32_t id[kRecolourMaxK];
  int cnt;
  int k;
  PCC_HD void insert(double dist, int32_t idx)
  {
    if (cnt == k) {
      const double w = d[k - 1];
      if (!(dist < w || (dist == w && idx < id[k - 1])))
        return;
    }
    int pos = cnt < k ? cnt : k - 1;
    while (pos > 0 && (d[pos - 1] > dist || (d[pos - 1] == dist && id[pos - 1] > idx))) {
      d[pos] = d[pos - 1];
      id[pos] = id[pos - 1];
      pos--;
    }
    d[pos] = dist;
    id[pos] = idx;
    if (cnt < k)
      cnt++;
  }
};

// squared distance as nanoflann's L2_Simple_Adaptor accumulates it
PCC_HD double
sqr_dist3(const double q[3], const int32_t* p)
{
  double r = 0.0;
  for (int k = 0; k < 3; k++) {
    const double diff = dsub(q[k], double(p[k]));
    r = dadd(r, dmul(diff, diff));
  }
  return r;
}

PCC_HD void
grid_knn(const PointGrid& g, const double q[3], KnnList& L)
{
  L.cnt = 0;
  int64_t qc[3];
  for (int k = 0; k < 3; k++)
    qc[k] = int64_t(floor(q[k])) >> g.shift;
  // the first ring that can touch the occupied box, the last one that has to
  int64_t r0 = 0, r1 = 0;
  for (int k = 0; k < 3; k++) {
    const int64_t lo = g.bbox[k], hi = g.bbox[3 + k];
    const int64_t below = lo - qc[k], above = qc[k] - hi;
    const int64_t gap = below > 0 ? below : above > 0 ? above : 0;
    r0 = gap > r0 ? gap : r0;
    const int64_t a = qc[k] - lo, b = hi - qc[k];
    const int64_t far = (a > b ? a : b);
    r1 = far > r1 ? far : r1;
  }
  const double cs = double(int64_t(1) << g.shift);
  bool scanAll = false;
  for (int64_t r = r0; r <= r1; r++) {
    if (r - r0 > kRecolourMaxRing) {
      scanAll = true;
      break;
    }
    for (int64_t dz = -r; dz <= r; dz++) {
      const int64_t cz = qc[2] + dz;
      if (cz < g.bbox[2] || cz > g.bbox[5])
        continue;
      for (int64_t dy = -r; dy <= r; dy++) {
        const int64_t cy = qc[1] + dy;
        if (cy < g.bbox[1] || cy > g.bbox[4])
          continue;
        const bool shell = dz == -r || dz == r || dy == -r || dy == r;
        // inside the shell's faces only the two end cells of the row belong to ring r
        for (int64_t dx = -r; dx <= r; dx += (shell || r == 0) ? 1 : 2 * r) {
          const int64_t cx = qc[0] + dx;
          if (cx < g.bbox[0] || cx > g.bbox[3])
            continue;
          const int64_t cell = morton_addr(int32_t(cx), int32_t(cy), int32_t(cz));
          int a = 0, b = g.n;
          while (a < b) {
            const int m = (a + b) >> 1;
            if (g.code[m] < cell)
              a = m + 1;
            else
              b = m;
          }
          for (int i = a; i < g.n && g.code[i] == cell; i++)
            L.insert(sqr_dist3(q, &g.spos[3 * size_t(i)]), g.order[i]);
        }
      }
    }
    // every unvisited point is farther than r cells along some axis
    if (L.cnt == L.k) {
      const double bound = dmul(double(r) * cs, double(r) * cs);
      if (L.d[L.k - 1] <= bound)
        return;
    }
  }
  if (scanAll) {
    L.cnt = 0;
    for (int i = 0; i < g.n; i++)
      L.insert(sqr_dist3(q, &g.spos[3 * size_t(i)]), g.order[i]);
  }
}

// one query per item: the target points in the source (forward,
// pointset_processing.cpp:306-313) or the source points in the target
// (backward, :409-418)
struct KnnQueryFn {
  PointGrid g;
  RecolourConfig cfg;
  const int32_t* qxyz;
  int backward;
  int k;
  double* outDist;    // nQueries x k
  int32_t* outIdx;
  PCC_HD void operator()(int64_t i) const
  {
    double q[3];
    for (int c = 0; c < 3; c++) {
      if (backward)  // posInTgt = source * scale - offset
        q[c] = dsub(dmul(double(qxyz[3 * i + c]), cfg.scale), double(cfg.off[c]));
      else  // posInSrc = (target + offset) * (1 / scale)
        q[c] = dmul(double(qxyz[3 * i + c] + cfg.off[c]), cfg.invScale);
    }
    KnnList L;
    L.k = k;
    grid_knn(g, q, L);
    for (int j = 0; j < k; j++) {
      outDist[size_t(i) * k + j] = j < L.cnt ? L.d[j] : 0.0;
      outIdx[size_t(i) * k + j] = j < L.cnt ? L.id[j] : -1;
    }
  }
};

// The reference pops its result vectors when the k-th neighbour is farther
// than maxGeometryDist2Fwd -- and never restores them (the vectors live
// outside the loop, pointset_processing.cpp:301-326): from the first such
// target on, every target sees one neighbour.  firstBad = that target.
struct FirstBadFn {
  const double* dist;
  int k;
  double maxGeom;
  int32_t* firstBad;
  PCC_HD void operator()(int64_t i) const
  {
    if (k > 1 && dist[size_t(i) * k + (k - 1)] > maxGeom)
      atomic_min_i32(firstBad, int32_t(i));
  }
};

PCC_HD double
clip_round(double v, double hi)
{
  const double r = round(v);
  return r < 0.0 ? 0.0 : r > hi ? hi : r;
}

// forward colour of every target (pointset_processing.cpp:301-399 / :660-744)
struct ForwardColourFn {
  RecolourConfig cfg;
  const double* dist;     // nTgt x kFwd
  const int32_t* idx;
  const int32_t* srcAttr; // nSrc x A
  const int32_t* firstBad;
  int32_t* refined1;      // nTgt x A
  PCC_HD void operator()(int64_t t) const
  {
    const int k = cfg.kFwd, A = cfg.A;
    const double* d = dist + size_t(t) * k;
    const int32_t* id = idx + size_t(t) * k;
    int nNN = t >= *firstBad ? 1 : k;
    if (cfg.skipFwd && d[0] < 0.0001)
      nNN = 1;
    while (nNN > 1) {
      double maxAttr = 2.2250738585072014e-308;  // std::numeric_limits<double>::min()
      for (int i = 0; i < nNN; i++)
        for (int j = 0; j < nNN; j++) {
          double s = 0.0;
          for (int c = 0; c < A; c++) {
            const double df = double(srcAttr[size_t(id[i]) * A + c])
              - double(srcAttr[size_t(id[j]) * A + c]);
            s = dadd(s, dmul(df, df));
          }
          if (s > maxAttr)
            maxAttr = s;
        }
      if (maxAttr > cfg.maxAttrFwd) {
        --nNN;
        continue;
      }
      double acc[3] = {0.0, 0.0, 0.0};
      if (cfg.weightedFwd) {
        double sumW = 0.0;
        for (int i = 0; i < nNN; i++) {
          const double w = 1 / dadd(d[i], cfg.distOffsetFwd);
          for (int c = 0; c < A; c++)
            acc[c] = dadd(acc[c], dmul(double(srcAttr[size_t(id[i]) * A + c]), w));
          sumW = dadd(sumW, w);
        }
        for (int c = 0; c < A; c++)
          acc[c] = acc[c] / sumW;
      } else {
        for (int i = 0; i < nNN; i++)
          for (int c = 0; c < A; c++)
            acc[c] = dadd(acc[c], double(srcAttr[size_t(id[i]) * A + c]));
        for (int c = 0; c < A; c++)
          acc[c] = acc[c] / double(nNN);
      }
      for (int c = 0; c < A; c++)
        refined1[size_t(t) * A + c] = int32_t(clip_round(acc[c], cfg.clipMax));
      return;
    }
    for (int c = 0; c < A; c++)
      refined1[size_t(t) * A + c] = srcAttr[size_t(id[0]) * A + c];
  }
};

// backward lists: the sources that name a target among their kBwd nearest
// (pointset_processing.cpp:409-428), as CSR: count, (scan), fill, sort.
struct BackwardCountFn {
  RecolourConfig cfg;
  const double* dist;   // nSrc x kBwd
  const int32_t* idx;
  int* count;           // nTgt (+1)
  PCC_HD void operator()(int64_t s) const
  {
    for (int j = 0; j < cfg.kBwd; j++) {
      const int32_t t = idx[size_t(s) * cfg.kBwd + j];
      if (t >= 0 && dist[size_t(s) * cfg.kBwd + j] <= cfg.maxGeomBwd)
        atomic_add_i32(&count[t], 1);
    }
  }
};

struct BackwardFillFn {
  RecolourConfig cfg;
  const double* dist;
  const int32_t* idx;
  const int* first;     // nTgt + 1 (exclusive scan of the counts)
  int* cursor;          // nTgt, zeroed
  double* listDist;
  int32_t* listSrc;
  PCC_HD void operator()(int64_t s) const
  {
    for (int j = 0; j < cfg.kBwd; j++) {
      const int32_t t = idx[size_t(s) * cfg.kBwd + j];
      const double d = dist[size_t(s) * cfg.kBwd + j];
      if (t >= 0 && d <= cfg.maxGeomBwd) {
        const int at = first[t] + atomic_fetch_add_i32(&cursor[t], 1);
        listDist[at] = d;
        listSrc[at] = int32_t(s);
      }
    }
  }
};

// final colour of every target (pointset_processing.cpp:430-611 / :773-921)
struct FinalColourFn {
  RecolourConfig cfg;
  const int32_t* refined1;
  const int32_t* srcAttr;
  const int* first;
  double* listDist;     // sorted in place by (dist, source index)
  int32_t* listSrc;
  int32_t* out;         // nTgt x A
  PCC_HD void operator()(int64_t t) const
  {
    const int A = cfg.A;
    const int lo = first[t];
    int L = first[t + 1] - lo;
    double* ld = listDist + lo;
    int32_t* ls = listSrc + lo;
    const int32_t* c1 = refined1 + size_t(t) * A;
    if (L == 0) {
      for (int c = 0; c < A; c++)
        out[size_t(t) * A + c] = c1[c];
      return;
    }
    // std::sort by distance (the reference's order among equal distances is
    // unspecified; here: by source index)
    for (int i = 1; i < L; i++) {
      const double d = ld[i];
      const int32_t s = ls[i];
      int p = i;
      while (p > 0 && (ld[p - 1] > d || (ld[p - 1] == d && ls[p - 1] > s))) {
        ld[p] = ld[p - 1];
        ls[p] = ls[p - 1];
        p--;
      }
      ld[p] = d;
      ls[p] = s;
    }
    double centroid2[3] = {0.0, 0.0, 0.0};
    bool done = false;
    if (cfg.skipBwd && ld[0] < 0.0001) {
      L = 1;
      done = true;
    }
    while (!done) {
      if (L == 1) {
        done = true;
        break;
      }
      double maxAttr = 2.2250738585072014e-308;
      for (int i = 0; i < L; i++)
        for (int j = 0; j < L; j++) {
          double s = 0.0;
          for (int c = 0; c < A; c++) {
            const double df = double(srcAttr[size_t(ls[i]) * A + c])
              - double(srcAttr[size_t(ls[j]) * A + c]);
            s = dadd(s, dmul(df, df));
          }
          if (s > maxAttr)
            maxAttr = s;
        }
      if (maxAttr <= cfg.maxAttrBwd) {
        if (cfg.weightedBwd) {
          double sumW = 0.0;
          for (int i = 0; i < L; i++) {
            const double w = 1 / dadd(sqrt(ld[i]), cfg.distOffsetBwd);
            for (int c = 0; c < A; c++)
              centroid2[c] = dadd(centroid2[c], dmul(double(srcAttr[size_t(ls[i]) * A + c]), w));
            sumW = dadd(sumW, w);
          }
          for (int c = 0; c < A; c++)
            centroid2[c] = centroid2[c] / sumW;
        } else {
          for (int i = 0; i < L; i++)
            for (int c = 0; c < A; c++)
              centroid2[c] = dadd(centroid2[c], double(srcAttr[size_t(ls[i]) * A + c]));
          for (int c = 0; c < A; c++)
            centroid2[c] = centroid2[c] / double(L);
        }
        break;
      }
      L--;  // pop_back
    }
    if (done)
      for (int c = 0; c < A; c++)
        centroid2[c] = double(srcAttr[size_t(ls[0]) * A + c]);
    // fixWeight (m42538): w = 0, the starting point is centroid2
    double c0[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < A; c++)
      c0[c] = clip_round(dadd(dmul(0.0, double(c1[c])), dmul(1.0, centroid2[c])), cfg.clipMax);
    const double rSource = 1.0 / double(cfg.nSrc);
    const double rTarget = 1.0 / double(cfg.nTgt);
    double minError = 1.7976931348623157e308;
    double best[3] = {c0[0], c0[1], c0[2]};
    const int R = cfg.searchRange;
    const int R1 = A == 3 ? R : 0;  // the single-component search is one loop
    double col[3] = {0.0, 0.0, 0.0};
    for (int s1 = -R; s1 <= R; s1++) {
      col[0] = c0[0] + s1 < 0.0 ? 0.0 : c0[0] + s1 > cfg.clipMax ? cfg.clipMax : c0[0] + s1;
      for (int s2 = -R1; s2 <= R1; s2++) {
        if (A == 3)
          col[1] = c0[1] + s2 < 0.0 ? 0.0 : c0[1] + s2 > cfg.clipMax ? cfg.clipMax : c0[1] + s2;
        for (int s3 = -R1; s3 <= R1; s3++) {
          if (A == 3)
            col[2] = c0[2] + s3 < 0.0 ? 0.0 : c0[2] + s3 > cfg.clipMax ? cfg.clipMax : c0[2] + s3;
          double e1 = 0.0;
          for (int c = 0; c < A; c++) {
            const double df = dsub(col[c], double(c1[c]));
            e1 = dadd(e1, dmul(df, df));
          }
          e1 = dmul(e1, rTarget);
          double e2 = 0.0;
          for (int i = 0; i < L; i++)
            for (int c = 0; c < A; c++) {
              const double df = dsub(col[c], double(srcAttr[size_t(ls[i]) * A + c]));
              e2 = dadd(e2, dmul(df, df));
            }
          e2 = dmul(e2, rSource);
          const double err = e1 > e2 ? e1 : e2;
          if (err < minError) {
            minError = err;
            for (int c = 0; c < A; c++)
              best[c] = col[c];
          }
        }
      }
    }
    for (int c = 0; c < A; c++)
      out[size_t(t) * A + c] = int32_t(best[c]);
  }
};

struct FillI32ValueFn {
  int32_t* p;
  int32_t v;
  PCC_HD void operator()(int64_t i) const { p[i] = v; }
};

//----------------------------------------------------------------------------
// schedule

// builds the grid over n points (executor memory); the cell size is the
// smallest power of two that leaves at most n / 2 occupied cells (about two or
// more points per occupied cell), found by trying shifts (each try is one sort)
template<class Exec>
int
build_point_grid(Exec& ex, const int32_t* xyz, int n, PointGrid& g)
{
  int32_t* cell = ex.template alloc<int32_t>(size_t(n) * 3);
  int64_t* code = ex.template alloc<int64_t>(n);
  int32_t* order = ex.template alloc<int32_t>(n);
  int* flag = ex.template alloc<int>(2);
  int shift = 0;
  for (;; shift++) {
    ex.zero(flag, 2 * sizeof(int));
    ex.foreach(n, CellCoordFn{xyz, shift, cell, flag});
    ex.morton_sort(cell, n, code, order);
    ex.foreach(n, DistinctCodeFn{code, flag + 1});
    int h[2];
    ex.download(h, flag, sizeof(h));
    if (h[0] & 1)
      return PCCB200_ERR_INVALID_ARG;
    if (h[1] <= (n + 1) / 2 || shift >= 20)
      break;
  }
  int32_t* spos = ex.template alloc<int32_t>(size_t(n) * 3);
  int32_t* bbox = ex.template alloc<int32_t>(6);
  const int32_t init[6] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN};
  ex.upload(bbox, init, sizeof(init));
  ex.foreach(n, GatherPosFn{xyz, order, shift, spos, bbox});
  g.code = code;
  g.order = order;
  g.spos = spos;
  g.n = n;
  g.shift = shift;
  g.bbox = bbox;
  return PCCB200_OK;
}

// srcXyz / srcAttr / tgtXyz / out: executor memory.  Returns a PCCB200_* status.
template<class Exec>
int
recolour_run(Exec& ex, const pccb200_recolour_params& rp, const int32_t* srcXyz,
             const int32_t* srcAttr, int A, int nSrc, double sourceToTargetScale,
             const int32_t off[3], const int32_t* tgtXyz, int nTgt, int bitdepth, int32_t* out)
{
  if (nSrc <= 0 || nTgt <= 0 || (A != 1 && A != 3) || bitdepth < 1 || bitdepth > 16
      || rp.num_neighbours_fwd < 1 || rp.num_neighbours_fwd > kRecolourMaxK
      || rp.num_neighbours_bwd < 1 || rp.num_neighbours_bwd > kRecolourMaxK
      || rp.num_neighbours_fwd > nSrc || rp.num_neighbours_bwd > nTgt || rp.search_range < 0
      || rp.search_range > 8 || !(sourceToTargetScale > 0.0))
    return PCCB200_ERR_INVALID_ARG;
  RecolourConfig cfg;
  const double big = 1.7976931348623157e308;
  cfg.distOffsetFwd = rp.dist_offset_fwd;
  cfg.distOffsetBwd = rp.dist_offset_bwd;
  cfg.maxGeomFwd = rp.max_geometry_dist2_fwd < 512 ? rp.max_geometry_dist2_fwd : big;
  cfg.maxGeomBwd = rp.max_geometry_dist2_bwd < 512 ? rp.max_geometry_dist2_bwd : big;
  cfg.maxAttrFwd = rp.max_attribute_dist2_fwd < 512 ? rp.max_attribute_dist2_fwd : big;
  cfg.maxAttrBwd = rp.max_attribute_dist2_bwd < 512 ? rp.max_attribute_dist2_bwd : big;
  cfg.searchRange = rp.search_range;
  cfg.kFwd = rp.num_neighbours_fwd;
  cfg.kBwd = rp.num_neighbours_bwd;
  cfg.weightedFwd = rp.use_dist_weighted_avg_fwd != 0;
  cfg.weightedBwd = rp.use_dist_weighted_avg_bwd != 0;
  cfg.skipFwd = rp.skip_avg_if_identical_source_point_present_fwd != 0;
  cfg.skipBwd = rp.skip_avg_if_identical_source_point_present_bwd != 0;
  cfg.A = A;
  cfg.clipMax = double((1 << bitdepth) - 1);
  cfg.scale = sourceToTargetScale;
  cfg.invScale = 1.0 / sourceToTargetScale;
  for (int k = 0; k < 3; k++)
    cfg.off[k] = off[k];
  cfg.nSrc = nSrc;
  cfg.nTgt = nTgt;

  ex.phase(0);
  PointGrid gs, gt;
  int rc = build_point_grid(ex, srcXyz, nSrc, gs);
  if (rc != PCCB200_OK)
    return rc;
  rc = build_point_grid(ex, tgtXyz, nTgt, gt);
  if (rc != PCCB200_OK)
    return rc;

  //-- forward: every target in the source
  ex.phase(2);
  double* fDist = ex.template alloc<double>(size_t(nTgt) * cfg.kFwd);
  int32_t* fIdx = ex.template alloc<int32_t>(size_t(nTgt) * cfg.kFwd);
  ex.foreach(nTgt, KnnQueryFn{gs, cfg, tgtXyz, 0, cfg.kFwd, fDist, fIdx});
  int32_t* firstBad = ex.template alloc<int32_t>(1);
  const int32_t never = INT32_MAX;
  ex.upload(firstBad, &never, sizeof(never));
  ex.foreach(nTgt, FirstBadFn{fDist, cfg.kFwd, cfg.maxGeomFwd, firstBad});
  int32_t* refined1 = ex.template alloc<int32_t>(size_t(nTgt) * A);
  ex.foreach(nTgt, ForwardColourFn{cfg, fDist, fIdx, srcAttr, firstBad, refined1});

  //-- backward: every source in the target, lists per target
  double* bDist = ex.template alloc<double>(size_t(nSrc) * cfg.kBwd);
  int32_t* bIdx = ex.template alloc<int32_t>(size_t(nSrc) * cfg.kBwd);
  ex.foreach(nSrc, KnnQueryFn{gt, cfg, srcXyz, 1, cfg.kBwd, bDist, bIdx});
  int* first = ex.template alloc<int>(size_t(nTgt) + 1);
  ex.zero(first, (size_t(nTgt) + 1) * sizeof(int));
  ex.foreach(nSrc, BackwardCountFn{cfg, bDist, bIdx, first});
  ex.exclusive_scan(first, int64_t(nTgt) + 1);
  int* cursor = ex.template alloc<int>(nTgt);
  ex.zero(cursor, size_t(nTgt) * sizeof(int));
  const size_t maxPairs = size_t(nSrc) * cfg.kBwd;
  double* listDist = ex.template alloc<double>(maxPairs);
  int32_t* listSrc = ex.template alloc<int32_t>(maxPairs);
  ex.foreach(nSrc, BackwardFillFn{cfg, bDist, bIdx, first, cursor, listDist, listSrc});

  //-- the colour of every target
  ex.phase(3);
  ex.foreach(nTgt, FinalColourFn{cfg, refined1, srcAttr, first, listDist, listSrc, out});
  return PCCB200_OK;
}

}  // namespace pccb200
