// raht_block_warp.cuh — the top-down block transform as a warp-cooperative
// dataflow kernel (device only).  Same arithmetic, statement for statement, as
// BlockFn in raht_core.cuh (which stays the host-testable definition); what
// changes is the mapping onto the machine:
//
//   * one warp per block of siblings; lane = component * 8 + child slot, so
//     the 2x2x2 block of a colour attribute fills 24 lanes and the three
//     butterfly stages are __shfl_xor exchanges with lanes 1, 2 and 4 away;
//     the 18 neighbour look-ups are binary searches run by 18 lanes at once;
//   * everything lives in registers (no per-thread arrays in local memory);
//   * warps claim blocks in Morton order through a global ticket, a few
//     consecutive blocks at a time, so a warp may spin on the ready flag of
//     any earlier block (sub-node prediction reads the reconstruction of
//     earlier neighbour blocks) and on the zero-run look-back chain (RDOQ)
//     without any risk of deadlock;
//   * blocks with a single child never reach this kernel (PrepFn).
//
// Reference: the block loop of uraht_process, tmc3/RAHT.cpp:1306-1808.
#pragma once

#include "raht_core.cuh"

namespace pccb200 {

struct WarpBlockArgs {
  RahtConfig cfg;
  const QpTables* qt;
  Stage S;
  Stage P;            // P.n == 0: root block
  int32_t* coef;
  int64_t coefStride;
  int64_t coefBase;
  int qpLayer;
  int acLayer;
  int predInLvl;
  int* tz;               // look-back words of this stage, indexed by worklist rank
  const int32_t* worklist;  // block indices in Morton order (null for the root)
  int32_t* geom;            // kGeomStride ints per worklist entry (see k_block_geom)
  const int* count;         // number of worklist entries (device memory)
  int64_t ab11a, ab11b;     // RahtKernel(1, 1), the commonest butterfly
};

constexpr int kWarpBlockThreads = 256;
constexpr int kWarpBlockChunk = 1;  // blocks claimed per ticket (consecutive blocks in one
                                    // warp would serialise the zero-run look-back chain)
constexpr int kGeomStride = 20;     // ints per block: 19 neighbour indices + count

__device__ __forceinline__ int64_t
shfl_xor_i64(int64_t v, int m)
{
  return (int64_t)__shfl_xor_sync(0xffffffffu, (long long)v, m);
}
__device__ __forceinline__ int64_t
shfl_i64(int64_t v, int src)
{
  return (int64_t)__shfl_sync(0xffffffffu, (long long)v, src);
}

// one butterfly stage; every lane calls it (the shuffle is unconditional)
struct Bfly {
  int64_t a, b;
  bool both, swap, lo;
};

__device__ __forceinline__ int64_t
bfly_fwd(int64_t x, const Bfly& f, int dist, bool haar)
{
  const int64_t y = shfl_xor_i64(x, dist);
  if (f.both) {
    if (haar) {
      // lo holds left, hi holds right; hf = right - left
      if (f.lo) {
        int64_t hf = y - x;
        return x + ((hf >> (1 + kFracBits)) << kFracBits);
      }
      return x - y;
    }
    return f.lo ? fx_mul(y, f.b) + fx_mul(f.a, x) : fx_mul(x, f.a) - fx_mul(f.b, y);
  }
  return f.swap ? y : x;
}

__device__ __forceinline__ int64_t
bfly_inv(int64_t x, const Bfly& f, int dist, bool haar)
{
  const int64_t y = shfl_xor_i64(x, dist);
  if (f.both) {
    if (haar) {
      if (f.lo)  // x = lf, y = hf
        return x - ((y >> (1 + kFracBits)) << kFracBits);
      // x = hf, y = lf
      return x + (y - ((x >> (1 + kFracBits)) << kFracBits));
    }
    return f.lo ? fx_mul(x, f.a) - fx_mul(f.b, y) : fx_mul(y, f.b) + fx_mul(f.a, x);
  }
  return f.swap ? y : x;
}

// reconstruction values are exchanged between blocks through L2: relaxed
// 64-bit accesses, the value itself says whether it has been produced
__device__ __forceinline__ int64_t
ld_rec(const int64_t* p)
{
  long long v;
  asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return int64_t(v);
}
__device__ __forceinline__ void
st_rec(int64_t* p, int64_t v)
{
  asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(p), "l"((long long)v) : "memory");
}
__device__ __forceinline__ int64_t
poll_rec(const int64_t* p)
{
  int64_t v = ld_rec(p);
  while (v == kRecNotReady) {
    __nanosleep(32);
    v = ld_rec(p);
  }
  return v;
}

// Warp-parallel decoupled look-back: the zero-run counter after block q-1
// (the resolved value of word q).  32 words are examined per step; transparent
// blocks are summed through, the walk ends at the nearest published exit
// state and waits only on words that have not been published at all.
__device__ __forceinline__ int
tz_lookback_warp(const int* tz, int q, const int lane)
{
  int acc = 0;
  for (;;) {
    const int idx = q - lane;
    const int w = idx >= 0 ? ld_acquire(&tz[idx]) : tz_pack(kTzExit, 0);
    const int st = tz_status(w);
    const unsigned exitMask = __ballot_sync(0xffffffffu, st == kTzExit);
    const unsigned noneMask = __ballot_sync(0xffffffffu, st == kTzNone);
    const int firstExit = exitMask ? __ffs(exitMask) - 1 : 32;
    const int firstNone = noneMask ? __ffs(noneMask) - 1 : 32;
    const int stop = firstExit < firstNone ? firstExit : firstNone;
    // transparent words nearer than the stopping lane
    acc += __reduce_add_sync(0xffffffffu, (lane < stop && st == kTzTransparent) ? tz_value(w) : 0);
    if (firstExit < firstNone)
      return acc + tz_value(__shfl_sync(0xffffffffu, w, firstExit));
    q -= stop;  // all 32 transparent (stop == 32), or wait at the unpublished word
    if (firstNone < 32)
      __nanosleep(40);
  }
}

// every kTzCheckpoint-th block resolves its exit state even if it is
// transparent, which bounds the length of every look-back walk
constexpr int kTzCheckpoint = 32;

// processes block p (worklist rank t); called by all 32 lanes
__device__ __forceinline__ void
warp_block(const WarpBlockArgs& a, const int p, const int t, const int lane)
{
  const RahtConfig& cfg = a.cfg;
  const Stage& S = a.S;
  const Stage& P = a.P;
  const int A = cfg.A;
  const int j = lane & 7;
  const int k = lane >> 3;
  const bool act = k < A;
  const bool root = P.n == 0;
  const bool haar = cfg.haar != 0;
  const bool ext = cfg.ext != 0;
  const bool enc = cfg.isEncoder != 0;

  const int c0 = root ? 0 : P.first[p];
  uint32_t occ;
  if (root) {
    uint32_t bit = lane < S.n ? 1u << int((S.key[lane] >> S.level) & 7) : 0u;
    occ = __reduce_or_sync(0xffffffffu, bit);
  } else {
    occ = P.occ[p];
  }
  const bool present = (occ >> j) & 1;
  const int cidx = c0 + __popc(occ & ((1u << j) - 1));
  const int w0 = present ? S.weight[cidx] : 0;
  int nodeQp0 = 0, nodeQp1 = 0;
  if (cfg.hasQp) {
    if (root) {
      if (lane == 0)
        descend_qps(S, 0, S.n, nullptr);
      __syncwarp();
    }
    if (present) {
      nodeQp0 = S.qpDown[2 * cidx] >> 4;
      nodeQp1 = S.qpDown[2 * cidx + 1] >> 4;
    }
  }
  int64_t buf = 0;
  if (enc && act && present)
    buf = fx_from_int(S.attr[size_t(cidx) * A + k]);

  //-- weight tree and butterfly constants (mkWeightTree + RahtKernel)
  Bfly bf[3];
  int wcur = w0;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int d = 1 << s;
    const int wp = __shfl_xor_sync(0xffffffffu, wcur, d);
    const bool lo = !(j & d);
    const int wl = lo ? wcur : wp;
    const int wr = lo ? wp : wcur;
    bf[s].lo = lo;
    bf[s].both = wl && wr;
    bf[s].swap = !wl && wr;
    bf[s].a = bf[s].b = 0;
    if (bf[s].both && !haar) {
      if (wl == 1 && wr == 1) {
        bf[s].a = a.ab11a;
        bf[s].b = a.ab11b;
      } else {
        raht_ab(wl, wr, bf[s].a, bf[s].b);
      }
    }
    wcur = (lo || bf[s].both) ? wl + wr : 0;
  }
  const int wfin = wcur;  // weights[24 + j]

  // 1/sqrt(w) scaling of this lane's child (used for the sums and the store)
  int rsShift = 0;
  int64_t rsMul = 0;
  if (!haar && w0 > 1) {
    rsShift = w0 > 1024 ? ilog2_u64(uint64_t(w0 - 1)) >> 1 : 0;
    rsMul = int64_t(irsqrt64(uint64_t(w0)) >> (40 - rsShift - kFracBits));
  }

  //-- prediction gating: neighbour indices and count come from k_block_geom
  bool enablePred = false;
  int pidxLane = -1;
  if (a.predInLvl) {
    const int g = lane < kGeomStride ? a.geom[size_t(t) * kGeomStride + lane] : -1;
    const int count = __shfl_sync(0xffffffffu, g, 19);
    pidxLane = lane < 19 ? g : -1;
    enablePred = count >= cfg.thr1 && __shfl_sync(0xffffffffu, g, 0) >= 0;
  } else if (root && present && k == 0) {
    S.nn[cidx] = 19;
  }

  //-- encoder: normalise and transform the sums
  if (enc) {
    if (rsMul)
      buf = fx_mul(buf >> rsShift, rsMul);
#pragma unroll
    for (int s = 0; s < 3; s++)
      buf = bfly_fwd(buf, bf[s], 1 << s, haar);
  }

  //-- prediction (intraDcPred)
  int64_t pred = 0;
  if (enablePred) {
    int wsum = -1;
    int64_t limLow = 0, limHigh = 0;
    const int64_t fracMul = ext ? 1 : (int64_t(1) << kFracBits);
    const int parentOnly = cfg.subnode ? 7 : 19;
    for (int i = 0; i < 19; i++) {
      const int q = __shfl_sync(0xffffffffu, pidxLane, i);
      if (q < 0)
        continue;
      const int64_t v0 = P.rec[size_t(q) * A];
      if (i) {
        if (10 * v0 <= limLow || 10 * v0 >= limHigh)
          continue;
      } else {
        limLow = 2 * v0;
        limHigh = 25 * v0;
      }
      const int64_t mine = act ? P.rec[size_t(q) * A + k] : 0;
      const int wp = cfg.predWeightParent[i];
      const uint32_t mask = uint32_t(neigh_mask(i)) & occ;
      uint32_t cmask = 0, nocc = 0;
      int shift = 0, cfirst = 0;
      if (i >= parentOnly && q < p) {
        const int ii = i - 7;
        const int sh = occu_shift(ii);
        shift = ii < 9 ? sh : -sh;
        nocc = P.occ[q];
        cmask = (ii < 9 ? (nocc >> sh) : (nocc << sh)) & mask & 0xffu;
        if (cmask)
          cfirst = P.first[q];
      }
      if ((mask >> j) & 1) {
        if ((cmask >> j) & 1) {
          const int wc = cfg.predWeightChild[i - 7];
          const int c = cfirst + __popc(nocc & ((1u << (j + shift)) - 1));
          wsum += wc;
          if (act)  // produced by an earlier block of this stage: poll the value
            pred += poll_rec(&S.rec[size_t(c) * A + k]) * (wc * fracMul);
        } else {
          wsum += wp;
          pred += mine * (wp * fracMul);
        }
      }
    }
    if (present && act) {
      const int d = wsum + 1;
      const int64_t div = (32768 + d / 2) / d;
      int64_t v = fx_mul(pred, div);
      if (haar)
        v = (v >> kFracBits) << kFracBits;
      else if (w0 > 1)
        v = fx_mul(v, int64_t(isqrt64(uint64_t(w0) << (2 * kFracBits))));
      pred = v;
    } else {
      pred = 0;
    }
#pragma unroll
    for (int s = 0; s < 3; s++)
      pred = bfly_fwd(pred, bf[s], 1 << s, haar);
  }

  //-- coefficients: lane (j, k) owns coefficient j of component k
  const bool exists = j == 0 ? root : wfin != 0;
  const uint32_t existsMask = __ballot_sync(0xffffffffu, exists) & 0xffu;
  // bit i' set in before(j): coefficient i' precedes j in scan order 0,4,2,1,6,5,3,7
  const uint32_t before =
    j == 0 ? 0x00u : j == 4 ? 0x01u : j == 2 ? 0x11u : j == 1 ? 0x15u
    : j == 6 ? 0x17u : j == 5 ? 0x57u : j == 3 ? 0x77u : 0x7fu;
  const int ncoef = __popc(existsMask);

  LayerQp lq;
  lq.luma = a.qt->layers[a.qpLayer][0];
  lq.chromaOffset = a.qt->layers[a.qpLayer][1];
  lq.maxQp = cfg.maxQp;
  lq.fixedPointQpOffset = cfg.fixedPointQpOffset;

  if (enc && enablePred && exists)
    buf -= pred;

  bool flagMine = false;
  const int myPos = __popc(existsMask & before);
  const bool rdoq = enc && !haar;
  if (rdoq) {
    int64_t d2 = 0, aq = 0, lam = 0;
    int rc = 0;
    if (exists) {
      Quantizer q[2];
      make_quantizers(lq, nodeQp0, nodeQp1, q);
      if (act) {
        const int64_t c = fx_round(buf);
        d2 = c * c;
        const int64_t qc = q[k < 1 ? k : 1].quantize(c << kAttrShift);
        aq = qc < 0 ? -qc : qc;
        rc = lut_log(aq);
      }
      const int64_t l0 = q[0].scale(1);
      lam = l0 * l0 * (A == 1 ? 25 : 35);
    }
    // sums over the components (lanes 8 and 16 away)
    d2 += shfl_xor_i64(d2, 8);
    d2 += shfl_xor_i64(d2, 16);
    aq += shfl_xor_i64(aq, 8);
    aq += shfl_xor_i64(aq, 16);
    rc += __shfl_xor_sync(0xffffffffu, rc, 8);
    rc += __shfl_xor_sync(0xffffffffu, rc, 16);
    const int kindMine = aq == 0 ? 0 : (aq < 3 ? 1 : 2);

    // the block's coefficients in scan order, replicated in every lane
    int kind[8], rcs[8];
    int64_t d2s[8], lams[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
      kind[m] = 0;
      rcs[m] = 0;
      d2s[m] = 0;
      lams[m] = 0;
    }
    int n = 0;
    const int kScan[8] = {0, 4, 2, 1, 6, 5, 3, 7};
#pragma unroll
    for (int si = 0; si < 8; si++) {
      const int idx = kScan[si];
      const int kd = __shfl_sync(0xffffffffu, kindMine, idx);
      const int r = __shfl_sync(0xffffffffu, rc, idx);
      const int64_t dd = shfl_i64(d2, idx);
      const int64_t ll = shfl_i64(lam, idx);
      const bool ex = (existsMask >> idx) & 1;  // uniform across the warp
#pragma unroll
      for (int m = 0; m < 8; m++)
        if (ex && m == n) {
          kind[m] = kd;
          rcs[m] = r;
          d2s[m] = dd;
          lams[m] = ll;
        }
      n += ex;
    }

    // zero-run protocol, as in BlockFn (all decisions are warp-uniform)
    const bool zeroMatters = cfg.numAcLayers > 0;
    int firstSensitive = -1, firstHard = -1, lastHard = -1;
    bool anySoft = false;
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (m < ncoef) {
        if ((kind[m] == 1 || (kind[m] == 0 && zeroMatters)) && firstSensitive < 0)
          firstSensitive = m;
        anySoft |= kind[m] == 1;
        if (kind[m] == 2) {
          if (firstHard < 0)
            firstHard = m;
          lastHard = m;
        }
      }
    const bool anyHard = lastHard >= 0;
    bool published = false;
    if (!anySoft && !anyHard && lane == 0)
      st_release(&a.tz[t + 1], tz_pack(kTzTransparent, ncoef));
    if (anyHard) {
      int tt = 0;
#pragma unroll
      for (int m = 0; m < 8; m++)
        if (m > lastHard && m < ncoef)
          tt = BlockFn::step_tz(tt, kind[m], d2s[m], lams[m], rcs[m], nullptr);
      if (lane == 0)
        st_release(&a.tz[t + 1], tz_pack(kTzExit, tt));
      published = true;
    }
    const bool needEntry =
      firstSensitive >= 0 && (firstHard < 0 || firstSensitive < firstHard);
    int tt = needEntry ? tz_lookback_warp(a.tz, t, lane) : 0;
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (m < ncoef) {
        bool f = false;
        tt = BlockFn::step_tz(tt, kind[m], d2s[m], lams[m], rcs[m], &f);
        if (m == myPos)
          flagMine = f;
      }
    if (!published && needEntry) {
      if (lane == 0)
        st_release(&a.tz[t + 1], tz_pack(kTzExit, tt));
      published = true;
    }
    if (!published && (t % kTzCheckpoint) == kTzCheckpoint - 1) {
      // transparent checkpoint block: resolve the exit state anyway
      const int e = tz_lookback_warp(a.tz, t, lane) + ncoef;
      if (lane == 0)
        st_release(&a.tz[t + 1], tz_pack(kTzExit, e));
    }
  }

  //-- quantise / dequantise (RAHT.cpp:1672-1723)
  {
    int off0 = nodeQp0, off1 = nodeQp1;
    if (j && a.acLayer < cfg.numAcLayers) {
      off0 += a.qt->acQps[a.acLayer][j - 1][0];
      off1 += a.qt->acQps[a.acLayer][j - 1][1];
    }
    if (exists && act) {
      Quantizer q[2];
      make_quantizers(lq, off0, off1, q);
      const Quantizer& qk = q[k < 1 ? k : 1];
      const int64_t pos = a.coefBase + c0 - (root ? 0 : p) + myPos;
      int64_t qc;
      if (enc) {
        const int64_t c = flagMine ? 0 : fx_round(buf);
        qc = qk.quantize(c << kAttrShift);
        a.coef[k * a.coefStride + pos] = int32_t(qc);
      } else {
        qc = a.coef[k * a.coefStride + pos];
      }
      pred += fx_from_int(div_exp2_round_half_up(qk.scale(qc), kAttrShift));
    }
  }

  //-- DC from the parent, inverse transform, store (RAHT.cpp:1726-1806)
  if (!root && j == 0 && act) {
    const int64_t v = P.recUs[size_t(p) * A + k];
    pred = ext ? v : v * (int64_t(1) << (kFracBits - 2));
  }
#pragma unroll
  for (int s = 2; s >= 0; s--)
    pred = bfly_inv(pred, bf[s], 1 << s, haar);
  if (present && act) {
    int64_t v = pred;
    S.recUs[size_t(cidx) * A + k] = ext ? v : fx_round(v * 4);
    if (rsMul)
      v = fx_mul(v >> rsShift, rsMul);
    st_rec(&S.rec[size_t(cidx) * A + k], ext ? v : fx_round(v));
  }
}

// Geometry-only part of a stage, one warp per transforming block, fully
// parallel: prediction gating (RAHT.cpp:1391-1432) and the 18 bounded
// neighbour searches of findNeighbours (RAHT.cpp:299-368), one per lane.
// Output per worklist entry t: geom[t*20 + i] = parent-stage index of
// neighbour i (i = 0 is the block's own parent) or -1; geom[t*20 + 19] = the
// neighbour count handed to the children (numParentNeigh).  A block whose
// grandparent count fails threshold0 gets index -1 in slot 0 (no search).
__global__ void __launch_bounds__(256)
k_block_geom(const WarpBlockArgs a)
{
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= *a.count)
    return;
  const RahtConfig& cfg = a.cfg;
  const Stage& S = a.S;
  const Stage& P = a.P;
  const int p = a.worklist[t];
  const uint32_t occ = P.occ[p];
  int pidx = -1;
  int count = 0;
  if (P.nn[p] >= cfg.thr0) {
    const int plevel = S.level + 3;
    const int64_t cur = P.key[p] >> plevel;
    const int64_t base = int64_t(morton3d_add(uint64_t(cur), ~uint64_t(0)));
    if (lane == 0)
      pidx = p;
    else if (lane < 19 && (occ & neigh_mask(lane)))
      pidx = find_parent_neighbour(P, p, plevel, cur, base, lane, cfg.searchRange);
    count = __popc(__ballot_sync(0xffffffffu, pidx >= 0));
  }
  if (lane < 19)
    a.geom[size_t(t) * kGeomStride + lane] = pidx;
  else if (lane == 19)
    a.geom[size_t(t) * kGeomStride + 19] = count;
  // the count is inherited by the children's blocks at the next stage
  if (lane < 8 && ((occ >> lane) & 1))
    S.nn[P.first[p] + __popc(occ & ((1u << lane) - 1))] = count;
}

__global__ void __launch_bounds__(kWarpBlockThreads)
k_block_warp(const WarpBlockArgs a, unsigned long long* ticket)
{
  const int lane = threadIdx.x & 31;
  const int n = *a.count;
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0)
      base = atomicAdd(ticket, (unsigned long long)kWarpBlockChunk);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned long long)n)
      return;
#pragma unroll 1
    for (int i = 0; i < kWarpBlockChunk; i++) {
      const int t = int(base) + i;
      if (t >= n)
        break;
      const int p = a.worklist ? a.worklist[t] : 0;
      warp_block(a, p, t, lane);
    }
  }
}

}  // namespace pccb200
