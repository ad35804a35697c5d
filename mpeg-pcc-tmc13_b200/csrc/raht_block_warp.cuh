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
  int experiment;           // timing experiments only (PCCB200_EXPERIMENT), 0 in production
  const struct TzRegion* regions;  // zero-run words/lists of every stage so far
  int stageIdx;                    // index of this stage in regions (0 = root)
};

// Zero-run bookkeeping of one stage, indexed by worklist rank t:
// words[t + 1] is block t's state word, lists[(t + 1) * 8 + i] its i-th
// coefficient in scan order (0: never resets the run, a > 0: resets it unless
// the run before it is at least a long).
struct TzRegion {
  int* words;
  int* lists;
  const int* count;
};

constexpr int kTzClassified = 3;  // word status: list published, outcome pending

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

// RDOQ threshold of a coefficient: the smallest zero-run length tz for which
// the reference's test (RAHT.cpp:1617-1636)
//     (Dist2 << 26) < lambda * (Rate(tz) + ((Ratecoeff + 128) >> 8))
// holds.  Rate(tz) is a non-decreasing step function; it changes at
// tz = 0,1,2,3,5,7,9 and at tz = 10 + 2^(a-1), a >= 1.  INT_MAX: never.
__device__ __forceinline__ int
rdoq_threshold(int64_t dist2, int64_t lambda, int rateCoeff)
{
  const int64_t lhs = dist2 << 26;
  const int rc = (rateCoeff + 128) >> 8;
  const int kTz[7] = {0, 1, 2, 3, 5, 7, 9};
#pragma unroll
  for (int i = 0; i < 7; i++)
    if (lhs < lambda * (zero_run_rate(kTz[i]) + rc))
      return kTz[i];
  for (int a = 1; a <= 30; a++) {
    const int tz = 10 + (1 << (a - 1));
    if (lhs < lambda * (zero_run_rate(tz) + rc))
      return tz;
  }
  return 0x7fffffff;
}

// Is the run of non-resetting coefficients that ends just before block t of
// stage a.stageIdx at least A long?  Walks back over the published
// classification of earlier blocks (this stage, then earlier stages); waits
// only for blocks that have not classified their coefficients yet, never for
// another block's own answer.  All lanes execute it uniformly.
__device__ __forceinline__ bool
tz_run_at_least(const WarpBlockArgs& a, int t, int A)
{
  if (A <= 0)
    return true;
  int req = A;   // positions 1..req behind the block must not reset the run
  int acc = 0;   // positions already verified
  int s = a.stageIdx;
  const TzRegion* rg = &a.regions[s];
  int u = t - 1;
  for (;;) {
    if (u < 0) {
      if (--s < 0)
        return acc >= req;  // start of the call: the counter starts at 0
      rg = &a.regions[s];
      u = *rg->count - 1;
      continue;
    }
    int w;
    while (tz_status(w = ld_acquire(&rg->words[u + 1])) == kTzNone)
      __nanosleep(40);
    const int st = tz_status(w), v = tz_value(w);
    if (st == kTzExit)
      return v + acc >= req;
    if (st == kTzClassified) {
      const int4 l0 = *reinterpret_cast<const int4*>(&rg->lists[size_t(u + 1) * 8]);
      const int4 l1 = *reinterpret_cast<const int4*>(&rg->lists[size_t(u + 1) * 8 + 4]);
      const int li[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int i = 7; i >= 0; i--)
        if (i < v) {
          const int pos = acc + (v - i);
          if (pos > req)
            return true;
          if (li[i] > 0 && pos + li[i] > req)
            req = li[i] >= 0x40000000 ? 0x7fffffff : pos + li[i];
        }
    }
    acc += v;
    if (acc >= req)
      return true;
    u--;
  }
}

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
    // classification of this lane's coefficient: 0 = never resets the run
    // (all components quantise to zero); kAlwaysRemoved = RDOQ removes it at
    // any run length (never resets either); INT_MAX = always resets the run
    // (sum of |q| >= 3, or RDOQ never fires); otherwise the run length from
    // which RDOQ removes it
    constexpr int kAlwaysRemoved = -2;
    int thrMine = 0;
    if (aq >= 3) {
      thrMine = 0x7fffffff;
    } else if (aq > 0) {
      thrMine = rdoq_threshold(d2, lam, rc);
      if (thrMine == 0)
        thrMine = kAlwaysRemoved;
    }

    // the block's coefficients in scan order, replicated in every lane
    int thr[8];
#pragma unroll
    for (int m = 0; m < 8; m++)
      thr[m] = 0;
    {
      int n = 0;
      const int kScan[8] = {0, 4, 2, 1, 6, 5, 3, 7};
#pragma unroll
      for (int si = 0; si < 8; si++) {
        const int idx = kScan[si];
        const int th = __shfl_sync(0xffffffffu, thrMine, idx);
        const bool ex = (existsMask >> idx) & 1;  // uniform across the warp
#pragma unroll
        for (int m = 0; m < 8; m++)
          if (ex && m == n)
            thr[m] = th;
        n += ex;
      }
    }
    bool hasS = false, hasH = false;
    int lastH = -1;
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (m < ncoef) {
        if (thr[m] == 0x7fffffff) {
          hasH = true;
          lastH = m;
        } else if (thr[m] > 0) {
          hasS = true;
        }
      }

    // publish what is known without looking at any other block
    const TzRegion rg = a.regions[a.stageIdx];
    if (hasH) {
      int tl = 0;
#pragma unroll
      for (int m = 0; m < 8; m++)
        if (m > lastH && m < ncoef)
          tl = tl >= thr[m] ? tl + 1 : 0;
      if (lane == 0)
        st_release(&rg.words[t + 1], tz_pack(kTzExit, tl));
    } else if (!hasS) {
      if (lane == 0)
        st_release(&rg.words[t + 1], tz_pack(kTzTransparent, ncoef));
    } else {
      int mine = 0;
#pragma unroll
      for (int m = 0; m < 8; m++)
        if (m == lane)
          mine = thr[m] > 0 ? thr[m] : 0;
      if (lane < 8)
        rg.lists[size_t(t + 1) * 8 + lane] = mine;
      __threadfence();
      __syncwarp();
      if (lane == 0)
        st_release(&rg.words[t + 1], tz_pack(kTzClassified, ncoef));
    }

    // resolve this block's own decisions
    bool linked = true;  // the run still reaches back beyond the block
    int z = 0;           // its length inside the block while linked
    int tl = 0;          // run length since the last reset inside the block
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (m < ncoef) {
        bool f;
        if (thr[m] <= 0)
          f = thr[m] == kAlwaysRemoved;
        else if (thr[m] == 0x7fffffff)
          f = false;
        else if (linked)
          f = a.experiment == 1 ? false : tz_run_at_least(a, t, thr[m] - z);
        else
          f = tl >= thr[m];
        const bool keeps = thr[m] <= 0 || f;
        if (linked) {
          if (keeps)
            z++;
          else {
            linked = false;
            tl = 0;
          }
        } else {
          tl = keeps ? tl + 1 : 0;
        }
        if (m == myPos)
          flagMine = f;
      }
    if (hasS && !hasH && lane == 0)
      st_release(&rg.words[t + 1],
                 linked ? tz_pack(kTzTransparent, ncoef) : tz_pack(kTzExit, tl));
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
