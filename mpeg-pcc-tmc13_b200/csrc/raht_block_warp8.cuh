// raht_block_warp8.cuh — the block transform with FOUR blocks per warp
// (device only).  Same arithmetic and the same published zero-run protocol as
// warp_block in raht_block_warp.cuh; what changes is the mapping:
//
//   * a block of siblings occupies a group of 8 lanes (lane = child slot); the
//     attribute components are a short in-lane loop.  Blocks of sparse clouds
//     have two or three children, so one warp per block left most lanes idle
//     and the kernel was bound by instruction issue once many slices / frames
//     were in flight; four independent blocks per warp cut the warp
//     instructions per block by 2-4x;
//   * all exchanges are shuffles / ballots restricted to the 8 lanes of the
//     group (member mask);
//   * NOTHING blocks inside a group.  Two blocks of one warp may depend on
//     each other (sub-node prediction reads the reconstruction of earlier
//     blocks; the zero-run walk reads their classification), and the compiler
//     reconverges the warp after every divergent region, so a group spinning
//     on a value that a sibling group has yet to produce would hang the warp.
//     Instead each group is a small state machine (predict -> resolve ->
//     done) that the warp steps in a converged loop: a step that finds an
//     input missing gives up and is retried in the next round;
//   * a warp claims one "column" of a 32-ticket tile: claim c serves tickets
//     (c / 8) * 32 + g * 8 + c % 8, g = 0..3, so that neighbouring blocks (which
//     depend on each other most often) sit in different warps.
//
// Reference: the block loop of uraht_process, tmc3/RAHT.cpp:1306-1808.
#pragma once

#include "raht_block_warp.cuh"

namespace pccb200 {

constexpr int kGroupsPerWarp = 4;
constexpr int kTicketInterleave = kWarpBlockThreads / 32;  // claims per tile

// RDOQ classification of a coefficient, 6 bits (eight of them, in scan order,
// make the block's published list):
//   0      all components quantise to zero: never resets the run
//   1      RDOQ removes it whatever the run length: never resets either
//   2      always resets the run (sum |q| >= 3, or RDOQ never fires)
//   3..8   removed if the run before it is at least 1,2,3,5,7,9 long
//   8 + a  removed if the run before it is at least 10 + 2^(a-1), a = 1..30
constexpr int kCodeZero = 0;
constexpr int kCodeRemoved = 1;
constexpr int kCodeHard = 2;

__device__ __forceinline__ int
thr_decode(int code)  // code >= 3
{
  return code < 9 ? int((0x975321u >> (4 * (code - 3))) & 15u) : 10 + (1 << (code - 9));
}

// rdoq_threshold (raht_block_warp.cuh), returning the code of the threshold
__device__ __forceinline__ int
rdoq_code(int64_t dist2, int64_t lambda, int rateCoeff)
{
  const int64_t lhs = dist2 << 26;
  const int rc = (rateCoeff + 128) >> 8;
  const int kRate[7] = {1, 2, 3, 5, 7, 9, 11};  // zero_run_rate of 0,1,2,3,5,7,9
#pragma unroll
  for (int i = 0; i < 7; i++)
    if (lhs < lambda * (kRate[i] + rc))
      return i == 0 ? kCodeRemoved : 2 + i;
  if (lhs >= lambda * (72 + rc))  // zero_run_rate(10 + 2^29) = 72
    return kCodeHard;
  for (int a = 1; a <= 30; a++)
    if (lhs < lambda * (12 + 2 * a + rc))  // zero_run_rate(10 + 2^(a-1))
      return 8 + a;
  return kCodeHard;
}

// resumable form of tz_run_at_least: 1 / 0 = answer, -1 = a block on the way
// has not published yet (call again later, the state is kept)
struct TzWalk {
  int req, acc, s, u;
  const int* words;
  const unsigned long long* lists;
};

__device__ __forceinline__ void
tz_walk_enter(const WarpBlockArgs& a, TzWalk& w, int s)
{
  const TzRegion rg = a.regions[s];
  w.s = s;
  w.words = rg.words;
  w.lists = reinterpret_cast<const unsigned long long*>(rg.lists);
  w.u = *rg.count - 1;
}

// Executed by the 8 lanes of a group together.  What a lane reads here changes
// under its feet (a word goes from "nothing" to "classified" to its final
// value), and lanes of a group are not guaranteed to run in lock step, so every
// lane follows the observation of the group's first lane: the decision
// "pending or not" must be the same in all eight, or the group would split
// between two states and wait for itself at the next shuffle.
__device__ __forceinline__ int
tz_walk(const WarpBlockArgs& a, TzWalk& w, unsigned gm)
{
  for (;;) {
    if (w.u < 0) {
      if (w.s == 0)
        return w.acc >= w.req;  // start of the call: the counter starts at 0
      tz_walk_enter(a, w, w.s - 1);
      continue;
    }
    const int word = __shfl_sync(gm, ld_acquire(&w.words[w.u + 1]), 0, 8);
    const int st = tz_status(word), v = tz_value(word);
    if (st == kTzNone)
      return -1;
    if (st == kTzExit)
      return v + w.acc >= w.req;
    if (st == kTzClassified) {
      // (lane 0 saw the word after its list was written)
      const unsigned long long L =
        (unsigned long long)__shfl_sync(gm, (long long)w.lists[w.u + 1], 0, 8);
      for (int i = v - 1; i >= 0; i--) {
        const int pos = w.acc + (v - i);
        if (pos > w.req)
          return 1;
        const int code = int((L >> (6 * i)) & 63);
        if (code >= 3) {
          const int li = thr_decode(code);
          if (pos + li > w.req)
            w.req = pos + li;
        }
      }
    }
    w.acc += v;
    if (w.acc >= w.req)
      return 1;
    w.u--;
  }
}

__device__ __forceinline__ int64_t
g_shfl_xor_i64(unsigned gm, int64_t v, int m)
{
  return (int64_t)__shfl_xor_sync(gm, (long long)v, m);
}

__device__ __forceinline__ int64_t
g_bfly_fwd(unsigned gm, int64_t x, const Bfly& f, int dist, bool haar)
{
  const int64_t y = g_shfl_xor_i64(gm, x, dist);
  if (f.both) {
    if (haar) {
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
g_bfly_inv(unsigned gm, int64_t x, const Bfly& f, int dist, bool haar)
{
  const int64_t y = g_shfl_xor_i64(gm, x, dist);
  if (f.both) {
    if (haar) {
      if (f.lo)
        return x - ((y >> (1 + kFracBits)) << kFracBits);
      return x + (y - ((x >> (1 + kFracBits)) << kFracBits));
    }
    return f.lo ? fx_mul(x, f.a) - fx_mul(f.b, y) : fx_mul(y, f.b) + fx_mul(f.a, x);
  }
  return f.swap ? y : x;
}

#ifdef PCCB200_WATCHDOG
// debugging aid: a group that has been stuck for seconds reports and traps
__device__ int g_wdMinTicket = 0x7fffffff;
__device__ unsigned long long g_wdDeadline = 0;
__device__ unsigned long long* g_wdTicket = nullptr;
__device__ __forceinline__ unsigned long long
global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

constexpr int kStPredict = 0;
constexpr int kStResolve = 1;
constexpr int kStDone = 2;

// The four blocks of one claim: group g (lanes 8g..8g+7) runs worklist rank t
// when `active`.  Called by all 32 lanes.
__device__ __forceinline__ void
group_blocks(const WarpBlockArgs& a, const int t, const bool active, const int lane)
{
  const RahtConfig& cfg = a.cfg;
  const Stage& S = a.S;
  const Stage& P = a.P;
  const int A = cfg.A;
  const int j = lane & 7;
  const int gbase = lane & ~7;
  const unsigned gm = 0xffu << gbase;  // member mask of the group
  const bool root = P.n == 0;
  const bool haar = cfg.haar != 0;
  const bool ext = cfg.ext != 0;
  const bool enc = cfg.isEncoder != 0;
  const bool rdoq = enc && !haar;

  int state = active ? kStPredict : kStDone;

  // block state (registers), meaningful while state != kStDone
  int p = 0, c0 = 0, cidx = 0, w0 = 0;
  uint32_t occ = 0;
  bool present = false;
  int nodeQp0 = 0, nodeQp1 = 0;
  int64_t buf[3] = {0, 0, 0};
  int64_t pred[3] = {0, 0, 0};
  Bfly bf[3];
  int wfin = 0;
  int rsShift = 0;
  int64_t rsMul = 0;
  bool enablePred = false;
  int nq[3] = {-1, -1, -1};          // neighbour indices i = j, j + 8, j + 16
  uint32_t nocc[3] = {0, 0, 0};      // their occupancy, when their children may be used
  int nfirst[3] = {0, 0, 0};         // and first child
  uint32_t validMask = 0;            // neighbours that contribute (19 bits, group uniform)
  const int64_t* missing = nullptr;  // an input of the prediction not yet produced

#pragma unroll
  for (int s = 0; s < 3; s++) {
    bf[s].a = bf[s].b = 0;
    bf[s].both = bf[s].swap = bf[s].lo = false;
  }

  //==========================================================================
  // everything that depends on finished stages only
  if (active) {
    p = a.worklist ? a.worklist[t] : 0;
    c0 = root ? 0 : P.first[p];
    if (root) {
      uint32_t bit = j < S.n ? 1u << int((S.key[j] >> S.level) & 7) : 0u;
      bit |= __shfl_xor_sync(gm, bit, 1);
      bit |= __shfl_xor_sync(gm, bit, 2);
      bit |= __shfl_xor_sync(gm, bit, 4);
      occ = bit;
    } else {
      occ = P.occ[p];
    }
    present = (occ >> j) & 1;
    cidx = c0 + __popc(occ & ((1u << j) - 1));
    w0 = present ? S.weight[cidx] : 0;
    if (cfg.hasQp) {
      if (root) {
        if (j == 0)
          descend_qps(S, 0, S.n, nullptr);
        __syncwarp(gm);
      }
      if (present) {
        nodeQp0 = S.qpDown[2 * cidx] >> 4;
        nodeQp1 = S.qpDown[2 * cidx + 1] >> 4;
      }
    }
    if (enc && present)
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (k < A)
          buf[k] = fx_from_int(S.attr[size_t(cidx) * A + k]);

    //-- weight tree and butterfly constants (mkWeightTree + RahtKernel)
    int wcur = w0;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const int d = 1 << s;
      const int wp = __shfl_xor_sync(gm, wcur, d);
      const bool lo = !(j & d);
      const int wl = lo ? wcur : wp;
      const int wr = lo ? wp : wcur;
      bf[s].lo = lo;
      bf[s].both = wl && wr;
      bf[s].swap = !wl && wr;
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
    wfin = wcur;  // weights[24 + j]

    if (!haar && w0 > 1) {
      rsShift = w0 > 1024 ? ilog2_u64(uint64_t(w0 - 1)) >> 1 : 0;
      rsMul = int64_t(irsqrt64(uint64_t(w0)) >> (40 - rsShift - kFracBits));
    }

    //-- prediction gating and the neighbours' parent-stage data, three
    //   neighbours per lane (geometry from k_block_geom; slot 19 = count)
    if (a.predInLvl) {
      const int32_t* gp = &a.geom[size_t(t) * kGeomStride];
      nq[0] = gp[j];
      nq[1] = gp[j + 8];
      const int g2 = j < 4 ? gp[j + 16] : -1;
      nq[2] = j < 3 ? g2 : -1;
      const int count = __shfl_sync(gm, g2, 3, 8);
      enablePred = count >= cfg.thr1 && __shfl_sync(gm, nq[0], 0, 8) >= 0;
      if (enablePred) {
        const int parentOnly = cfg.subnode ? 7 : 19;
        int64_t v0[3] = {0, 0, 0};
#pragma unroll
        for (int x = 0; x < 3; x++) {
          const int q = nq[x];
          if (q >= 0) {
            v0[x] = P.rec[size_t(q) * A];
            if (j + 8 * x >= parentOnly && q < p) {
              nocc[x] = P.occ[q];
              nfirst[x] = P.first[q];
            }
          }
        }
        const int64_t self = (int64_t)__shfl_sync(gm, (long long)v0[0], 0, 8);
        const int64_t limLow = 2 * self, limHigh = 25 * self;
#pragma unroll
        for (int x = 0; x < 3; x++) {
          const bool ok = nq[x] >= 0
            && ((x == 0 && j == 0) || (10 * v0[x] > limLow && 10 * v0[x] < limHigh));
          validMask |= ((__ballot_sync(gm, ok) >> gbase) & 0xffu) << (8 * x);
        }
      }
    } else if (root && present) {
      S.nn[cidx] = 19;
    }

    //-- encoder: normalise and transform the sums
    if (enc) {
#pragma unroll
      for (int k = 0; k < 3; k++)
        if (k < A) {
          if (rsMul)
            buf[k] = fx_mul(buf[k] >> rsShift, rsMul);
#pragma unroll
          for (int s = 0; s < 3; s++)
            buf[k] = g_bfly_fwd(gm, buf[k], bf[s], 1 << s, haar);
        }
    }
  }

  // coefficient bookkeeping (set once the prediction is done)
  bool exists = false;
  int ncoef = 0, myPos = 0;
  Quantizer qz[2];
  qz[0].step = qz[1].step = 0;
  qz[0].recip = qz[1].recip = 0;
  unsigned long long codes = 0;  // the block's classification in scan order
  bool hasS = false, hasH = false;
  bool flagMine = false;
  // resolve state
  int m = 0, z = 0, tl = 0;
  bool linked = true, walking = false;
  TzWalk walk;
  walk.req = walk.acc = walk.s = walk.u = 0;
  walk.words = nullptr;
  walk.lists = nullptr;

#ifdef PCCB200_WATCHDOG
  const unsigned long long wdStart = global_ns();
  bool wdReported = false, wdPrinted = false;
#endif

  //==========================================================================
  for (;;) {
    bool progressed = false;
#ifdef PCCB200_WATCHDOG
    {
      // the first group stuck for 3 s opens a 1 s window in which every
      // unfinished group registers; then the lowest tickets report
      const unsigned long long now = global_ns();
      if (now - wdStart > 3000000000ull)
        atomicCAS(&g_wdDeadline, 0ull, now + 1000000000ull);
      const unsigned long long dl = *(volatile unsigned long long*)&g_wdDeadline;
      if (dl) {
        if (state != kStDone && !wdReported)
          atomicMin(&g_wdMinTicket, t);
        wdReported = true;
        if (now > dl) {
          if (!wdPrinted && state != kStDone && t <= *(volatile int*)&g_wdMinTicket + 64) {
            wdPrinted = true;
            if (j == 0)
              printf("stuck stage %d t %d/%d p %d state %d m %d/%d linked %d walking %d "
                     "walk(s %d u %d req %d acc %d) codes %llx hasS %d hasH %d pred %d vm %x "
                     "ticket %llu age %llu ms\n",
                     a.stageIdx, t, *a.count, p, state, m, ncoef, int(linked), int(walking),
                     walk.s, walk.u, walk.req, walk.acc, codes, int(hasS), int(hasH),
                     int(enablePred), validMask, *(volatile unsigned long long*)g_wdTicket,
                     (now - wdStart) / 1000000ull);
            if (missing)
              printf("   t %d lane %d missing rec[%lld] (node %lld)\n", t, j,
                     (long long)(missing - S.rec), (long long)(missing - S.rec) / A);
          }
          __nanosleep(1000000);
          if (now > dl + 500000000ull)
            __trap();
        }
      }
    }
#endif

    //------------------------------------------------------------------------
    // prediction (intraDcPred), classification, publication
    if (state == kStPredict) {
      bool ready = true;
      if (missing) {
        ready = ld_rec(missing) != kRecNotReady;
        if (ready)
          missing = nullptr;
      }
      if (!__ballot_sync(gm, !ready)) {
        bool ok = true;
        if (enablePred) {
          int wsum = -1;
          const int64_t fracMul = ext ? 1 : (int64_t(1) << kFracBits);
#pragma unroll
          for (int k = 0; k < 3; k++)
            pred[k] = 0;
          uint32_t vm = validMask;
          while (vm) {
            const int i = __ffs(vm) - 1;
            vm &= vm - 1;
            const int x = i >> 3;  // group uniform
            const int src = i & 7;
            const int q = __shfl_sync(gm, x == 0 ? nq[0] : x == 1 ? nq[1] : nq[2], src, 8);
            const uint32_t no =
              __shfl_sync(gm, x == 0 ? nocc[0] : x == 1 ? nocc[1] : nocc[2], src, 8);
            const uint32_t mask = uint32_t(neigh_mask(i)) & occ;
            uint32_t cmask = 0;
            int shift = 0;
            if (no) {  // only set for i >= parentOnly && q < p
              const int ii = i - 7;
              const int sh = occu_shift(ii);
              shift = ii < 9 ? sh : -sh;
              cmask = (ii < 9 ? (no >> sh) : (no << sh)) & mask & 0xffu;
            }
            int cfirst = 0;
            if (cmask)
              cfirst =
                __shfl_sync(gm, x == 0 ? nfirst[0] : x == 1 ? nfirst[1] : nfirst[2], src, 8);
            if ((mask >> j) & 1) {
              if ((cmask >> j) & 1) {
                // produced by an earlier block of this stage
                const int wc = cfg.predWeightChild[i - 7];
                const int c = cfirst + __popc(no & ((1u << (j + shift)) - 1));
                wsum += wc;
#pragma unroll
                for (int k = 0; k < 3; k++)
                  if (k < A) {
                    const int64_t* src64 = &S.rec[size_t(c) * A + k];
                    const int64_t v = ld_rec(src64);
                    if (v == kRecNotReady && !missing)
                      missing = src64;
                    pred[k] += v * (wc * fracMul);
                  }
              } else {
                const int wp = cfg.predWeightParent[i];
                wsum += wp;
#pragma unroll
                for (int k = 0; k < 3; k++)
                  if (k < A)
                    pred[k] += P.rec[size_t(q) * A + k] * (wp * fracMul);
              }
            }
          }
          ok = !__ballot_sync(gm, missing != nullptr);
          if (ok) {
            int64_t div = 0, sq = 0;
            if (present) {
              const int d = wsum + 1;
              div = (32768 + d / 2) / d;
              if (!haar && w0 > 1)
                sq = int64_t(isqrt64(uint64_t(w0) << (2 * kFracBits)));
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
              if (k < A) {
                int64_t v = 0;
                if (present) {
                  v = fx_mul(pred[k], div);
                  if (haar)
                    v = (v >> kFracBits) << kFracBits;
                  else if (w0 > 1)
                    v = fx_mul(v, sq);
                }
                pred[k] = v;
#pragma unroll
                for (int s = 0; s < 3; s++)
                  pred[k] = g_bfly_fwd(gm, pred[k], bf[s], 1 << s, haar);
              }
          }
        }
        if (ok) {
          progressed = true;
          //-- coefficients: lane j owns coefficient j (all components)
          exists = j == 0 ? root : wfin != 0;
          const uint32_t existsMask = (__ballot_sync(gm, exists) >> gbase) & 0xffu;
          // bit i' set in before(j): coefficient i' precedes j in scan order 0,4,2,1,6,5,3,7
          const uint32_t before =
            j == 0 ? 0x00u : j == 4 ? 0x01u : j == 2 ? 0x11u : j == 1 ? 0x15u
            : j == 6 ? 0x17u : j == 5 ? 0x57u : j == 3 ? 0x77u : 0x7fu;
          ncoef = __popc(existsMask);
          myPos = __popc(existsMask & before);

          LayerQp lq;
          lq.luma = a.qt->layers[a.qpLayer][0];
          lq.chromaOffset = a.qt->layers[a.qpLayer][1];
          lq.maxQp = cfg.maxQp;
          lq.fixedPointQpOffset = cfg.fixedPointQpOffset;
          {
            // (the encoder never reaches this kernel with AC qp offsets, so the
            // RDOQ test and the quantisation share the quantisers)
            int off0 = nodeQp0, off1 = nodeQp1;
            if (j && a.acLayer < cfg.numAcLayers) {
              off0 += a.qt->acQps[a.acLayer][j - 1][0];
              off1 += a.qt->acQps[a.acLayer][j - 1][1];
            }
            make_quantizers(lq, off0, off1, qz);
          }

          if (enc && enablePred && exists)
#pragma unroll
            for (int k = 0; k < 3; k++)
              if (k < A)
                buf[k] -= pred[k];

          if (rdoq) {
            int code = kCodeZero;
            if (exists) {
              int64_t d2 = 0, aq = 0;
              int rc = 0;
#pragma unroll
              for (int k = 0; k < 3; k++)
                if (k < A) {
                  const int64_t c = fx_round(buf[k]);
                  d2 += c * c;
                  const int64_t qc = qz[k < 1 ? k : 1].quantize(c << kAttrShift);
                  const int64_t mag = qc < 0 ? -qc : qc;
                  aq += mag;
                  rc += lut_log(mag);
                }
              if (aq >= 3) {
                code = kCodeHard;
              } else if (aq > 0) {
                const int64_t l0 = qz[0].scale(1);
                code = rdoq_code(d2, l0 * l0 * (A == 1 ? 25 : 35), rc);
              }
            }
            // the block's coefficients in scan order, replicated in every lane
            codes = exists ? (unsigned long long)code << (6 * myPos) : 0ull;
            codes |= (unsigned long long)g_shfl_xor_i64(gm, (int64_t)codes, 1);
            codes |= (unsigned long long)g_shfl_xor_i64(gm, (int64_t)codes, 2);
            codes |= (unsigned long long)g_shfl_xor_i64(gm, (int64_t)codes, 4);
            hasS = hasH = false;
            int lastH = -1;
            for (int mm = 0; mm < ncoef; mm++) {
              const int cd = int((codes >> (6 * mm)) & 63);
              if (cd == kCodeHard) {
                hasH = true;
                lastH = mm;
              } else if (cd >= 3) {
                hasS = true;
              }
            }
            // publish what is known without looking at any other block
            const TzRegion rg = a.regions[a.stageIdx];
            if (hasH) {
              int e = 0;
              for (int mm = lastH + 1; mm < ncoef; mm++) {
                const int cd = int((codes >> (6 * mm)) & 63);
                e = (cd < 3 || e >= thr_decode(cd)) ? e + 1 : 0;
              }
              if (j == 0)
                st_release(&rg.words[t + 1], tz_pack(kTzExit, e));
            } else if (!hasS) {
              if (j == 0)
                st_release(&rg.words[t + 1], tz_pack(kTzTransparent, ncoef));
            } else if (j == 0) {
              reinterpret_cast<unsigned long long*>(rg.lists)[t + 1] = codes;
              __threadfence();
              st_release(&rg.words[t + 1], tz_pack(kTzClassified, ncoef));
            }
            m = 0;
            z = 0;
            tl = 0;
            linked = true;
            walking = false;
          }
          state = kStResolve;
        }
      }
    }

    //------------------------------------------------------------------------
    // RDOQ decisions of this block, then quantisation and reconstruction
    if (state == kStResolve) {
      bool pending = false;
      if (rdoq) {
        while (m < ncoef) {
          const int cd = int((codes >> (6 * m)) & 63);
          bool f;
          if (cd < 3) {
            f = cd == kCodeRemoved;
          } else {
            const int th = thr_decode(cd);
            if (linked) {
              int r = 1;
              if (!walking && th - z > 0) {
                walk.req = th - z;
                walk.acc = 0;
                tz_walk_enter(a, walk, a.stageIdx);
                walk.u = t - 1;
                walking = true;
              }
              if (walking) {
                r = a.experiment == 1 ? 0 : tz_walk(a, walk, gm);
                if (r < 0) {
                  pending = true;
                  break;
                }
                walking = false;
              }
              f = r != 0;
            } else {
              f = tl >= th;
            }
          }
          const bool keeps = cd < 2 || f;
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
          m++;
          progressed = true;
        }
        if (!pending && hasS && !hasH && j == 0)
          st_release(&a.regions[a.stageIdx].words[t + 1],
                     linked ? tz_pack(kTzTransparent, ncoef) : tz_pack(kTzExit, tl));
      }
      if (!pending) {
        progressed = true;
        //-- quantise / dequantise (RAHT.cpp:1672-1723)
        if (exists) {
          const int64_t pos = a.coefBase + c0 - (root ? 0 : p) + myPos;
#pragma unroll
          for (int k = 0; k < 3; k++)
            if (k < A) {
              const Quantizer& qk = qz[k < 1 ? k : 1];
              int64_t qc;
              if (enc) {
                const int64_t c = flagMine ? 0 : fx_round(buf[k]);
                qc = qk.quantize(c << kAttrShift);
                a.coef[k * a.coefStride + pos] = int32_t(qc);
              } else {
                qc = a.coef[k * a.coefStride + pos];
              }
              pred[k] += fx_from_int(div_exp2_round_half_up(qk.scale(qc), kAttrShift));
            }
        }
        //-- DC from the parent, inverse transform, store (RAHT.cpp:1726-1806)
#pragma unroll
        for (int k = 0; k < 3; k++)
          if (k < A) {
            if (!root && j == 0) {
              const int64_t v = P.recUs[size_t(p) * A + k];
              pred[k] = ext ? v : v * (int64_t(1) << (kFracBits - 2));
            }
#pragma unroll
            for (int s = 2; s >= 0; s--)
              pred[k] = g_bfly_inv(gm, pred[k], bf[s], 1 << s, haar);
            if (present) {
              int64_t v = pred[k];
              S.recUs[size_t(cidx) * A + k] = ext ? v : fx_round(v * 4);
              if (rsMul)
                v = fx_mul(v >> rsShift, rsMul);
              st_rec(&S.rec[size_t(cidx) * A + k], ext ? v : fx_round(v));
            }
          }
        state = kStDone;
      }
    }

    if (!__ballot_sync(0xffffffffu, state != kStDone))
      break;
    if (!__any_sync(0xffffffffu, progressed))
      __nanosleep(32);
  }
}

__global__ void __launch_bounds__(kWarpBlockThreads)
k_block_warp8(const WarpBlockArgs a, unsigned long long* ticket)
{
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3;
  const int n = *a.count;
#ifdef PCCB200_WATCHDOG
  g_wdTicket = ticket;
#endif
  constexpr int kTile = kGroupsPerWarp * kTicketInterleave;
  for (;;) {
    unsigned long long c = 0;
    if (lane == 0)
      c = atomicAdd(ticket, 1ull);
    c = __shfl_sync(0xffffffffu, c, 0);
    const long long tile = (long long)(c / kTicketInterleave) * kTile;
    if (tile >= n)
      return;
    const int t = int(tile) + g * kTicketInterleave + int(c % kTicketInterleave);
    group_blocks(a, t, t < n, lane);
  }
}

}  // namespace pccb200
