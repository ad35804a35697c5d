/* raht_oracle.c — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Sequential plain-C restatement of TMC13's region-adaptive hierarchical
 * transform (the "upsampled / predictive" RAHT of release-23.0-rc2), intra
 * mode, forward (encoder) and inverse (decoder):
 *
 *   pcc::regionAdaptiveHierarchicalTransform / ...InverseTransform
 *       tmc3/RAHT.cpp:1997-2058 -> uraht_process<isEncoder, rahtExtension>
 *       tmc3/RAHT.cpp:977-1976
 *
 * The restatement is organised differently from the reference: instead of
 * the LF/HF vectors that are reduced and re-expanded one binary level at a
 * time (RAHT.cpp:108-264), every transform stage (each 3rd binary level that
 * adds nodes, RAHT.cpp:1205-1209) is materialised once as an array of nodes in
 * Morton order; the descent then walks blocks of siblings.  The results are
 * required to be bit-identical to the reference.  Parity is PINNED: checked
 * against the compiled reference (oracle/_ref) and the committed golden
 * vectors by tests/test_oracle_vs_reference.py.
 *
 * Not covered (out of scope): inter-frame prediction
 * (enableAttrInterPred), RAHT.cpp:805-972 and the *_ref branches.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pcc_attr_b200.h"
#include "pcc_arith_oracle.h"

typedef struct {
  int level;      /* bit level of the nodes: unique in (key >> level) */
  int n;
  int64_t* key;   /* Morton code of the node's first leaf (UrahtNode::pos) */
  int32_t* weight;
  int32_t* attr;  /* n*A: attribute sums, or Haar low-pass (attrsLf) */
  int32_t* qp;    /* n*2: region qp offset << 4, averaged over the node's
                     subtree on the way up (UrahtNode::qp after reduceLevel) */
  int32_t* qpd;   /* n*2: the qp the node carries on the way down: expandLevel
                     (RAHT.cpp:210-264) restores weights and sums but leaves
                     the kept ("left") node with the averaged qp */
  int32_t* first; /* n+1: first child in the next finer stage */
  int32_t* nn;    /* numParentNeigh of the node */
  uint8_t* occ;   /* child occupancy (UrahtNode::occupancy) */
  int64_t* rec;   /* n*A attrRec: reconstruction scaled by 1/sqrt(w) */
  int64_t* recus; /* n*A attrRecUs: un-scaled reconstruction */
} stage_t;

typedef struct {
  const pccb200_raht_params* pp;
  const pccb200_qpset* qs;
  int A;
  int isEncoder;
  int ext;
  int haar;
  int trainZeros;      /* RDOQ zero-run state, RAHT.cpp:1154 */
  int32_t* coef[3];    /* write / read cursors of the planar buffers */
} ctx_t;

static void
stage_alloc(stage_t* s, int n, int A)
{
  s->n = n;
  s->key = (int64_t*)calloc((size_t)n, sizeof(int64_t));
  s->weight = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  s->attr = (int32_t*)calloc((size_t)n * A, sizeof(int32_t));
  s->qp = (int32_t*)calloc((size_t)n * 2, sizeof(int32_t));
  s->qpd = (int32_t*)calloc((size_t)n * 2, sizeof(int32_t));
  s->first = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
  s->nn = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  s->occ = (uint8_t*)calloc((size_t)n, 1);
  s->rec = (int64_t*)calloc((size_t)n * A, sizeof(int64_t));
  s->recus = (int64_t*)calloc((size_t)n * A, sizeof(int64_t));
}

static void
stage_free(stage_t* s)
{
  free(s->key);
  free(s->weight);
  free(s->attr);
  free(s->qp);
  free(s->qpd);
  free(s->first);
  free(s->nn);
  free(s->occ);
  free(s->rec);
  free(s->recus);
}

/* QpSet::quantizers(qpLayer, qpOffset), quantization.cpp:169-178 */
static void
mk_quantizers(const pccb200_qpset* qs, int layer, int off0, int off1,
              orc_quantizer q[2])
{
  int qp0 = qs->layers[layer][0] + off0;
  if (qp0 < 4) qp0 = 4;
  if (qp0 > qs->max_qp) qp0 = qs->max_qp;
  int qp1 = qs->layers[layer][1] + off1 + qp0;
  if (qp1 < 4) qp1 = 4;
  if (qp1 > qs->max_qp) qp1 = qs->max_qp;
  q[0] = orc_mkquant(qp0 + qs->fixed_point_qp_offset);
  q[1] = orc_mkquant(qp1 + qs->fixed_point_qp_offset);
}

/* RahtKernel constructor, RAHT.cpp:596-604 */
static void
raht_ab(int wl, int wr, int64_t* a, int64_t* b)
{
  uint64_t w = (uint64_t)wl + (uint64_t)wr;
  uint64_t isw = orc_irsqrt(w);
  *a = (int64_t)((orc_isqrt((uint64_t)wl << 30) * isw) >> 40);
  *b = (int64_t)((orc_isqrt((uint64_t)wr << 30) * isw) >> 40);
}

static const int kPairA[12] = {0, 2, 4, 6, 0, 4, 1, 5, 0, 1, 2, 3};
static const int kPairB[12] = {1, 3, 5, 7, 2, 6, 3, 7, 4, 5, 6, 7};

/* mkWeightTree, RAHT.cpp:742-771: weights[0..7] children; [8..15], [16..23],
 * [24..31] the pair sums after each of the three stages, the upper four of
 * each group being the high-pass weights (0 when the pair was not a pair). */
static void
mk_weight_tree(int w[32])
{
  for (int g = 0; g < 3; g++) {
    int* in = w + 8 * g;
    int* out = w + 8 * (g + 1);
    for (int i = 0; i < 4; i++) {
      int s = in[2 * i] + in[2 * i + 1];
      out[i] = s;
      out[4 + i] = (in[2 * i] && in[2 * i + 1]) ? s : 0;
    }
  }
}

/* fwdTransformBlock222 / invTransformBlock222, RAHT.cpp:671-737, with the
 * RahtKernel (RAHT.cpp:606-640) or HaarKernel (RAHT.cpp:653-665) butterfly. */
static void
fwd_block(int nbuf, int64_t buf[][8], const int w[32], int haar)
{
  for (int i = 0, iw = 0; i < 12; i++, iw += 2) {
    int i0 = kPairA[i], i1 = kPairB[i];
    if (w[iw] + w[iw + 1] == 0)
      continue;
    if (!w[iw] || !w[iw + 1]) {
      if (!w[iw])
        for (int k = 0; k < nbuf; k++) {
          int64_t t = buf[k][i0];
          buf[k][i0] = buf[k][i1];
          buf[k][i1] = t;
        }
      continue;
    }
    if (haar) {
      for (int k = 0; k < nbuf; k++) {
        int64_t l = buf[k][i0], r = buf[k][i1];
        int64_t hf = r - l;
        buf[k][i0] = l + ((hf >> (1 + ORC_FRAC_BITS)) << ORC_FRAC_BITS);
        buf[k][i1] = hf;
      }
      continue;
    }
    int64_t a, b;
    raht_ab(w[iw], w[iw + 1], &a, &b);
    for (int k = 0; k < nbuf; k++) {
      int64_t l = buf[k][i0], r = buf[k][i1];
      buf[k][i0] = orc_fxmul(r, b) + orc_fxmul(a, l);
      buf[k][i1] = orc_fxmul(r, a) - orc_fxmul(b, l);
    }
  }
}

static void
inv_block(int nbuf, int64_t buf[][8], const int w[32], int haar)
{
  for (int i = 11, iw = 22; i >= 0; i--, iw -= 2) {
    int i0 = kPairA[i], i1 = kPairB[i];
    if (w[iw] + w[iw + 1] == 0)
      continue;
    if (!w[iw] || !w[iw + 1]) {
      if (!w[iw])
        for (int k = 0; k < nbuf; k++) {
          int64_t t = buf[k][i0];
          buf[k][i0] = buf[k][i1];
          buf[k][i1] = t;
        }
      continue;
    }
    if (haar) {
      for (int k = 0; k < nbuf; k++) {
        int64_t lf = buf[k][i0], hf = buf[k][i1];
        int64_t l = lf - ((hf >> (1 + ORC_FRAC_BITS)) << ORC_FRAC_BITS);
        buf[k][i0] = l;
        buf[k][i1] = hf + l;
      }
      continue;
    }
    int64_t a, b;
    raht_ab(w[iw], w[iw + 1], &a, &b);
    for (int k = 0; k < nbuf; k++) {
      int64_t lf = buf[k][i0], hf = buf[k][i1];
      buf[k][i0] = orc_fxmul(lf, a) - orc_fxmul(b, hf);
      buf[k][i1] = orc_fxmul(lf, b) + orc_fxmul(a, hf);
    }
  }
}

/* scale an (un-normalised) value by 1/sqrt(w), RAHT.cpp:1474-1481,1780-1787 */
static int64_t
scale_rsqrt(int64_t v, int w)
{
  int shift = w > 1024 ? orc_ilog2((uint64_t)(w - 1)) >> 1 : 0;
  int64_t rs = (int64_t)(orc_irsqrt((uint64_t)w) >> (40 - shift - ORC_FRAC_BITS));
  return orc_fxmul(v >> shift, rs);
}

/* index of the node with (key >> level) == pos in [0, n), or -1 */
static int
find_pos(const stage_t* P, int level, int64_t pos)
{
  int lo = 0, hi = P->n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((P->key[mid] >> level) < pos)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (lo < P->n && (P->key[lo] >> level) == pos)
    return lo;
  return -1;
}

static const uint8_t kNeighMasks[19] = {255, 240, 204, 170, 192, 160, 136,
                                        3,   5,   15,  17,  51,  85,  10,
                                        34,  12,  68,  48,  80};
static const uint8_t kNeighOffset[19] = {0, 35, 21, 14, 49, 42, 28, 1,  2, 3,
                                         4, 5,  6,  10, 12, 17, 20, 33, 34};
static const uint8_t kOccuShift[12] = {6, 5, 4, 3, 2, 1, 3, 1, 2, 1, 2, 3};

/* findNeighbours, RAHT.cpp:299-416.  P = parent stage, p = the parent,
 * S = child stage.  A neighbour counts as already visited at this depth iff
 * it precedes p in Morton order (its occupancy has been set, RAHT.cpp:1393). */
static void
find_neighbours(const stage_t* P, const stage_t* S, int p, int plevel,
                uint8_t occ, int subnode, int range, int pidx[19],
                int cidx[12][8])
{
  int64_t cur = P->key[p] >> plevel;
  int64_t base = (int64_t)orc_morton3d_add((uint64_t)cur, (uint64_t)-1ll);
  pidx[0] = p;
  for (int i = 1; i < 19; i++) {
    pidx[i] = -1;
    if (!(occ & kNeighMasks[i]))
      continue;
    int64_t np = (int64_t)orc_morton3d_add((uint64_t)base, kNeighOffset[i]);
    int q = find_pos(P, plevel, np);
    if (q < 0)
      continue;
    int d = q - p;
    if (d < 0)
      d = -d;
    if (d > range)
      continue;
    pidx[i] = q;
  }
  if (!subnode)
    return;
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 8; j++)
      cidx[i][j] = -1;
  for (int i = 0; i < 12; i++) {
    int q = pidx[7 + i];
    if (q < 0 || q >= p)
      continue; /* not found, or not yet visited: occupancy still 0 */
    uint8_t nocc = P->occ[q];
    int sh = kOccuShift[i];
    uint8_t mask = (uint8_t)((i < 9 ? (nocc >> sh) : (nocc << sh)) & occ
                             & kNeighMasks[7 + i]);
    if (!mask)
      continue;
    for (int c = P->first[q]; c < P->first[q + 1]; c++) {
      int slot = (int)((S->key[c] >> S->level) & 7);
      int j = i < 9 ? slot - sh : slot + sh;
      if (j >= 0 && j < 8 && ((mask >> j) & 1))
        cidx[i][j] = c;
    }
  }
}

/* intraDcPred, RAHT.cpp:421-589 */
static void
intra_dc_pred(const ctx_t* cx, const stage_t* P, const stage_t* S,
              const int pidx[19], int cidx[12][8], int occ, int64_t pred[][8])
{
  const pccb200_raht_params* pp = cx->pp;
  const int A = cx->A;
  int wsum[8];
  for (int j = 0; j < 8; j++)
    wsum[j] = -1;
  int64_t limLow = 0, limHigh = 0;
  const int fracMul = cx->ext ? 1 : (1 << ORC_FRAC_BITS);

  int parentOnly = pp->subnode_prediction_enabled ? 7 : 19;
  for (int i = 0; i < parentOnly; i++) {
    if (pidx[i] < 0)
      continue;
    int64_t v[3];
    for (int k = 0; k < A; k++)
      v[k] = P->rec[(size_t)pidx[i] * A + k];
    if (i) {
      if (10 * v[0] <= limLow || 10 * v[0] >= limHigh)
        continue;
    } else {
      limLow = 2 * v[0];
      limHigh = 25 * v[0];
    }
    int w = pp->pred_weight_parent[i];
    for (int k = 0; k < A; k++)
      v[k] *= (int64_t)w * fracMul;
    int mask = kNeighMasks[i] & occ;
    for (int j = 0; mask; j++, mask >>= 1)
      if (mask & 1) {
        wsum[j] += w;
        for (int k = 0; k < A; k++)
          pred[k][j] += v[k];
      }
  }
  if (pp->subnode_prediction_enabled) {
    for (int i = 0; i < 12; i++) {
      if (pidx[7 + i] < 0)
        continue;
      int64_t v[3] = {0, 0, 0};
      for (int k = 0; k < A; k++)
        v[k] = P->rec[(size_t)pidx[7 + i] * A + k];
      if (10 * v[0] <= limLow || 10 * v[0] >= limHigh)
        continue;
      int wp = pp->pred_weight_parent[7 + i];
      int wc = pp->pred_weight_child[i];
      for (int k = 0; k < A; k++)
        v[k] *= (int64_t)wp * fracMul;
      int mask = kNeighMasks[7 + i] & occ;
      for (int j = 0; mask; j++, mask >>= 1) {
        if (!(mask & 1))
          continue;
        if (cidx[i][j] >= 0) {
          wsum[j] += wc;
          for (int k = 0; k < A; k++)
            pred[k][j] +=
              S->rec[(size_t)cidx[i][j] * A + k] * ((int64_t)wc * fracMul);
        } else {
          wsum[j] += wp;
          for (int k = 0; k < A; k++)
            pred[k][j] += v[k];
        }
      }
    }
  }
  /* normalise by the Q.15 reciprocal of the weight sum (kDivisors) */
  for (int j = 0; j < 8; j++) {
    if (!((occ >> j) & 1))
      continue;
    int d = wsum[j] + 1;
    int64_t div = (32768 + d / 2) / d; /* == kDivisors[wsum], RAHT.cpp:445-451 */
    for (int k = 0; k < A; k++) {
      pred[k][j] = orc_fxmul(pred[k][j], div);
      if (cx->haar)
        pred[k][j] = (pred[k][j] >> ORC_FRAC_BITS) << ORC_FRAC_BITS;
    }
  }
}

/* RDOQ rate of a zero run, RAHT.cpp:1578-1634 */
static const int kLUTlog[16] = {0,   256, 406, 512, 594, 662, 719,  768,
                                812, 850, 886, 918, 947, 975, 1000, 1024};
static const int kLUTbins[11] = {1, 2, 3, 5, 5, 7, 7, 9, 9, 11, 11};

static int
zero_run_rate(int tz)
{
  int rate = kLUTbins[tz > 10 ? 10 : tz];
  if (tz > 10) {
    int t = tz - 11 + 1, a = 0;
    while (t) {
      a++;
      t >>= 1;
    }
    rate += 2 * a - 1 + 2;
  }
  return rate;
}

static const int kScan[8] = {0, 4, 2, 1, 6, 5, 3, 7};

/* Region-qp carried by each child of a block during the descent.  Going up,
 * reduceLevel averages the qps of a merged pair into the kept (left) node
 * (RAHT.cpp:187-188) and pushes the right node, with the qp of its own
 * subtree, to the high-pass list; going down, expandLevel re-creates the
 * pair but never undoes the average.  Hence the first child of a block
 * inherits the parent's descent qp, and the first node of every right-hand
 * subtree of the block's binary merge tree carries that subtree's average. */
static void
descent_qps(stage_t* S, int c0, int c1, const int32_t parentQp[2])
{
  int32_t up[8][2];
  int present[8] = {0};
  int child[8];
  memset(up, 0, sizeof(up));
  for (int c = c0; c < c1; c++) {
    int slot = (int)((S->key[c] >> S->level) & 7);
    present[slot] = 1;
    child[slot] = c;
    up[slot][0] = S->qp[2 * c];
    up[slot][1] = S->qp[2 * c + 1];
  }
  /* first child of the block */
  S->qpd[2 * c0] = parentQp[0];
  S->qpd[2 * c0 + 1] = parentQp[1];
  int first[8], has[8];
  for (int j = 0; j < 8; j++) {
    first[j] = j;
    has[j] = present[j];
  }
  /* sizes 1, 2, 4: when two non-empty halves meet, the right half's first
   * child takes the right half's accumulated average */
  for (int step = 1; step < 8; step <<= 1) {
    for (int lo = 0; lo < 8; lo += 2 * step) {
      int hi = lo + step;
      if (!has[hi])
        continue;
      if (!has[lo]) {
        has[lo] = 1;
        first[lo] = first[hi];
        up[lo][0] = up[hi][0];
        up[lo][1] = up[hi][1];
        continue;
      }
      int c = child[first[hi]];
      S->qpd[2 * c] = up[hi][0];
      S->qpd[2 * c + 1] = up[hi][1];
      up[lo][0] = (up[lo][0] + up[hi][0]) >> 1;
      up[lo][1] = (up[lo][1] + up[hi][1]) >> 1;
    }
  }
}

/* One block of siblings: the body of the loop RAHT.cpp:1306-1808.
 * P == NULL for the root block (inheritDc == false). */
static void
process_block(ctx_t* cx, stage_t* S, stage_t* P, int p, int c0, int c1,
              int predInLvl, int qpLayer, int acLayer)
{
  const int A = cx->A;
  const pccb200_raht_params* pp = cx->pp;
  const int inheritDc = P != NULL;
  int64_t buf[6][8];
  int w[32];
  int nodeQp[8][2];
  uint8_t occ = 0;
  memset(buf, 0, sizeof(buf));
  memset(w, 0, sizeof(w));
  memset(nodeQp, 0, sizeof(nodeQp));
  int64_t(*pred)[8] = &buf[A];

  for (int c = c0; c < c1; c++) {
    int slot = (int)((S->key[c] >> S->level) & 7);
    w[slot] = S->weight[c];
    nodeQp[slot][0] = S->qpd[2 * c] >> 4;
    nodeQp[slot][1] = S->qpd[2 * c + 1] >> 4;
    occ |= (uint8_t)(1 << slot);
    if (cx->isEncoder)
      for (int k = 0; k < A; k++)
        buf[k][slot] = orc_fxfromint(S->attr[(size_t)c * A + k]);
  }
  int nodeCnt = cx->ext ? c1 - c0 : 0;
  mk_weight_tree(w);

  if (!inheritDc)
    for (int c = c0; c < c1; c++)
      S->nn[c] = 19;

  int enablePred = predInLvl;
  if (predInLvl) {
    P->occ[p] = occ;
    int pidx[19], cidx[12][8];
    int count = 0;
    if (cx->ext && nodeCnt == 1) {
      enablePred = 0;
      count = 19;
    } else if (P->nn[p] < pp->prediction_threshold0) {
      enablePred = 0;
    } else {
      find_neighbours(P, S, p, S->level + 3, occ,
                      pp->subnode_prediction_enabled,
                      pp->prediction_search_range, pidx, cidx);
      for (int i = 0; i < 19; i++)
        count += pidx[i] != -1;
      if (count < pp->prediction_threshold1)
        enablePred = 0;
      else
        intra_dc_pred(cx, P, S, pidx, cidx, occ, pred);
    }
    for (int c = c0; c < c1; c++)
      S->nn[c] = count;
  }

  /* normalise: sums by 1/sqrt(w), predictions by sqrt(w), RAHT.cpp:1445-1499 */
  if (!cx->haar) {
    for (int j = 0; j < 8; j++) {
      if (w[j] <= 1)
        continue;
      if (cx->isEncoder)
        for (int k = 0; k < A; k++)
          buf[k][j] = scale_rsqrt(buf[k][j], w[j]);
      if (enablePred) {
        int64_t sq = orc_isqrt((uint64_t)w[j] << (2 * ORC_FRAC_BITS));
        for (int k = 0; k < A; k++)
          pred[k][j] = orc_fxmul(pred[k][j], sq);
      }
    }
  }

  /* forward transform of sums and of the prediction, RAHT.cpp:1504-1549 */
  if (cx->isEncoder && enablePred)
    fwd_block(2 * A, buf, w, cx->haar);
  else if (cx->isEncoder)
    fwd_block(A, buf, w, cx->haar);
  else if (enablePred)
    fwd_block(A, pred, w, cx->haar);

  /* per coefficient, in scan order (scanBlock, RAHT.cpp:776-791) */
  const pccb200_qpset* qs = cx->qs;
  for (int si = 0; si < 8; si++) {
    int idx = kScan[si];
    if (si && !w[24 + idx])
      continue;
    if (inheritDc && !idx)
      continue;

    if (cx->isEncoder && enablePred)
      for (int k = 0; k < A; k++)
        buf[k][idx] -= pred[k][idx];

    /* RDOQ decision, RAHT.cpp:1576-1670 */
    int flagRDOQ = 0;
    if (cx->isEncoder && !cx->haar) {
      int64_t sumCoeff = 0, dist2 = 0, lambda0 = 0;
      int rateCoeff = 0;
      orc_quantizer q[2];
      mk_quantizers(qs, qpLayer, nodeQp[idx][0], nodeQp[idx][1], q);
      for (int k = 0; k < A; k++) {
        orc_quantizer qk = q[k < 1 ? k : 1];
        int64_t c = orc_fxround(buf[k][idx]);
        dist2 += c * c;
        int64_t qc = orc_quantize(qk, c << ORC_ATTR_SHIFT);
        int64_t aq = qc < 0 ? -qc : qc;
        sumCoeff += aq;
        rateCoeff += kLUTlog[aq < 15 ? aq : 15];
        if (!k)
          lambda0 = orc_scale(qk, 1);
      }
      int64_t lambda = lambda0 * lambda0 * (A == 1 ? 25 : 35);
      if (sumCoeff < 3) {
        int rate = zero_run_rate(cx->trainZeros) + ((rateCoeff + 128) >> 8);
        flagRDOQ = (dist2 << 26) < lambda * rate;
      }
      if (flagRDOQ || sumCoeff == 0)
        cx->trainZeros++;
      else
        cx->trainZeros = 0;
    }

    /* quantiser for this coefficient, RAHT.cpp:1672-1682 */
    int off0 = nodeQp[idx][0], off1 = nodeQp[idx][1];
    if (idx && acLayer < qs->num_ac_coeff_qp_layers) {
      off0 += qs->ac_coeff_qps[acLayer][idx - 1][0];
      off1 += qs->ac_coeff_qps[acLayer][idx - 1][1];
    }
    orc_quantizer q[2];
    mk_quantizers(qs, qpLayer, off0, off1, q);
    for (int k = 0; k < A; k++) {
      orc_quantizer qk = q[k < 1 ? k : 1];
      int64_t qc;
      if (cx->isEncoder) {
        if (flagRDOQ)
          buf[k][idx] = 0;
        int64_t c = orc_fxround(buf[k][idx]);
        qc = orc_quantize(qk, c << ORC_ATTR_SHIFT);
        *cx->coef[k]++ = (int32_t)qc;
      } else {
        qc = *cx->coef[k]++;
      }
      pred[k][idx] += orc_fxfromint(
        orc_div_exp2_half_up(orc_scale(qk, qc), ORC_ATTR_SHIFT));
    }
  }

  /* DC comes from the parent's un-scaled reconstruction, RAHT.cpp:1726-1742 */
  if (inheritDc)
    for (int k = 0; k < A; k++) {
      int64_t v = P->recus[(size_t)p * A + k];
      if (cx->ext)
        pred[k][0] = v;
      else if (v > 0)
        pred[k][0] = v << (ORC_FRAC_BITS - 2);
      else
        pred[k][0] = -((-v) << (ORC_FRAC_BITS - 2));
    }

  inv_block(A, pred, w, cx->haar);

  /* store reconstructions, RAHT.cpp:1754-1806 */
  for (int c = c0; c < c1; c++) {
    int slot = (int)((S->key[c] >> S->level) & 7);
    for (int k = 0; k < A; k++) {
      int64_t v = pred[k][slot];
      S->recus[(size_t)c * A + k] = cx->ext ? v : orc_fxround(v << 2);
      if (!cx->haar && w[slot] > 1)
        v = scale_rsqrt(v, w[slot]);
      S->rec[(size_t)c * A + k] = cx->ext ? v : orc_fxround(v);
    }
  }
}

/* merge the per-slot values of one block into its parent node through the
 * three binary levels (reduceLevel, RAHT.cpp:157-205): weights and sums add,
 * region qps average pairwise, Haar low-pass lifts pairwise. */
static void
merge_children(const stage_t* S, int c0, int c1, int A, int haar,
               int32_t* weightOut, int32_t* attrOut, int32_t* qpOut)
{
  int present[8] = {0};
  int32_t qp[8][2];
  uint32_t at[8][3];
  memset(qp, 0, sizeof(qp));
  memset(at, 0, sizeof(at));
  uint32_t wsum = 0;
  for (int c = c0; c < c1; c++) {
    int slot = (int)((S->key[c] >> S->level) & 7);
    present[slot] = 1;
    qp[slot][0] = S->qp[2 * c];
    qp[slot][1] = S->qp[2 * c + 1];
    for (int k = 0; k < A; k++)
      at[slot][k] = (uint32_t)S->attr[(size_t)c * A + k];
    wsum += (uint32_t)S->weight[c];
  }
  for (int step = 1; step < 8; step <<= 1) {
    for (int lo = 0; lo < 8; lo += 2 * step) {
      int hi = lo + step;
      if (!present[hi])
        continue;
      if (!present[lo]) {
        present[lo] = 1;
        present[hi] = 0;
        qp[lo][0] = qp[hi][0];
        qp[lo][1] = qp[hi][1];
        for (int k = 0; k < A; k++)
          at[lo][k] = at[hi][k];
        continue;
      }
      present[hi] = 0;
      qp[lo][0] = (qp[lo][0] + qp[hi][0]) >> 1;
      qp[lo][1] = (qp[lo][1] + qp[hi][1]) >> 1;
      for (int k = 0; k < A; k++) {
        if (haar) {
          int32_t d = (int32_t)(at[hi][k] - at[lo][k]);
          at[lo][k] += (uint32_t)(d >> 1);
        } else {
          at[lo][k] += at[hi][k]; /* int32 wrap-around as in the reference */
        }
      }
    }
  }
  *weightOut = (int32_t)wsum;
  qpOut[0] = qp[0][0];
  qpOut[1] = qp[0][1];
  for (int k = 0; k < A; k++)
    attrOut[k] = (int32_t)at[0][k];
}

static int
bitlen64(uint64_t x)
{
  int n = 0;
  while (x) {
    n++;
    x >>= 1;
  }
  return n;
}

/* The whole transform.  Returns 0, or -1 on bad arguments. */
int
oracle_raht(int forward, const pccb200_raht_params* pp,
            const pccb200_qpset* qs, const int32_t* pointQpOffsets,
            const int64_t* morton, int32_t* attrs, int A, int N,
            int32_t* coeffs)
{
  if (N <= 0 || A < 1 || A > 3)
    return -1;
  ctx_t cx;
  cx.pp = pp;
  cx.qs = qs;
  cx.A = A;
  cx.isEncoder = forward != 0;
  cx.ext = pp->raht_extension != 0;
  cx.haar = pp->integer_haar != 0;
  cx.trainZeros = 0;
  for (int k = 0; k < 3; k++)
    cx.coef[k] = coeffs + (size_t)N * k;

  /* single point, RAHT.cpp:998-1017 */
  if (N == 1) {
    orc_quantizer q[2];
    int o0 = pointQpOffsets ? pointQpOffsets[0] : 0;
    int o1 = pointQpOffsets ? pointQpOffsets[1] : 0;
    mk_quantizers(qs, 0, o0, o1, q);
    for (int k = 0; k < A; k++) {
      orc_quantizer qk = q[k < 1 ? k : 1];
      int64_t c;
      if (cx.isEncoder) {
        c = orc_quantize(qk, (int64_t)attrs[k] << ORC_ATTR_SHIFT);
        coeffs[k * (size_t)N] = (int32_t)c;
      } else {
        c = coeffs[k * (size_t)N];
      }
      attrs[k] =
        (int32_t)orc_div_exp2_half_up(orc_scale(qk, c), ORC_ATTR_SHIFT);
    }
    return 0;
  }

  /* ---- leaves: merge duplicate positions (reduceUnique, RAHT.cpp:108-152) */
  int nLeaves = 0;
  for (int i = 0; i < N; i++)
    if (i == 0 || morton[i] != morton[i - 1])
      nLeaves++;
  stage_t* stages = (stage_t*)calloc(32, sizeof(stage_t));
  int nStages = 0;
  stage_t* L = &stages[nStages++];
  stage_alloc(L, nLeaves, A);
  /* high-pass values of the duplicates, in point order (attrsHf level 0) */
  int32_t* dupHf = (int32_t*)calloc((size_t)N * A + 1, sizeof(int32_t));
  {
    int u = -1;
    for (int i = 0; i < N; i++) {
      if (i == 0 || morton[i] != morton[i - 1]) {
        u++;
        L->key[u] = morton[i];
        L->weight[u] = 1;
        L->first[u] = i;
        L->qp[2 * u] = (pointQpOffsets ? pointQpOffsets[2 * i] : 0) << 4;
        L->qp[2 * u + 1] = (pointQpOffsets ? pointQpOffsets[2 * i + 1] : 0) << 4;
        for (int k = 0; k < A; k++)
          L->attr[(size_t)u * A + k] = attrs[(size_t)i * A + k];
        continue;
      }
      L->weight[u]++;
      for (int k = 0; k < A; k++) {
        uint32_t lf = (uint32_t)L->attr[(size_t)u * A + k];
        uint32_t in = (uint32_t)attrs[(size_t)i * A + k];
        if (cx.haar) {
          int32_t d = (int32_t)(in - lf);
          dupHf[(size_t)i * A + k] = d;
          lf += (uint32_t)(d >> 1);
        } else {
          dupHf[(size_t)i * A + k] = (int32_t)in;
          lf += in;
        }
        L->attr[(size_t)u * A + k] = (int32_t)lf;
      }
    }
    L->first[nLeaves] = N;
  }
  const int numDup = N - nLeaves;

  /* ---- stage levels: every 3rd binary level below the first level at which
   * all points agree, skipping levels that add no nodes (RAHT.cpp:1086,
   * 1205-1209). cnt[s/3] = number of distinct (key >> s). */
  int lmax = nLeaves > 1 ? bitlen64((uint64_t)(morton[0] ^ morton[N - 1])) : 0;
  int rootLevel = lmax > 0 ? 3 * ((lmax - 1) / 3) : -1;
  int cnt[24];
  for (int s = 0; s <= rootLevel + 3; s += 3) {
    int c = 1;
    for (int u = 1; u < nLeaves; u++)
      if ((L->key[u] >> s) != (L->key[u - 1] >> s))
        c++;
    cnt[s / 3] = c;
  }

  int qpLayer = 0;
  if (rootLevel >= 0) {
    /* finest processed level: the leaves live there */
    int s = 0;
    while (s < rootLevel && cnt[s / 3] == cnt[s / 3 + 1])
      s += 3;
    L->level = s;
    /* build coarser stages */
    while (stages[nStages - 1].level < rootLevel) {
      stage_t* F = &stages[nStages - 1];
      int sp = F->level + 3;
      while (sp < rootLevel && cnt[sp / 3] == cnt[sp / 3 + 1])
        sp += 3;
      stage_t* C = &stages[nStages++];
      stage_alloc(C, cnt[sp / 3], A);
      C->level = sp;
      int u = -1;
      for (int c = 0; c < F->n; c++) {
        if (c == 0 || (F->key[c] >> (F->level + 3)) != (F->key[c - 1] >> (F->level + 3))) {
          u++;
          C->key[u] = F->key[c];
          C->first[u] = c;
        }
      }
      C->first[C->n] = F->n;
      for (u = 0; u < C->n; u++)
        merge_children(F, C->first[u], C->first[u + 1], A, cx.haar,
                       &C->weight[u], &C->attr[(size_t)u * A], &C->qp[2 * u]);
    }

    /* ---- descent: coarse to fine */
    int acLayer = -1;
    for (int si = nStages - 1; si >= 0; si--) {
      stage_t* S = &stages[si];
      stage_t* P = si == nStages - 1 ? NULL : &stages[si + 1];
      qpLayer = qpLayer + 1 < qs->num_layers ? qpLayer + 1 : qs->num_layers - 1;
      acLayer++;
      if (!P) {
        int32_t w0, at0[3], rootQp[2];
        merge_children(S, 0, S->n, A, cx.haar, &w0, at0, rootQp);
        descent_qps(S, 0, S->n, rootQp);
        process_block(&cx, S, NULL, 0, 0, S->n, 0, qpLayer, acLayer);
        continue;
      }
      int predInLvl = pp->prediction_enabled != 0;
      if (predInLvl)
        memset(P->occ, 0, (size_t)P->n);
      for (int p = 0; p < P->n; p++) {
        descent_qps(S, P->first[p], P->first[p + 1], &P->qpd[2 * p]);
        process_block(&cx, S, P, p, P->first[p], P->first[p + 1], predInLvl,
                      qpLayer, acLayer);
      }
    }
  }

  /* ---- duplicate points, RAHT.cpp:1840-1964, and write-back :1967-1975 */
  int64_t* out = (int64_t*)calloc((size_t)N * A, sizeof(int64_t));
  if (!numDup) {
    memcpy(out, L->rec, sizeof(int64_t) * (size_t)N * A);
  } else {
    for (int u = 0; u < nLeaves; u++) {
      int i0 = L->first[u];
      int wt = L->weight[u];
      if (wt == 1) {
        for (int k = 0; k < A; k++)
          out[(size_t)i0 * A + k] = rootLevel >= 0 ? L->rec[(size_t)u * A + k] : 0;
        continue;
      }
      int64_t attrSum[3], recDc[3];
      int64_t sq = orc_isqrt((uint64_t)wt << (2 * ORC_FRAC_BITS));
      for (int k = 0; k < A; k++) {
        attrSum[k] = orc_fxfromint(L->attr[(size_t)u * A + k]);
        int64_t r = rootLevel >= 0 ? L->rec[(size_t)u * A + k] : 0;
        recDc[k] = cx.ext ? r : orc_fxfromint(r);
        if (!cx.haar)
          recDc[k] = orc_fxmul(recDc[k], sq);
      }
      orc_quantizer q[2];
      {
        int32_t d0 = rootLevel >= 0 ? L->qpd[2 * u] : L->qp[2 * u];
        int32_t d1 = rootLevel >= 0 ? L->qpd[2 * u + 1] : L->qp[2 * u + 1];
        mk_quantizers(qs, qpLayer, d0 >> 4, d1 >> 4, q);
      }
      for (int w = wt - 1; w > 0; w--) {
        int64_t a, b;
        raht_ab(w, 1, &a, &b);
        for (int k = 0; k < A; k++) {
          orc_quantizer qk = q[k < 1 ? k : 1];
          int64_t t0, t1;
          if (cx.isEncoder) {
            t1 = orc_fxfromint(dupHf[(size_t)(i0 + w) * A + k]);
            if (cx.haar) {
              attrSum[k] -= t1 >> 1;
              t1 += attrSum[k];
              t0 = attrSum[k];
              int64_t hf = t1 - t0;
              t0 = t0 + ((hf >> (1 + ORC_FRAC_BITS)) << ORC_FRAC_BITS);
              t1 = hf;
            } else {
              attrSum[k] -= t1;
              t0 = scale_rsqrt(attrSum[k], w);
              int64_t lf = orc_fxmul(t1, b) + orc_fxmul(a, t0);
              int64_t hf = orc_fxmul(t1, a) - orc_fxmul(b, t0);
              t0 = lf;
              t1 = hf;
            }
            int64_t c = orc_quantize(qk, orc_fxround(t1) << ORC_ATTR_SHIFT);
            *cx.coef[k]++ = (int32_t)c;
            t1 = orc_fxfromint(
              orc_div_exp2_half_up(orc_scale(qk, c), ORC_ATTR_SHIFT));
          } else {
            int64_t c = *cx.coef[k]++;
            t1 = orc_fxfromint(
              orc_div_exp2_half_up(orc_scale(qk, c), ORC_ATTR_SHIFT));
          }
          t0 = recDc[k];
          int64_t left, right;
          if (cx.haar) {
            left = t0 - ((t1 >> (1 + ORC_FRAC_BITS)) << ORC_FRAC_BITS);
            right = t1 + left;
          } else {
            left = orc_fxmul(t0, a) - orc_fxmul(b, t1);
            right = orc_fxmul(t0, b) + orc_fxmul(a, t1);
          }
          recDc[k] = left;
          out[(size_t)(i0 + w) * A + k] = cx.ext ? right : orc_fxround(right);
          if (w == 1)
            out[(size_t)i0 * A + k] = cx.ext ? left : orc_fxround(left);
        }
      }
    }
  }
  for (size_t i = 0; i < (size_t)N * A; i++)
    attrs[i] = cx.ext ? (int32_t)((out[i] + ORC_ONE_HALF) >> ORC_FRAC_BITS)
                      : (int32_t)out[i];

  free(out);
  free(dupHf);
  for (int i = 0; i < nStages; i++)
    stage_free(&stages[i]);
  free(stages);
  return 0;
}

/* Morton code + stable sort (mortonAddr + std::sort with index tie-break,
 * AttributeEncoder.cpp:1316-1321, PCCTMC3Common.h:176-191). */
typedef struct {
  int64_t key;
  int32_t idx;
} keyidx_t;

static int
keyidx_cmp(const void* a, const void* b)
{
  const keyidx_t* x = (const keyidx_t*)a;
  const keyidx_t* y = (const keyidx_t*)b;
  if (x->key != y->key)
    return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int
oracle_morton_sort(const int32_t* xyz, int n, int64_t* keys, int32_t* order)
{
  keyidx_t* v = (keyidx_t*)malloc(sizeof(keyidx_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    v[i].key = orc_morton_addr(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    v[i].idx = i;
  }
  qsort(v, (size_t)n, sizeof(keyidx_t), keyidx_cmp);
  for (int i = 0; i < n; i++) {
    keys[i] = v[i].key;
    order[i] = v[i].idx;
  }
  free(v);
  return 0;
}

/* scalar exports for known-answer tests */
uint32_t oracle_isqrt(uint64_t x) { return orc_isqrt(x); }
uint64_t oracle_irsqrt(uint64_t x) { return orc_irsqrt(x); }
int64_t oracle_morton_addr(int32_t x, int32_t y, int32_t z) { return orc_morton_addr(x, y, z); }
uint64_t oracle_morton3d_add(uint64_t a, uint64_t b) { return orc_morton3d_add(a, b); }
int64_t oracle_quantize(int qp, int64_t x) { return orc_quantize(orc_mkquant(qp), x); }
int64_t oracle_scale(int qp, int64_t x) { return orc_scale(orc_mkquant(qp), x); }
int64_t oracle_fixed_mul(int64_t a, int64_t b) { return orc_fxmul(a, b); }
int64_t oracle_div_approx(int64_t a, uint64_t b, int32_t s) { return orc_div_approx(a, b, s); }
