/* lod_oracle.c — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Sequential plain-C restatement of the level-of-detail build of the
 * predicting / lifting transforms (intra, non-scalable):
 *
 *   AttributeLods::generate            tmc3/AttributeCommon.cpp:45-72
 *   buildPredictorsFast                tmc3/PCCTMC3Common.h:2300-2469
 *   subsampleByDistance / ByDecimation / ByOctree(+WithCentroid)
 *                                      tmc3/PCCTMC3Common.h:1984-2250
 *   computeNearestNeighbors            tmc3/PCCTMC3Common.h:1147-1953
 *   updateNearestNeigh*                tmc3/PCCTMC3Common.h:944-1143
 *   updatePredictors                   tmc3/PCCTMC3Common.h:2273-2296
 *   PCCPredictor::computeWeights / blendWeights  tmc3/PCCTMC3Common.h:589-693
 *
 * The restatement is organised differently from the reference: the reference
 * walks the refined points of a LoD in Morton order carrying cursors (`j`, the
 * atlas fill cursor `cubeIndex`, the cached 27-cell candidate list) and a
 * hash atlas; here every query is a pure function of (query, retained list)
 * — cell ranges are binary searches over the Morton-sorted retained list, `j`
 * is an upper bound, and the one genuinely stateful behaviour of the
 * reference (the atlas fill cursor stalls for good once some atlas that holds
 * retained points holds no query, PCCTMC3Common.h:1337-1349) is reproduced by
 * computing the first such atlas per LoD.  Parity is PINNED against the
 * compiled reference by tests/test_oracle_vs_reference.py::test_live_lod.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pcc_attr_b200.h"
#include "pcc_arith_oracle.h"

typedef struct {
  int64_t code;
  int32_t pos[3];
  int32_t idx;
} voxel_t;

static int
voxel_cmp(const void* a, const void* b)
{
  const voxel_t* x = (const voxel_t*)a;
  const voxel_t* y = (const voxel_t*)b;
  if (x->code != y->code)
    return x->code < y->code ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

typedef struct {
  int32_t min[3], max[3];
} box_t;

static void
box_init(box_t* b)
{
  for (int k = 0; k < 3; k++) {
    b->min[k] = INT32_MAX;
    b->max[k] = INT32_MIN;
  }
}
static void
box_insert(box_t* b, const int32_t* p)
{
  for (int k = 0; k < 3; k++) {
    if (p[k] < b->min[k]) b->min[k] = p[k];
    if (p[k] > b->max[k]) b->max[k] = p[k];
  }
}
static void
box_merge(box_t* b, const box_t* o)
{
  for (int k = 0; k < 3; k++) {
    if (o->min[k] < b->min[k]) b->min[k] = o->min[k];
    if (o->max[k] > b->max[k]) b->max[k] = o->max[k];
  }
}
/* Box3::getDist1, PCCMath.h:504-510 (int32 arithmetic) */
static int32_t
box_dist1(const box_t* b, const int32_t* p)
{
  int32_t s = 0;
  for (int k = 0; k < 3; k++) {
    int32_t a = b->min[k] - p[k], c = p[k] - b->max[k];
    int32_t d = a > 0 ? a : 0;
    if (c > d) d = c;
    s += d;
  }
  return s;
}

/* three-level bounding-box hierarchy over 32 / 1024 / 32768 consecutive
 * entries (BoxHierarchy<5,3>, PCCTMC3Common.h:58-107) */
typedef struct {
  box_t* lvl[3];
  int n[3];
} boxh_t;

static void
boxh_build(boxh_t* h, const int32_t (*bpos)[3], const uint32_t* list, int count)
{
  int c = count;
  for (int l = 0; l < 3; l++) {
    c = (c + 31) >> 5;
    h->n[l] = c;
    h->lvl[l] = (box_t*)malloc(sizeof(box_t) * (size_t)(c > 0 ? c : 1));
    for (int i = 0; i < c; i++)
      box_init(&h->lvl[l][i]);
  }
  for (int i = 0; i < count; i++)
    box_insert(&h->lvl[0][i >> 5], bpos[list[i]]);
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < h->n[l]; i++)
      box_merge(&h->lvl[l + 1][i >> 5], &h->lvl[l][i]);
}
static void
boxh_free(boxh_t* h)
{
  for (int l = 0; l < 3; l++)
    free(h->lvl[l]);
}

/* search state of one query (localIndexes / minDistances / index2) */
typedef struct {
  int32_t li[6];
  int64_t md[6];
  int index2;
  int distribution;
} nn_t;

static int64_t
norm1(const int32_t* a, const int32_t* b)
{
  int32_t s = 0;
  for (int k = 0; k < 3; k++) {
    int32_t d = a[k] - b[k];
    s += d < 0 ? -d : d;
  }
  return s;
}

/* updateNearestNeigh / updateNearestNeighByDistanceAndDistribution
 * (PCCTMC3Common.h:944-1069) */
static void
nn_update(nn_t* s, const int32_t* p0, const int32_t* p1, int32_t index)
{
  int64_t d = norm1(p0, p1);
  if (!s->distribution) {
    if (d >= s->md[2])
      return;
    if (d < s->md[0]) {
      s->md[2] = s->md[1]; s->md[1] = s->md[0]; s->md[0] = d;
      s->li[2] = s->li[1]; s->li[1] = s->li[0]; s->li[0] = index;
    } else if (d < s->md[1]) {
      s->md[2] = s->md[1]; s->md[1] = d;
      s->li[2] = s->li[1]; s->li[1] = index;
    } else {
      s->md[2] = d;
      s->li[2] = index;
    }
    return;
  }
  if (d > s->md[2]) {
    /* nothing */
  } else if (d < s->md[2]) {
    /* the evicted third neighbour becomes a spare candidate */
    if (s->li[2] != -1)
      s->li[s->index2++] = s->li[2];
    if (d < s->md[0]) {
      s->md[2] = s->md[1]; s->md[1] = s->md[0]; s->md[0] = d;
      s->li[2] = s->li[1]; s->li[1] = s->li[0]; s->li[0] = index;
    } else if (d < s->md[1]) {
      s->md[2] = s->md[1]; s->md[1] = d;
      s->li[2] = s->li[1]; s->li[1] = index;
    } else {
      s->md[2] = d;
      s->li[2] = index;
    }
  } else if (s->li[5] == -1) {
    s->li[s->index2++] = index; /* exact tie with the third neighbour */
  }
  if (s->index2 == 6)
    s->index2 = 3;
}

/* ...WithCheck variants (PCCTMC3Common.h:1073-1143) */
static void
nn_update_check(nn_t* s, const int32_t* p0, const int32_t* p1, int32_t index)
{
  int lim = s->distribution ? 6 : 3;
  for (int h = 0; h < lim; h++)
    if (s->li[h] == index)
      return;
  nn_update(s, p0, p1, index);
}

static const uint8_t kNeighOffset27[27] = {7,  3,  5,  6,  35, 21, 14, 28, 42,
                                           49, 12, 10, 17, 20, 34, 33, 4,  2,
                                           1,  56, 24, 40, 48, 32, 16, 8,  0};
static const uint8_t kNeighOffset20[20] = {7,  3,  5,  6,  12, 10, 17, 20, 34, 33,
                                           4,  2,  1,  24, 40, 48, 32, 16, 8,  0};

/* [lo, hi) of entries of `list` whose (code >> shift) == cell */
static void
cell_range(const voxel_t* v, const uint32_t* list, int n, int shift, int64_t cell,
           int* lo, int* hi)
{
  int a = 0, b = n;
  while (a < b) {
    int m = (a + b) >> 1;
    if ((v[list[m]].code >> shift) < cell) a = m + 1; else b = m;
  }
  *lo = a;
  b = n;
  while (a < b) {
    int m = (a + b) >> 1;
    if ((v[list[m]].code >> shift) <= cell) a = m + 1; else b = m;
  }
  *hi = a;
}

/* window search over [lo, hi] of the retained list with bounding-box pruning
 * (PCCTMC3Common.h:1422-1521); dir > 0 ascending, dir < 0 descending */
static void
window_search(nn_t* s, const boxh_t* h, const int32_t (*bpos)[3], const uint32_t* list,
              const int32_t* bp, int lo, int hi, int dir)
{
  if (lo > hi)
    return;
  const int b2lo = lo >> 15, b2hi = hi >> 15;
  const int b1lo = lo >> 10, b1hi = hi >> 10;
  const int b0lo = lo >> 5, b0hi = hi >> 5;
  for (int t2 = 0; t2 <= b2hi - b2lo; t2++) {
    int b2 = dir > 0 ? b2lo + t2 : b2hi - t2;
    if (s->li[2] != -1 && box_dist1(&h->lvl[2][b2], bp) >= s->md[2])
      continue;
    int a1 = b2 << 5;
    int s1 = b1lo > a1 ? b1lo : a1, e1 = b1hi < a1 + 31 ? b1hi : a1 + 31;
    for (int t1 = 0; t1 <= e1 - s1; t1++) {
      int b1 = dir > 0 ? s1 + t1 : e1 - t1;
      if (s->li[2] != -1 && box_dist1(&h->lvl[1][b1], bp) >= s->md[2])
        continue;
      int a0 = b1 << 5;
      int s0 = b0lo > a0 ? b0lo : a0, e0 = b0hi < a0 + 31 ? b0hi : a0 + 31;
      for (int t0 = 0; t0 <= e0 - s0; t0++) {
        int b0 = dir > 0 ? s0 + t0 : e0 - t0;
        if (s->li[2] != -1 && box_dist1(&h->lvl[0][b0], bp) >= s->md[2])
          continue;
        int a = b0 << 5;
        int k0 = lo > a ? lo : a, k1 = hi < a + 31 ? hi : a + 31;
        for (int t = 0; t <= k1 - k0; t++) {
          int k = dir > 0 ? k0 + t : k1 - t;
          nn_update_check(s, bp, bpos[list[k]], k);
        }
      }
    }
  }
}

typedef struct {
  uint32_t neighborCount;
  uint64_t weight[3];
  uint32_t index[3]; /* point index, later predictor index */
} pred_t;

/* PCCPredictor::computeWeights (PCCTMC3Common.h:589-633) */
static void
compute_weights(pred_t* p)
{
  const uint32_t shift = 1u << 8;
  int n = 0;
  while ((p->weight[0] >> n) >= shift)
    n++;
  if (n > 0)
    for (uint32_t i = 0; i < p->neighborCount; i++)
      p->weight[i] = (p->weight[i] + (1ull << (n - 1))) >> n;
  while (p->neighborCount > 1) {
    if (p->weight[p->neighborCount - 1] >= (p->weight[0] << 8))
      p->neighborCount--;
    else
      break;
  }
  if (p->neighborCount <= 1) {
    p->weight[0] = shift;
  } else if (p->neighborCount == 2) {
    uint64_t d0 = p->weight[0], d1 = p->weight[1];
    uint64_t w1 = (uint64_t)orc_div_approx((int64_t)d0, d0 + d1, 8);
    p->weight[0] = (uint32_t)(shift - w1);
    p->weight[1] = (uint32_t)w1;
  } else {
    p->neighborCount = 3;
    uint64_t d0 = p->weight[0], d1 = p->weight[1], d2 = p->weight[2];
    uint64_t sum = d1 * d2 + d0 * d2 + d0 * d1;
    uint64_t w2 = (uint64_t)orc_div_approx((int64_t)(d0 * d1), sum, 8);
    uint64_t w1 = (uint64_t)orc_div_approx((int64_t)(d0 * d2), sum, 8);
    p->weight[0] = (uint32_t)(shift - (w1 + w2));
    p->weight[1] = (uint32_t)w1;
    p->weight[2] = (uint32_t)w2;
  }
}

static int64_t
norm2(const int32_t* a, const int32_t* b)
{
  int64_t s = 0;
  for (int k = 0; k < 3; k++) {
    int64_t d = (int64_t)a[k] - b[k];
    s += d * d;
  }
  return s;
}

static int
dir_of(const int32_t* nb, const int32_t* p)
{
  return ((nb[0] - p[0] >= 0) << 2) + ((nb[1] - p[1] >= 0) << 1) + (nb[2] - p[2] >= 0);
}

int
oracle_lod_build(const pccb200_lod_params* lp, const int32_t* xyz, int N,
                 pccb200_predictor* predsOut, uint32_t* indexesOut,
                 uint32_t* nplOut, int32_t* lodCountOut)
{
  if (N <= 0 || lp->num_detail_levels < 1 || lp->num_detail_levels > PCCB200_MAX_LODS)
    return -1;
  voxel_t* v = (voxel_t*)malloc(sizeof(voxel_t) * (size_t)N);
  for (int i = 0; i < N; i++) {
    v[i].pos[0] = xyz[3 * i];
    v[i].pos[1] = xyz[3 * i + 1];
    v[i].pos[2] = xyz[3 * i + 2];
    v[i].code = orc_morton_addr(v[i].pos[0], v[i].pos[1], v[i].pos[2]);
    v[i].idx = i;
  }
  qsort(v, (size_t)N, sizeof(voxel_t), voxel_cmp);
  int32_t(*bpos)[3] = (int32_t(*)[3])malloc(sizeof(int32_t) * 3 * (size_t)N);
  for (int i = 0; i < N; i++)
    for (int k = 0; k < 3; k++)
      bpos[i][k] = v[i].pos[k] * lp->lod_neigh_bias[k];

  uint32_t* input = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
  uint32_t* retained = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
  uint32_t* indexes = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
  uint32_t* p2p = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)N);
  pred_t* preds = (pred_t*)calloc((size_t)N, sizeof(pred_t));
  int nInput = N, nIndexes = 0;
  for (int i = 0; i < N; i++)
    input[i] = (uint32_t)i;
  uint32_t npl[PCCB200_MAX_LODS + 2];
  int nNpl = 0;
  npl[nNpl++] = (uint32_t)N;
  int predIndex = N;
  const int L = lp->num_detail_levels;

  for (int lod = 0; nInput > 0 && lod < L; lod++) {
    const int start = nIndexes;
    int nRet = 0;
    if (lod == L - 1) {
      for (int i = 0; i < nInput; i++)
        indexes[nIndexes++] = input[i];
    } else if (lp->lod_decimation_type == 1) {
      /* subsampleByDecimation */
      int period = lp->lod_sampling_period[lod];
      for (int i = 0, j = 1; i < nInput; i++) {
        if (--j)
          indexes[nIndexes++] = input[i];
        else {
          retained[nRet++] = input[i];
          j = period;
        }
      }
    } else if (lp->lod_decimation_type == 2) {
      /* subsampleByOctree with centroid, backward direction */
      if (nInput == 1) {
        indexes[nIndexes++] = input[0];
      } else {
        const int nodeLog2 = lp->dist2 + lod;
        const int q = 3 * (nodeLog2 + 1);
        const int period = lp->lod_sampling_period[lod];
        const uint32_t mask = nodeLog2 ? (uint32_t)-1 << nodeLog2 : (uint32_t)-1;
        int g0 = 0;
        for (int i = 0; i < nInput; i++) {
          uint64_t cur = (uint64_t)(v[input[i]].code >> q);
          uint64_t nxt = i < nInput - 1 ? (uint64_t)(v[input[i + 1]].code >> q) : cur;
          if (!(i == nInput - 1 || cur < nxt))
            continue;
          int size = i - g0 + 1;
          if (size < period && i != nInput - 1)
            continue;
          int32_t cen[3] = {0, 0, 0};
          for (int t = g0; t <= i; t++)
            for (int k = 0; k < 3; k++)
              cen[k] += (int32_t)((uint32_t)v[input[t]].pos[k] & mask);
          int pick = i;
          int64_t best = INT64_MAX;
          for (int t = i; t >= g0; t--) {
            int32_t pp[3];
            for (int k = 0; k < 3; k++)
              pp[k] = (int32_t)((uint32_t)v[input[t]].pos[k] & mask) * size;
            int64_t m = norm1(pp, cen);
            if (best > m) {
              best = m;
              pick = t;
            }
          }
          for (int t = g0; t <= i; t++) {
            if (t == pick)
              retained[nRet++] = input[t];
            else
              indexes[nIndexes++] = input[t];
          }
          g0 = i + 1;
        }
      }
    } else {
      /* subsampleByDistance */
      if (nInput == 1) {
        indexes[nIndexes++] = input[0];
      } else {
        const int shiftBits0 = lp->dist2 + lod;
        const int64_t radius2 = 3ll << (shiftBits0 << 1);
        const int sb3 = 3 * (shiftBits0 + 1);
        const int atlasBoundaryBit = sb3 + 21 < 63 ? sb3 + 21 : 63;
        int64_t lastCell = -1;
        for (int i = 0; i < nInput; i++) {
          const voxel_t* pv = &v[input[i]];
          const int64_t atlasId = pv->code >> atlasBoundaryBit;
          const int64_t cell = pv->code >> sb3;
          if (nRet == 0) {
            retained[nRet++] = input[i];
            lastCell = cell;
            continue;
          }
          if (lastCell == cell) {
            indexes[nIndexes++] = input[i];
            continue;
          }
          const uint64_t base = orc_morton3d_add((uint64_t)cell, (uint64_t)-1ll);
          int found = 0;
          for (int n = 0; n < 20 && !found; n++) {
            const int64_t nb = (int64_t)orc_morton3d_add(base, kNeighOffset20[n]);
            if ((nb >> 21) != atlasId)
              continue;
            /* retained points of that cell that belong to the query's atlas
             * (the reference clears its atlas whenever the atlas id changes) */
            int lo, hi;
            cell_range(v, retained, nRet, sb3, nb, &lo, &hi);
            for (int k = lo; k < hi; k++) {
              if ((v[retained[k]].code >> atlasBoundaryBit) != atlasId)
                continue;
              if (norm2(v[retained[k]].pos, pv->pos) <= radius2) {
                found = 1;
                break;
              }
            }
          }
          if (found)
            indexes[nIndexes++] = input[i];
          else {
            retained[nRet++] = input[i];
            lastCell = cell;
          }
        }
      }
    }
    const int end = nIndexes;

    /* ---- nearest neighbours of indexes[start, end) among `retained` ---- */
    {
      const int R = nRet;
      const int shiftBits = 1 + lp->dist2 + lod;
      const int sb3 = 3 * shiftBits;
      const int atlasBoundaryBit = sb3 + 21 < 63 ? sb3 + 21 : 63;
      const int intra = lod >= lp->intra_lod_prediction_skip_layers;
      boxh_t hb, hi;
      boxh_build(&hb, (const int32_t(*)[3])bpos, retained, R);
      if (intra)
        boxh_build(&hi, (const int32_t(*)[3])bpos, indexes + start, end - start);

      /* first atlas that holds retained points but no query: from there on
       * the reference's fill cursor never moves again */
      int64_t stuckAtlas = INT64_MAX;
      {
        int qi = start;
        for (int r = 0; r < R;) {
          int64_t a = v[retained[r]].code >> atlasBoundaryBit;
          while (qi < end && (v[indexes[qi]].code >> atlasBoundaryBit) < a)
            qi++;
          if (!(qi < end && (v[indexes[qi]].code >> atlasBoundaryBit) == a)) {
            stuckAtlas = a;
            break;
          }
          while (r < R && (v[retained[r]].code >> atlasBoundaryBit) == a)
            r++;
        }
      }

      uint32_t* qidx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(end - start + 1));
      memcpy(qidx, indexes + start, sizeof(uint32_t) * (size_t)(end - start));
      for (int i = start; i < end; i++) {
        const uint32_t index = qidx[i - start];
        const voxel_t* pv = &v[index];
        const int32_t* bp = bpos[index];
        nn_t s;
        for (int h = 0; h < 6; h++) {
          s.li[h] = -1;
          s.md[h] = INT64_MAX;
        }
        s.index2 = 3;
        s.distribution = lp->prediction_with_distribution != 0;
        indexes[i] = (uint32_t)pv->idx;
        pred_t* pr = &preds[--predIndex];
        p2p[pv->idx] = (uint32_t)predIndex;

        if (R) {
          /* j: retained entries whose code does not exceed the query's */
          int j;
          {
            int a = 0, b = R;
            while (a < b) {
              int m = (a + b) >> 1;
              if (v[retained[m]].code <= pv->code) a = m + 1; else b = m;
            }
            j = a < R - 1 ? a : R - 1;
          }
          const int64_t atlasId = pv->code >> atlasBoundaryBit;
          const int64_t cell = pv->code >> sb3;
          if (atlasId < stuckAtlas) {
            const uint64_t base = orc_morton3d_add((uint64_t)cell, (uint64_t)-1ll);
            for (int n = 0; n < 27; n++) {
              const int64_t nb = (int64_t)orc_morton3d_add(base, kNeighOffset27[n]);
              if ((nb >> 21) != atlasId)
                continue;
              int lo, hi2;
              cell_range(v, retained, R, sb3, nb, &lo, &hi2);
              for (int k = lo; k < hi2; k++)
                if ((v[retained[k]].code >> atlasBoundaryBit) == atlasId)
                  nn_update(&s, bp, bpos[retained[k]], k);
            }
          }
          if (s.li[2] == -1) {
            const int center = s.li[0] == -1 ? j : s.li[0];
            const int range = lp->inter_lod_search_range;
            const int k0 = center - range > 0 ? center - range : 0;
            const int k1 = (int64_t)center + range < R - 1 ? center + range : R - 1;
            nn_update_check(&s, bp, bpos[retained[center]], center);
            for (int n = 1; n <= 2; n++) {
              if (center + n <= k1)
                nn_update_check(&s, bp, bpos[retained[center + n]], center + n);
              if (center - n >= k0)
                nn_update_check(&s, bp, bpos[retained[center - n]], center - n);
            }
            const int p1 = center + 3 < R - 1 ? center + 3 : R - 1;
            const int p0 = center - 3 > 0 ? center - 3 : 0;
            window_search(&s, &hb, (const int32_t(*)[3])bpos, retained, bp, p1, k1, +1);
            window_search(&s, &hb, (const int32_t(*)[3])bpos, retained, bp, k0, p0, -1);
          }
          /* retained-list positions -> sorted-voxel indices */
          for (int h = 0; h < 6; h++)
            if (s.li[h] != -1 && (h < 3 || s.distribution))
              s.li[h] = (int32_t)retained[s.li[h]];
        }

        if (intra) {
          /* candidates inside the same LoD: the following entries */
          const int k00 = i + 1;
          const int k01 = end - 1 < k00 + 2 ? end - 1 : k00 + 2;
          for (int k = k00; k <= k01; k++)
            nn_update(&s, bp, bpos[qidx[k - start]], (int32_t)qidx[k - start]);
          const int w0 = k01 + 1 - start;
          const int w1 = (end - 1 < k00 + lp->intra_lod_search_range
                            ? end - 1 : k00 + lp->intra_lod_search_range) - start;
          /* same pruning walk as window_search, but without the duplicate
           * check and with sorted-voxel indices as candidate ids */
          if (w0 <= w1) {
            const int b2lo = w0 >> 15, b2hi = w1 >> 15, b1lo = w0 >> 10, b1hi = w1 >> 10;
            const int b0lo = w0 >> 5, b0hi = w1 >> 5;
            for (int b2 = b2lo; b2 <= b2hi; b2++) {
              if (s.li[2] != -1 && box_dist1(&hi.lvl[2][b2], bp) >= s.md[2])
                continue;
              int a1 = b2 << 5;
              int s1 = b1lo > a1 ? b1lo : a1, e1 = b1hi < a1 + 31 ? b1hi : a1 + 31;
              for (int b1 = s1; b1 <= e1; b1++) {
                if (s.li[2] != -1 && box_dist1(&hi.lvl[1][b1], bp) >= s.md[2])
                  continue;
                int a0 = b1 << 5;
                int s0 = b0lo > a0 ? b0lo : a0, e0 = b0hi < a0 + 31 ? b0hi : a0 + 31;
                for (int b0 = s0; b0 <= e0; b0++) {
                  if (s.li[2] != -1 && box_dist1(&hi.lvl[0][b0], bp) >= s.md[2])
                    continue;
                  int a = b0 << 5;
                  int h0 = w0 > a ? w0 : a, h1 = w1 < a + 31 ? w1 : a + 31;
                  for (int h = h0; h <= h1; h++)
                    nn_update(&s, bp, bpos[qidx[h]], (int32_t)qidx[h]);
                }
              }
            }
          }
        }

        int nc = (s.li[0] != -1) + (s.li[1] != -1) + (s.li[2] != -1);
        if (nc > lp->num_pred_nearest_neighbours)
          nc = lp->num_pred_nearest_neighbours;
        if (s.distribution) {
          const int nc1 = 3 + (s.li[3] != -1) + (s.li[4] != -1) + (s.li[5] != -1);
          for (int m = 3; m < nc1; m++)
            if (s.md[m] == INT64_MAX)
              s.md[m] = norm1(bp, bpos[s.li[m]]);
          for (int m = 3; m < nc1; m++)
            for (int l = m + 1; l < nc1; l++)
              if (s.md[l] < s.md[m]) {
                int32_t ti = s.li[l]; s.li[l] = s.li[m]; s.li[m] = ti;
                int64_t td = s.md[l]; s.md[l] = s.md[m]; s.md[m] = td;
              }
          if (nc >= 3) {
            static const int loose[8][3] = {{3, 5, 6}, {2, 4, 7}, {1, 4, 7}, {0, 5, 6},
                                            {1, 2, 7}, {0, 3, 6}, {0, 3, 5}, {1, 2, 4}};
            int dir[6] = {-1, -1, -1, -1, -1, -1};
            int numend = 3;
            for (; numend < nc1; numend++)
              if ((s.md[numend] << 5) >= s.md[2] * 54)
                break;
            for (int h = 0; h < numend; h++)
              dir[h] = dir_of(bpos[s.li[h]], bp);
            int replace = 1, ridx = -1;
            if (dir[1] == 7 - dir[0] || dir[2] == 7 - dir[0] || dir[2] == 7 - dir[1])
              replace = 0;
            for (int h = 3; replace && h < numend; h++)
              if (dir[h] == 7 - dir[0] || dir[h] == 7 - dir[1]) {
                replace = 0;
                ridx = h;
              }
            const int e01 = dir[0] == dir[1], e02 = dir[0] == dir[2], e12 = dir[1] == dir[2];
            const int* ld = loose[dir[0]];
#define IN_LOOSE(x) ((x) == ld[0] || (x) == ld[1] || (x) == ld[2])
            if (replace) {
              if ((e02 || e12) && e01) {
                for (int h = 3; replace && h < numend; h++)
                  if (IN_LOOSE(dir[h])) {
                    replace = 0;
                    ridx = h;
                  }
              } else if ((e02 || e12) && !e01) {
                if (!IN_LOOSE(dir[1]))
                  for (int h = 3; replace && h < numend; h++)
                    if (dir[h] != dir[0] && dir[h] != dir[1]) {
                      replace = 0;
                      ridx = h;
                    }
              } else if (e01) {
                if (!IN_LOOSE(dir[2]))
                  for (int h = 3; replace && h < numend; h++)
                    if (IN_LOOSE(dir[h])) {
                      replace = 0;
                      ridx = h;
                    }
              }
            }
#undef IN_LOOSE
            if (ridx >= 0)
              s.li[2] = s.li[ridx];
          }
        }
        pr->neighborCount = (uint32_t)nc;
        for (int h = 0; h < nc; h++) {
          pr->index[h] = (uint32_t)v[s.li[h]].idx;
          pr->weight[h] = (uint64_t)norm2(bpos[s.li[h]], bp);
        }
        /* order by squared distance (PCCTMC3Common.h:1941-1951) */
#define SWAPN(a, b) do { uint64_t tw = pr->weight[a]; pr->weight[a] = pr->weight[b]; pr->weight[b] = tw; \
                         uint32_t tx = pr->index[a]; pr->index[a] = pr->index[b]; pr->index[b] = tx; } while (0)
        if (nc > 1) {
          if (pr->weight[0] > pr->weight[1])
            SWAPN(0, 1);
          if (nc == 3 && pr->weight[1] > pr->weight[2]) {
            SWAPN(1, 2);
            if (pr->weight[0] > pr->weight[1])
              SWAPN(0, 1);
          }
        }
#undef SWAPN
      }
      free(qidx);
      boxh_free(&hb);
      if (intra)
        boxh_free(&hi);
    }

    if (nRet)
      npl[nNpl++] = (uint32_t)nRet;
    memcpy(input, retained, sizeof(uint32_t) * (size_t)nRet);
    nInput = nRet;
  }

  /* reverse the order: coarse to fine */
  for (int i = 0; i < N / 2; i++) {
    uint32_t t = indexes[i];
    indexes[i] = indexes[N - 1 - i];
    indexes[N - 1 - i] = t;
  }
  /* updatePredictors + computeWeights (+ blendWeights) */
  for (int i = 0; i < N; i++) {
    pred_t* p = &preds[i];
    if (p->neighborCount < 2) {
      p->weight[0] = 1;
    } else if (p->weight[0] == 0) {
      p->neighborCount = 1;
      p->weight[0] = 1;
    }
    for (uint32_t k = 0; k < p->neighborCount; k++)
      p->index[k] = p2p[p->index[k]];
    compute_weights(p);
    if (lp->pred_weight_blending && p->neighborCount == 3) {
      const int32_t* n0 = &xyz[3 * (size_t)indexes[p->index[0]]];
      const int32_t* n1 = &xyz[3 * (size_t)indexes[p->index[1]]];
      const int32_t* n2 = &xyz[3 * (size_t)indexes[p->index[2]]];
      int64_t d01 = norm2(n0, n1), d02 = norm2(n0, n2), d12 = norm2(n1, n2);
      int w0 = (int)p->weight[0], w1 = (int)p->weight[1], w2 = (int)p->weight[2];
      int b1 = d01 <= d02 ? 1 : 5;
      int b2 = d01 <= d12 ? 5 : 1;
      int b3 = d02 <= d12 ? 1 : 5;
      int r0 = (w0 * 10 + w1 * (16 - 10 - b2) + w2 * b3) >> 4;
      int r1 = (w0 * b1 + w1 * 10 + w2 * (16 - 10 - b3)) >> 4;
      p->weight[0] = (uint64_t)r0;
      p->weight[1] = (uint64_t)r1;
      p->weight[2] = (uint64_t)(256 - r0 - r1);
    }
    predsOut[i].neighbor_count = p->neighborCount;
    for (int k = 0; k < 3; k++) {
      predsOut[i].predictor_index[k] = k < (int)p->neighborCount ? p->index[k] : 0;
      predsOut[i].weight[k] = k < (int)p->neighborCount ? (uint32_t)p->weight[k] : 0;
    }
    indexesOut[i] = indexes[i];
  }
  *lodCountOut = nNpl;
  for (int i = 0; i < nNpl; i++)
    nplOut[i] = npl[nNpl - 1 - i];

  free(v);
  free(bpos);
  free(input);
  free(retained);
  free(indexes);
  free(p2p);
  free(preds);
  return 0;
}
