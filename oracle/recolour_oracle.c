/* recolour_oracle.c — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Plain-C restatement of the reference's attribute transfer, recolourColour /
 * recolourReflectance (tmc3/pointset_processing.cpp:253-923), with brute-force
 * nearest-neighbour searches in place of the nanoflann kd-trees.  Only tests/
 * may load it.
 *
 * What is restated, with the reference lines:
 *   forward search and the (never restored) result-vector pops    :301-326 / :660-685
 *   forward colour: identical-point shortcut, attribute-distance
 *     pruning, (distance-weighted) average                        :328-399 / :687-744
 *   backward search and per-target lists, sorted by distance      :401-435 / :746-778
 *   backward centroid with pops from the far end                  :437-516 / :780-848
 *   candidate search around the centroid (fixWeight: w = 0)       :517-608 / :849-919
 *
 * Two things the reference leaves to its libraries are fixed here, and the
 * product follows the same rules (so product == oracle bit for bit):
 *   - among equidistant candidates the lower point index wins (nanoflann keeps
 *     whichever its tree traversal met first);
 *   - entries of a backward list with equal distance are ordered by source
 *     index (std::sort is not stable).
 * tests/test_recolour.py pins this oracle against the compiled reference and
 * states how far the two rules move the result on clouds with distance ties.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../include/pcc_attr_b200.h"

typedef struct {
  double d;
  int32_t id;
} Cand;

static int
cand_before(double d0, int32_t i0, double d1, int32_t i1)
{
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

/* the k nearest of n points (int positions) to q, ascending (distance, index) */
static void
knn_brute(const int32_t* pts, int n, const double q[3], int k, Cand* out)
{
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    double r = 0.0;
    for (int c = 0; c < 3; c++) {
      const double diff = q[c] - (double)pts[3 * i + c];
      r += diff * diff; /* L2_Simple_Adaptor: summed in axis order */
    }
    if (cnt == k && !cand_before(r, i, out[k - 1].d, out[k - 1].id))
      continue;
    int pos = cnt < k ? cnt : k - 1;
    while (pos > 0 && cand_before(r, i, out[pos - 1].d, out[pos - 1].id)) {
      out[pos] = out[pos - 1];
      pos--;
    }
    out[pos].d = r;
    out[pos].id = i;
    if (cnt < k)
      cnt++;
  }
}

static double
clip_round(double v, double hi)
{
  const double r = round(v);
  return r < 0.0 ? 0.0 : r > hi ? hi : r;
}

static double
max_attr_dist2(const int32_t* attr, int A, const int32_t* ids, int n)
{
  double m = DBL_MIN;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int c = 0; c < A; c++) {
        const double df = (double)attr[(size_t)ids[i] * A + c] - (double)attr[(size_t)ids[j] * A + c];
        s += df * df;
      }
      if (s > m)
        m = s;
    }
  return m;
}

typedef struct {
  double d;
  int32_t src;
} ListEntry;

static int
list_cmp(const void* a, const void* b)
{
  const ListEntry* x = (const ListEntry*)a;
  const ListEntry* y = (const ListEntry*)b;
  if (x->d != y->d)
    return x->d < y->d ? -1 : 1;
  return x->src < y->src ? -1 : x->src > y->src;
}

int
oracle_recolour(const pccb200_recolour_params* p, const int32_t* sxyz, const int32_t* sattr, int A,
                int ns, double scale, const int32_t* off, const int32_t* txyz, int nt, int bitdepth,
                int32_t* out)
{
  const int kF = p->num_neighbours_fwd, kB = p->num_neighbours_bwd;
  if (ns <= 0 || nt <= 0 || (A != 1 && A != 3) || kF < 1 || kF > 16 || kB < 1 || kB > 16
      || kF > ns || kB > nt)
    return -1;
  const double invScale = 1.0 / scale;
  const double clipMax = (double)((1 << bitdepth) - 1);
  const double maxGeomF = p->max_geometry_dist2_fwd < 512 ? p->max_geometry_dist2_fwd : DBL_MAX;
  const double maxGeomB = p->max_geometry_dist2_bwd < 512 ? p->max_geometry_dist2_bwd : DBL_MAX;
  const double maxAttrF = p->max_attribute_dist2_fwd < 512 ? p->max_attribute_dist2_fwd : DBL_MAX;
  const double maxAttrB = p->max_attribute_dist2_bwd < 512 ? p->max_attribute_dist2_bwd : DBL_MAX;

  int32_t* refined1 = (int32_t*)malloc((size_t)nt * A * sizeof(int32_t));
  Cand cand[16];
  int32_t ids[16];

  /* forward direction; vecSize is the size of the reference's result vectors,
   * which shrink for good once a k-th neighbour is too far */
  int vecSize = kF;
  for (int t = 0; t < nt; t++) {
    double q[3];
    for (int c = 0; c < 3; c++)
      q[c] = (double)(txyz[3 * t + c] + off[c]) * invScale;
    knn_brute(sxyz, ns, q, kF, cand);
    while (vecSize != 1 && !(cand[kF - 1].d <= maxGeomF))
      vecSize--;
    for (int i = 0; i < kF; i++)
      ids[i] = cand[i].id;
    int nNN = vecSize;
    if (p->skip_avg_if_identical_source_point_present_fwd && cand[0].d < 0.0001)
      nNN = 1;
    while (nNN > 1 && max_attr_dist2(sattr, A, ids, nNN) > maxAttrF)
      nNN--;
    if (nNN == 1) {
      for (int c = 0; c < A; c++)
        refined1[(size_t)t * A + c] = sattr[(size_t)ids[0] * A + c];
      continue;
    }
    double acc[3] = {0.0, 0.0, 0.0};
    if (p->use_dist_weighted_avg_fwd) {
      double sumW = 0.0;
      for (int i = 0; i < nNN; i++) {
        const double w = 1 / (cand[i].d + p->dist_offset_fwd);
        for (int c = 0; c < A; c++)
          acc[c] += sattr[(size_t)ids[i] * A + c] * w;
        sumW += w;
      }
      for (int c = 0; c < A; c++)
        acc[c] /= sumW;
    } else {
      for (int i = 0; i < nNN; i++)
        for (int c = 0; c < A; c++)
          acc[c] += sattr[(size_t)ids[i] * A + c];
      for (int c = 0; c < A; c++)
        acc[c] /= nNN;
    }
    for (int c = 0; c < A; c++)
      refined1[(size_t)t * A + c] = (int32_t)clip_round(acc[c], clipMax);
  }

  /* backward direction: lists per target */
  int* count = (int*)calloc((size_t)nt + 1, sizeof(int));
  ListEntry* pairs = (ListEntry*)malloc((size_t)ns * kB * sizeof(ListEntry));
  int32_t* pairTgt = (int32_t*)malloc((size_t)ns * kB * sizeof(int32_t));
  size_t np = 0;
  for (int s = 0; s < ns; s++) {
    double q[3];
    for (int c = 0; c < 3; c++)
      q[c] = (double)sxyz[3 * s + c] * scale - (double)off[c];
    knn_brute(txyz, nt, q, kB, cand);
    for (int j = 0; j < kB; j++)
      if (cand[j].d <= maxGeomB) {
        pairs[np].d = cand[j].d;
        pairs[np].src = s;
        pairTgt[np] = cand[j].id;
        count[cand[j].id]++;
        np++;
      }
  }
  int* first = (int*)malloc(((size_t)nt + 1) * sizeof(int));
  int acc0 = 0;
  for (int t = 0; t <= nt; t++) {
    first[t] = acc0;
    if (t < nt)
      acc0 += count[t];
  }
  ListEntry* lists = (ListEntry*)malloc((np ? np : 1) * sizeof(ListEntry));
  int* cur = (int*)calloc((size_t)nt, sizeof(int));
  for (size_t i = 0; i < np; i++)
    lists[first[pairTgt[i]] + cur[pairTgt[i]]++] = pairs[i];

  const double rSource = 1.0 / (double)ns, rTarget = 1.0 / (double)nt;
  int32_t* lid = (int32_t*)malloc((np ? np : 1) * sizeof(int32_t));
  for (int t = 0; t < nt; t++) {
    const int32_t* c1 = refined1 + (size_t)t * A;
    ListEntry* L = lists + first[t];
    int n = first[t + 1] - first[t];
    if (n == 0) {
      for (int c = 0; c < A; c++)
        out[(size_t)t * A + c] = c1[c];
      continue;
    }
    qsort(L, n, sizeof(ListEntry), list_cmp);
    for (int i = 0; i < n; i++)
      lid[i] = L[i].src;
    double centroid2[3] = {0.0, 0.0, 0.0};
    int single = 0;
    if (p->skip_avg_if_identical_source_point_present_bwd && L[0].d < 0.0001) {
      n = 1;
      single = 1;
    }
    while (!single) {
      if (n == 1) {
        single = 1;
        break;
      }
      if (max_attr_dist2(sattr, A, lid, n) <= maxAttrB) {
        if (p->use_dist_weighted_avg_bwd) {
          double sumW = 0.0;
          for (int i = 0; i < n; i++) {
            const double w = 1 / (sqrt(L[i].d) + p->dist_offset_bwd);
            for (int c = 0; c < A; c++)
              centroid2[c] += sattr[(size_t)lid[i] * A + c] * w;
            sumW += w;
          }
          for (int c = 0; c < A; c++)
            centroid2[c] /= sumW;
        } else {
          for (int i = 0; i < n; i++)
            for (int c = 0; c < A; c++)
              centroid2[c] += sattr[(size_t)lid[i] * A + c];
          for (int c = 0; c < A; c++)
            centroid2[c] /= n;
        }
        break;
      }
      n--;
    }
    if (single)
      for (int c = 0; c < A; c++)
        centroid2[c] = sattr[(size_t)lid[0] * A + c];
    double c0[3] = {0, 0, 0}, best[3], col[3] = {0, 0, 0};
    for (int c = 0; c < A; c++)
      best[c] = c0[c] = clip_round(0.0 * c1[c] + 1.0 * centroid2[c], clipMax);
    double minError = DBL_MAX;
    const int R = p->search_range, R1 = A == 3 ? R : 0;
    for (int s1 = -R; s1 <= R; s1++) {
      col[0] = fmax(0.0, fmin(c0[0] + s1, clipMax));
      for (int s2 = -R1; s2 <= R1; s2++) {
        if (A == 3)
          col[1] = fmax(0.0, fmin(c0[1] + s2, clipMax));
        for (int s3 = -R1; s3 <= R1; s3++) {
          if (A == 3)
            col[2] = fmax(0.0, fmin(c0[2] + s3, clipMax));
          double e1 = 0.0, e2 = 0.0;
          for (int c = 0; c < A; c++) {
            const double df = col[c] - c1[c];
            e1 += df * df;
          }
          e1 *= rTarget;
          for (int i = 0; i < n; i++)
            for (int c = 0; c < A; c++) {
              const double df = col[c] - sattr[(size_t)lid[i] * A + c];
              e2 += df * df;
            }
          e2 *= rSource;
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
      out[(size_t)t * A + c] = (int32_t)best[c];
  }
  free(lid);
  free(cur);
  free(lists);
  free(first);
  free(pairTgt);
  free(pairs);
  free(count);
  free(refined1);
  return 0;
}
