/* dist2_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of estimateDist2, tmc3/AttributeEncoder.cpp:1683-1720
 * (nearest neighbour in coding order per sample, percentile by selection,
 * smallest shift s with 3 << 2s >= the selected distance). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int
cmp_i64(const void* a, const void* b)
{
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : x > y;
}

int
oracle_estimate_dist2(const int32_t* xyz, int32_t n, int32_t sampling_period, int32_t search_range,
                      float percentile)
{
  if (n < 2)
    return 0;
  int64_t* dists = malloc(sizeof(int64_t) * (size_t)(n / sampling_period + 1));
  size_t m = 0;
  for (int32_t index = 0; index < n; index += sampling_period) {
    int k0 = index - search_range > 0 ? index - search_range : 0;
    int k1 = index + search_range < n - 1 ? index + search_range : n - 1;
    int64_t d2 = INT64_MAX;
    for (int k = k0; k <= k1; k++) {
      if (k == index)
        continue;
      int64_t s = 0;
      for (int c = 0; c < 3; c++) {
        int64_t d = (int64_t)xyz[3 * (size_t)index + c] - xyz[3 * (size_t)k + c];
        s += d * d;
      }
      if (s < d2)
        d2 = s;
    }
    dists[m++] = d2;
  }
  int p = (int)floorf((float)m * percentile);
  qsort(dists, m, sizeof(int64_t), cmp_i64); /* nth_element: the p-th smallest */
  int64_t dist2 = dists[p];
  free(dists);
  int shift = 0;
  while (((int64_t)3 << (shift << 1)) < dist2 && shift < 20)
    ++shift;
  return shift;
}
