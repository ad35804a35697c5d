/* spherical_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md 2).
 *
 * Plain-C restatement of the spherical-coordinate conversion used for attribute
 * coding of LiDAR slices:
 *   convertXyzToRpl, offsetAndScale   tmc3/coordinate_conversion.cpp:44-69,108-117
 *   findLaser                         tmc3/geometry_octree.cpp:855-874
 *   iatan2Core, iatan2                tmc3/misc.cpp:278-309
 * The arcsine table is derived here from its definition (round(asin(i/512) * 2^20),
 * last entry repeated), not shared with the product.  Pinned against the compiled
 * reference by tests/test_oracle_vs_reference.py::test_spherical_*. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "pcc_arith_oracle.h"

static int32_t g_asin[364];
static int g_asin_ready;

static void
asin_init(void)
{
  if (g_asin_ready)
    return;
  for (int i = 0; i < 363; i++)
    g_asin[i] = (int32_t)floor(asin(i / 512.0) * 1048576.0 + 0.5);
  g_asin[363] = g_asin[362];
  g_asin_ready = 1;
}

static int
atan_core(int y, int x) /* 0 <= y <= x */
{
  if (x == 0)
    return 0;
  uint64_t rinv = orc_irsqrt((uint64_t)x * (uint64_t)x + (uint64_t)y * (uint64_t)y);
  int r = (int)(((uint64_t)(int64_t)y * rinv) >> 20);
  int idx = r >> 11;
  int lambda = r - (idx << 11);
  return g_asin[idx] + ((lambda * (g_asin[idx + 1] - g_asin[idx])) >> 11);
}

int
oracle_iatan2(int y, int x)
{
  asin_init();
  int xa = abs(x), ya = abs(y);
  int t = ya <= xa ? atan_core(ya, xa) : 1647099 - atan_core(xa, ya);
  if (x < 0)
    t = 3294199 - t;
  return y < 0 ? -t : t;
}

int
oracle_find_laser(const int32_t pos[3], const int32_t* theta, int num_theta)
{
  if (num_theta == 1)
    return 0;
  int64_t xl = (int64_t)pos[0] << 8, yl = (int64_t)pos[1] << 8;
  int64_t rinv = (int64_t)orc_irsqrt((uint64_t)(xl * xl + yl * yl));
  int theta32 = (int)(((int64_t)pos[2] * rinv) >> 14);
  /* std::upper_bound over theta[1 .. num_theta - 2] */
  int it = num_theta - 1;
  for (int i = 1; i < num_theta - 1; i++)
    if (theta[i] > theta32) {
      it = i;
      break;
    }
  if (theta32 - theta[it - 1] <= theta[it] - theta32)
    it--;
  return it;
}

/* bbox: min[3], max[3] */
void
oracle_xyz_to_rpl(const int32_t origin[3], const int32_t* theta, int num_theta,
                  const int32_t* xyz, int64_t n, int32_t* rpl, int32_t bbox[6])
{
  for (int k = 0; k < 3; k++) {
    bbox[k] = INT32_MAX;
    bbox[3 + k] = INT32_MIN;
  }
  for (int64_t i = 0; i < n; i++) {
    int32_t pos[3];
    for (int k = 0; k < 3; k++)
      pos[k] = xyz[3 * i + k] - origin[k];
    int laser = oracle_find_laser(pos, theta, num_theta);
    /* the shift is done in 32 bits, then widened */
    int64_t xl = (int32_t)((uint32_t)pos[0] << 8);
    int64_t yl = (int32_t)((uint32_t)pos[1] << 8);
    int32_t out[3];
    out[0] = (int32_t)(orc_isqrt((uint64_t)(xl * xl + yl * yl)) >> 8);
    out[1] = (oracle_iatan2((int)yl, (int)xl) + 3294199) >> 8;
    out[2] = laser;
    for (int k = 0; k < 3; k++) {
      rpl[3 * i + k] = out[k];
      if (out[k] < bbox[k])
        bbox[k] = out[k];
      if (out[k] > bbox[3 + k])
        bbox[3 + k] = out[k];
    }
  }
}

void
oracle_offset_and_scale(const int32_t min_pos[3], const int32_t weight[3], int32_t* pos, int64_t n)
{
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      uint32_t d = (uint32_t)pos[3 * i + k] - (uint32_t)min_pos[k];
      int32_t v = (int32_t)(d * (uint32_t)weight[k] + 128u);
      pos[3 * i + k] = v >> 8;
    }
}
