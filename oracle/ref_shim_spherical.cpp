// ref_shim_spherical.cpp — TEST INFRASTRUCTURE ONLY.
// C entry points into the UNMODIFIED reference's spherical-coordinate
// conversion (tmc3/coordinate_conversion.cpp, findLaser in
// tmc3/geometry_octree.cpp, iatan2 in tmc3/misc.cpp), linked into
// _ref/libtmc13_lift.so (`make liftref`: all reference objects).
#include <cstdint>
#include <vector>

#include "PCCMath.h"
#include "PCCMisc.h"
#include "coordinate_conversion.h"
#include "geometry_octree.h"

using namespace pcc;

extern "C" {

int
tmc13ref_iatan2(int y, int x)
{
  return iatan2(y, x);
}

int
tmc13ref_find_laser(const int32_t pos[3], const int32_t* theta, int numTheta)
{
  return findLaser(point_t(pos[0], pos[1], pos[2]), theta, numTheta);
}

void
tmc13ref_xyz_to_rpl(const int32_t origin[3], const int32_t* theta, int numTheta,
                    const int32_t* xyz, int64_t n, int32_t* rpl, int32_t bbox[6])
{
  std::vector<Vec3<int>> src(n), dst(n);
  for (int64_t i = 0; i < n; i++)
    src[i] = Vec3<int>(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  Box3<int> box = convertXyzToRpl(
    Vec3<int>(origin[0], origin[1], origin[2]), theta, numTheta, src.data(), src.data() + n,
    dst.data());
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++)
      rpl[3 * i + k] = dst[i][k];
  for (int k = 0; k < 3; k++) {
    bbox[k] = box.min[k];
    bbox[3 + k] = box.max[k];
  }
}

void
tmc13ref_offset_and_scale(const int32_t minPos[3], const int32_t weight[3], int32_t* pos, int64_t n)
{
  std::vector<Vec3<int>> p(n);
  for (int64_t i = 0; i < n; i++)
    p[i] = Vec3<int>(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
  offsetAndScale(
    Vec3<int>(minPos[0], minPos[1], minPos[2]), Vec3<int>(weight[0], weight[1], weight[2]),
    p.data(), p.data() + n);
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++)
      pos[3 * i + k] = p[i][k];
}

// normalisedAxesWeights(Box3{0, {r, twoPi, maxLaserIdx}}, forcedMaxLog2)
void
tmc13ref_normalised_axes_weights(const int32_t boxMax[3], int forcedMaxLog2, int32_t out[3])
{
  Box3<int> box{0, {boxMax[0], boxMax[1], boxMax[2]}};
  auto w = normalisedAxesWeights(box, forcedMaxLog2);
  for (int k = 0; k < 3; k++)
    out[k] = w[k];
}

}  // extern "C"
