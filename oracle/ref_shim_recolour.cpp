// ref_shim_recolour.cpp — TEST INFRASTRUCTURE ONLY (oracle).
//
// extern "C" wrapper around the UNMODIFIED reference recolouring
// (pcc::recolourColour / pcc::recolourReflectance,
// tmc3/pointset_processing.cpp:253-923, with its nanoflann kd-trees from
// dependencies/nanoflann), compiled where the sources lie by `make -C oracle
// recolourref` into oracle/_ref/libtmc13_recolour.so.  Only tests/ may load it.
#include <cstdint>

#include "PCCPointSet.h"
#include "hls.h"
#include "pointset_processing.h"

#include "pcc_attr_b200.h"

using namespace pcc;

extern "C" int
ref_recolour(const pccb200_recolour_params* p, const int32_t* sxyz, const int32_t* sattr, int A,
             int ns, double scale, const int32_t* off, const int32_t* txyz, int nt, int bitdepth,
             int32_t* out)
{
  PCCPointSet3 src, tgt;
  src.resize(ns);
  if (A == 3)
    src.addColors();
  else
    src.addReflectances();
  for (int i = 0; i < ns; i++) {
    src[i] = point_t(sxyz[3 * i], sxyz[3 * i + 1], sxyz[3 * i + 2]);
    if (A == 3)
      src.setColor(i, Vec3<attr_t>(attr_t(sattr[3 * i]), attr_t(sattr[3 * i + 1]), attr_t(sattr[3 * i + 2])));
    else
      src.setReflectance(i, attr_t(sattr[i]));
  }
  tgt.resize(nt);
  for (int i = 0; i < nt; i++)
    tgt[i] = point_t(txyz[3 * i], txyz[3 * i + 1], txyz[3 * i + 2]);

  AttributeDescription desc{};
  desc.attr_num_dimensions_minus1 = A - 1;
  desc.bitdepth = bitdepth;
  desc.attributeLabel = A == 3 ? KnownAttributeLabel::kColour : KnownAttributeLabel::kReflectance;

  RecolourParams rp;
  rp.distOffsetFwd = p->dist_offset_fwd;
  rp.distOffsetBwd = p->dist_offset_bwd;
  rp.maxGeometryDist2Fwd = p->max_geometry_dist2_fwd;
  rp.maxGeometryDist2Bwd = p->max_geometry_dist2_bwd;
  rp.maxAttributeDist2Fwd = p->max_attribute_dist2_fwd;
  rp.maxAttributeDist2Bwd = p->max_attribute_dist2_bwd;
  rp.searchRange = p->search_range;
  rp.numNeighboursFwd = p->num_neighbours_fwd;
  rp.numNeighboursBwd = p->num_neighbours_bwd;
  rp.useDistWeightedAvgFwd = p->use_dist_weighted_avg_fwd != 0;
  rp.useDistWeightedAvgBwd = p->use_dist_weighted_avg_bwd != 0;
  rp.skipAvgIfIdenticalSourcePointPresentFwd = p->skip_avg_if_identical_source_point_present_fwd != 0;
  rp.skipAvgIfIdenticalSourcePointPresentBwd = p->skip_avg_if_identical_source_point_present_bwd != 0;

  const point_t offset(off[0], off[1], off[2]);
  bool ok = A == 3 ? recolourColour(desc, rp, src, scale, offset, tgt)
                   : recolourReflectance(desc, rp, src, scale, offset, tgt);
  if (!ok)
    return -1;
  for (int i = 0; i < nt; i++) {
    if (A == 3) {
      const auto c = tgt.getColor(i);
      out[3 * i] = c[0];
      out[3 * i + 1] = c[1];
      out[3 * i + 2] = c[2];
    } else {
      out[i] = tgt.getReflectance(i);
    }
  }
  return 0;
}
