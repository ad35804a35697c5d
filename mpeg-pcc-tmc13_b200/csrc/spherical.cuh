// spherical.cuh — spherical coordinates for attribute coding of LiDAR slices
// (SURVEY 8(f) row N2): the step right before the attribute transforms when
// attr_aps.spherical_coord_flag is set.  Per-point integer arithmetic, one
// thread per point; positions stay on the device for the transforms that follow.
//
// Reference: convertXyzToRpl / offsetAndScale, tmc3/coordinate_conversion.cpp:44-122
// (called from tmc3/encoder.cpp:1151-1197, tmc3/decoder.cpp:874-920);
// findLaser, tmc3/geometry_octree.cpp:855-874; iatan2, tmc3/misc.cpp:278-309.
#pragma once

#include "pcc_arith.cuh"

namespace pccb200 {

// asin(i / 512) in Q20, i = 0..362, last entry repeated (tools/gen_asin_table.py)
#define PCC_ASIN_Q20_VALUES \
  0, 2048, 4096, 6144, 8192, 10240, 12288, 14336, 16385, 18433, \
  20481, 22530, 24578, 26627, 28676, 30724, 32773, 34822, 36872, 38921, \
  40970, 43020, 45070, 47120, 49170, 51220, 53271, 55322, 57373, 59424, \
  61475, 63527, 65579, 67631, 69683, 71736, 73789, 75842, 77896, 79949, \
  82004, 84058, 86113, 88168, 90223, 92279, 94335, 96392, 98449, 100506, \
  102563, 104621, 106680, 108739, 110798, 112858, 114918, 116978, 119040, 121101, \
  123163, 125225, 127288, 129352, 131416, 133480, 135545, 137611, 139677, 141743, \
  143810, 145878, 147946, 150015, 152085, 154155, 156225, 158297, 160368, 162441, \
  164514, 166588, 168662, 170737, 172813, 174890, 176967, 179045, 181123, 183203, \
  185283, 187363, 189445, 191527, 193610, 195694, 197779, 199864, 201950, 204037, \
  206125, 208214, 210303, 212393, 214485, 216577, 218669, 220763, 222858, 224954, \
  227050, 229148, 231246, 233345, 235445, 237547, 239649, 241752, 243856, 245961, \
  248068, 250175, 252283, 254392, 256502, 258614, 260726, 262840, 264954, 267070, \
  269187, 271305, 273424, 275544, 277666, 279788, 281912, 284037, 286163, 288290, \
  290419, 292549, 294680, 296812, 298945, 301080, 303216, 305354, 307492, 309632, \
  311773, 313916, 316060, 318206, 320352, 322500, 324650, 326801, 328953, 331107, \
  333262, 335419, 337577, 339737, 341898, 344061, 346225, 348391, 350558, 352727, \
  354897, 357069, 359243, 361418, 363595, 365773, 367953, 370135, 372318, 374503, \
  376690, 378879, 381069, 383261, 385455, 387650, 389847, 392046, 394247, 396450, \
  398655, 400861, 403069, 405279, 407491, 409705, 411921, 414139, 416359, 418581, \
  420804, 423030, 425258, 427488, 429720, 431954, 434190, 436428, 438668, 440910, \
  443155, 445401, 447650, 449901, 452155, 454410, 456668, 458928, 461190, 463455, \
  465722, 467991, 470262, 472536, 474813, 477091, 479373, 481656, 483942, 486231, \
  488522, 490815, 493111, 495410, 497711, 500015, 502322, 504631, 506943, 509257, \
  511574, 513894, 516217, 518542, 520870, 523201, 525535, 527872, 530211, 532553, \
  534899, 537247, 539598, 541952, 544310, 546670, 549033, 551399, 553769, 556142, \
  558517, 560896, 563278, 565664, 568052, 570444, 572839, 575238, 577640, 580045, \
  582454, 584866, 587282, 589701, 592123, 594549, 596979, 599412, 601849, 604290, \
  606734, 609183, 611634, 614090, 616549, 619013, 621480, 623951, 626426, 628905, \
  631388, 633875, 636366, 638862, 641361, 643865, 646373, 648885, 651401, 653922, \
  656447, 658976, 661510, 664049, 666592, 669139, 671691, 674248, 676809, 679375, \
  681946, 684522, 687103, 689688, 692278, 694874, 697474, 700080, 702690, 705306, \
  707927, 710553, 713184, 715821, 718463, 721111, 723764, 726423, 729087, 731757, \
  734433, 737115, 739802, 742495, 745194, 747899, 750611, 753328, 756051, 758781, \
  761517, 764259, 767008, 769763, 772525, 775294, 778069, 780850, 783639, 786435, \
  789237, 792047, 794863, 797687, 800518, 803357, 806202, 809056, 811917, 814785, \
  817662, 820546, 823438, 823438

PCC_TABLE(int32_t, kAsinQ20, 364, {PCC_ASIN_Q20_VALUES})

// atan(y / x) in Q20 for 0 <= y <= x (iatan2Core)
PCC_HD int
iatan2_core(int y, int x)
{
  if (x == 0)
    return 0;
  const uint64_t rinv = irsqrt64(uint64_t(x) * uint64_t(x) + uint64_t(y) * uint64_t(y));
  const int r = int((uint64_t(int64_t(y)) * rinv) >> 20);  // sin of the angle, 20 bits
  const int idx = r >> 11;
  const int lambda = r - (idx << 11);
  return kAsinQ20(idx) + ((lambda * (kAsinQ20(idx + 1) - kAsinQ20(idx))) >> 11);
}

// four-quadrant arc tangent in Q20 radians: pi = 3294199
PCC_HD int
iatan2_q20(int y, int x)
{
  const int xa = x < 0 ? -x : x;
  const int ya = y < 0 ? -y : y;
  int t = ya <= xa ? iatan2_core(ya, xa) : 1647099 - iatan2_core(xa, ya);
  if (x < 0)
    t = 3294199 - t;
  return y < 0 ? -t : t;
}

// index of the laser whose elevation tangent is nearest to the point's
PCC_HD int
find_laser(int px, int py, int pz, const int32_t* theta, int numTheta)
{
  if (numTheta == 1)
    return 0;
  const int64_t xl = int64_t(px) << 8;
  const int64_t yl = int64_t(py) << 8;
  const int64_t rinv = int64_t(irsqrt64(uint64_t(xl * xl + yl * yl)));
  const int theta32 = int((int64_t(pz) * rinv) >> 14);
  // first entry of theta[1 .. numTheta-2] greater than theta32 (else numTheta-1)
  int lo = 1, hi = numTheta - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (theta[mid] > theta32)
      hi = mid;
    else
      lo = mid + 1;
  }
  if (theta32 - theta[lo - 1] <= theta[lo] - theta32)
    lo--;
  return lo;
}

PCC_HD void
atomic_min_i32(int32_t* p, int32_t v)
{
#if defined(__CUDA_ARCH__)
  atomicMin(p, v);
#else
  if (v < *p)
    *p = v;
#endif
}
PCC_HD void
atomic_max_i32(int32_t* p, int32_t v)
{
#if defined(__CUDA_ARCH__)
  atomicMax(p, v);
#else
  if (v > *p)
    *p = v;
#endif
}

// (x, y, z) -> (radius, azimuth, laser index), and the bounding box of the result
struct XyzToRplFn {
  int32_t origin[3];
  const int32_t* theta;
  int numTheta;
  const int32_t* xyz;
  int32_t* rpl;
  int32_t* bbox;  // min[3], max[3]; initialised to INT32_MAX / INT32_MIN
  PCC_HD void operator()(int64_t i) const
  {
    const int px = xyz[3 * i] - origin[0];
    const int py = xyz[3 * i + 1] - origin[1];
    const int pz = xyz[3 * i + 2] - origin[2];
    const int laser = find_laser(px, py, pz, theta, numTheta);
    // (the reference shifts in 32 bits before widening)
    const int64_t xl = int64_t(int32_t(uint32_t(px) << 8));
    const int64_t yl = int64_t(int32_t(uint32_t(py) << 8));
    int32_t out[3];
    out[0] = int32_t(isqrt64(uint64_t(xl * xl + yl * yl)) >> 8);
    out[1] = (iatan2_q20(int(yl), int(xl)) + 3294199) >> 8;
    out[2] = laser;
    for (int k = 0; k < 3; k++) {
      rpl[3 * i + k] = out[k];
      if (out[k] < bbox[k])
        atomic_min_i32(&bbox[k], out[k]);
      if (out[k] > bbox[3 + k])
        atomic_max_i32(&bbox[3 + k], out[k]);
    }
  }
};

// (pos - minPos) * axisWeight, rounded, / 256 (32-bit arithmetic like the reference)
struct OffsetScaleFn {
  int32_t minPos[3];
  int32_t weight[3];
  int32_t* pos;
  PCC_HD void operator()(int64_t i) const
  {
    for (int k = 0; k < 3; k++) {
      const uint32_t d = uint32_t(pos[3 * i + k]) - uint32_t(minPos[k]);
      const int32_t v = int32_t(d * uint32_t(weight[k]) + 128u);
      pos[3 * i + k] = v >> 8;
    }
  }
};

struct BboxInitFn {
  int32_t* bbox;
  PCC_HD void operator()(int64_t k) const { bbox[k] = k < 3 ? INT32_MAX : INT32_MIN; }
};

// convertXyzToRpl then, if weight != null, offsetAndScale with minPos (or the
// bounding-box minimum when minPos == null, as the intra encoder does,
// encoder.cpp:1185-1194).  hostBbox receives min[3], max[3] of the conversion.
template<class Exec>
void
run_xyz_to_rpl(Exec& ex, const int32_t origin[3], const int32_t* dTheta, int numTheta,
               const int32_t* dXyz, int64_t n, int32_t* dRpl, int32_t hostBbox[6],
               const int32_t* minPos, const int32_t* weight)
{
  int32_t* dBox = ex.template alloc<int32_t>(6);
  ex.foreach(6, BboxInitFn{dBox});
  XyzToRplFn fn;
  for (int k = 0; k < 3; k++)
    fn.origin[k] = origin[k];
  fn.theta = dTheta;
  fn.numTheta = numTheta;
  fn.xyz = dXyz;
  fn.rpl = dRpl;
  fn.bbox = dBox;
  ex.foreach(n, fn);
  ex.download(hostBbox, dBox, 6 * sizeof(int32_t));
  if (weight) {
    OffsetScaleFn os;
    for (int k = 0; k < 3; k++) {
      os.minPos[k] = minPos ? minPos[k] : hostBbox[k];
      os.weight[k] = weight[k];
    }
    os.pos = dRpl;
    ex.foreach(n, os);
  }
}

}  // namespace pccb200
