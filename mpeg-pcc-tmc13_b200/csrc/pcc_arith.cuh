// pcc_arith.cuh — integer / fixed-point primitives of the attribute-transform
// path, usable from device code (and from host code, so that the kernel
// bodies can be unit-tested without a GPU).
//
// Bit-exact counterparts of (paths relative to the TMC13 tree):
//   FixedPoint           tmc3/FixedPoint.h:44-122      -> fx_mul, fx_round, fx_from_int
//   irsqrt / isqrt       tmc3/misc.cpp:138-225         -> irsqrt64, isqrt64
//   ilog2                tmc3/PCCMisc.h:149-165        -> ilog2_u64
//   Quantizer            tmc3/quantization.h:53-102,
//                        tmc3/quantization.cpp:46-52   -> Quantizer
//   QpSet::quantizers    tmc3/quantization.cpp:169-178 -> make_quantizers
//   morton3dAdd          tmc3/PCCMisc.h:244-256        -> morton3d_add
//   mortonAddr           tmc3/PCCMath.h:605-626        -> morton_addr
//   divExp2RoundHalf*    tmc3/PCCMath.h:650-685
//   divApprox            tmc3/PCCMath.h:714-736
#pragma once

#include <stdint.h>

#include "pcc_attr_b200.h"

#if defined(__CUDACC__)
#  define PCC_HD __host__ __device__ __forceinline__
#  define PCC_HD_NOINLINE __host__ __device__
#else
#  define PCC_HD inline
#  define PCC_HD_NOINLINE
#endif

namespace pccb200 {

constexpr int kFracBits = 15;
constexpr int64_t kOneHalf = 1 << (kFracBits - 1);
constexpr int kAttrShift = 8;  // kFixedPointAttributeShift

//----------------------------------------------------------------------------
// count leading zeros / bit length

PCC_HD int
clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

PCC_HD int
clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

PCC_HD int
popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// floor(log2(x)), -1 for 0
PCC_HD int
ilog2_u64(uint64_t x)
{
  return 63 - clz64(x);
}

//----------------------------------------------------------------------------
// Q.15 fixed point

PCC_HD int64_t
fx_mul(int64_t a, int64_t b)
{
  int64_t v = a * b;
  return v < 0 ? -((kOneHalf - v) >> kFracBits) : ((kOneHalf + v) >> kFracBits);
}

PCC_HD int64_t
fx_round(int64_t v)
{
  return v > 0 ? ((kOneHalf + v) >> kFracBits) : -((kOneHalf - v) >> kFracBits);
}

// sign-magnitude shift == multiplication by 2^15
PCC_HD int64_t
fx_from_int(int64_t v)
{
  return v * (int64_t(1) << kFracBits);
}

PCC_HD int64_t
div_exp2_round_half_up(int64_t x, int shift)
{
  return shift ? (x + (int64_t(1) << (shift - 1))) >> shift : x;
}

PCC_HD int64_t
div_exp2_round_half_inf(int64_t x, int shift)
{
  if (!shift)
    return x;
  int64_t s0 = int64_t(1) << (shift - 1);
  return x >= 0 ? (s0 + x) >> shift : -((s0 - x) >> shift);
}

PCC_HD uint64_t
div_exp2_round_half_inf_u(uint64_t x, int shift)
{
  return shift ? ((uint64_t(1) << (shift - 1)) + x) >> shift : x;
}

// Small lookup tables indexed at run time: a local array would be rebuilt on
// the stack of every device thread, so device code reads a __constant__ copy.
//   PCC_TABLE(uint8_t, kFoo, 3, {1, 2, 3})   defines   kFoo(i)
#if defined(__CUDACC__)
#  define PCC_TABLE(type, name, n, ...) \
    static const type name##_host[n] = __VA_ARGS__; \
    static __device__ __constant__ type name##_dev[n] = __VA_ARGS__; \
    PCC_HD type name(int i) \
    { \
      PCC_TABLE_BODY(name) \
    }
#  if defined(__CUDA_ARCH__)
#    define PCC_TABLE_BODY(name) return name##_dev[i];
#  else
#    define PCC_TABLE_BODY(name) return name##_host[i];
#  endif
#else
#  define PCC_TABLE(type, name, n, ...) \
    static const type name##_host[n] = __VA_ARGS__; \
    PCC_HD type name(int i) { return name##_host[i]; }
#endif

//----------------------------------------------------------------------------
// inverse square root: 96-entry seed, two Newton iterations.  The seed tables
// are normative constants of the G-PCC specification.

#define PCC_IRSQRT_3R_VALUES \
  3196059648u, 3145728000u, 3107979264u, 3057647616u, 3019898880u, 2969567232u, \
  2931818496u, 2894069760u, 2868903936u, 2831155200u, 2793406464u, 2768240640u, \
  2730491904u, 2705326080u, 2667577344u, 2642411520u, 2617245696u, 2592079872u, \
  2566914048u, 2541748224u, 2516582400u, 2491416576u, 2466250752u, 2441084928u, \
  2428502016u, 2403336192u, 2378170368u, 2365587456u, 2340421632u, 2327838720u, \
  2302672896u, 2290089984u, 2264924160u, 2252341248u, 2239758336u, 2214592512u, \
  2202009600u, 2189426688u, 2164260864u, 2151677952u, 2139095040u, 2126512128u, \
  2113929216u, 2101346304u, 2088763392u, 2076180480u, 2051014656u, 2038431744u, \
  2025848832u, 2013265920u, 2000683008u, 2000683008u, 1988100096u, 1962934272u, \
  1962934272u, 1950351360u, 1937768448u, 1925185536u, 1912602624u, 1900019712u, \
  1900019712u, 1887436800u, 1874853888u, 1862270976u, 1849688064u, 1849688064u, \
  1837105152u, 1824522240u, 1811939328u, 1811939328u, 1799356416u, 1786773504u, \
  1786773504u, 1774190592u, 1761607680u, 1761607680u, 1749024768u, 1736441856u, \
  1736441856u, 1723858944u, 1723858944u, 1711276032u, 1698693120u, 1698693120u, \
  1686110208u, 1686110208u, 1673527296u, 1660944384u, 1660944384u, 1648361472u, \
  1648361472u, 1635778560u, 1635778560u, 1623195648u, 1623195648u, 1610612736u

#define PCC_IRSQRT_R3_VALUES \
  4195081216u, 3999986688u, 3857709056u, 3673323520u, 3538940928u, 3364924416u, \
  3238224896u, 3114735616u, 3034196992u, 2915990528u, 2800922624u, 2725880832u, \
  2615890944u, 2544223232u, 2439185408u, 2370818048u, 2303728640u, 2237913088u, \
  2173355008u, 2110061568u, 2048008192u, 1987165184u, 1927563264u, 1869150208u, \
  1840392192u, 1783783424u, 1728321536u, 1701024768u, 1647311872u, 1620883456u, \
  1568898048u, 1543306240u, 1492993024u, 1468236800u, 1443762176u, 1395656704u, \
  1372007424u, 1348605952u, 1302626304u, 1280060416u, 1257736192u, 1235650560u, \
  1213861888u, 1192294400u, 1171008512u, 1149979648u, 1108673536u, 1088379904u, \
  1068352512u, 1048567808u, 1029031936u, 1029036032u, 1009729536u, 971888640u, \
  971882496u,  953319424u,  934993920u,  916897792u,  899011584u,  881389568u, \
  881392640u,  864009216u,  846846976u,  829900800u,  813182976u,  813201408u, \
  796721152u,  780459008u,  764412928u,  764417024u,  748601344u,  732995584u, \
  733017088u,  717624320u,  702468096u,  702466048u,  687520768u,  672786432u, \
  672787456u,  658258944u,  658256896u,  643947520u,  629854208u,  629862400u, \
  615976960u,  615952384u,  602276864u,  588779520u,  588804096u,  575512576u, \
  575526912u,  562433024u,  562439168u,  549556224u,  549564416u,  536876032u

static const uint32_t kIrsqrt3R_host[96] = {PCC_IRSQRT_3R_VALUES};
static const uint32_t kIrsqrtR3_host[96] = {PCC_IRSQRT_R3_VALUES};
#if defined(__CUDACC__)
static __device__ __constant__ uint32_t kIrsqrt3R_dev[96] = {PCC_IRSQRT_3R_VALUES};
static __device__ __constant__ uint32_t kIrsqrtR3_dev[96] = {PCC_IRSQRT_R3_VALUES};
#endif

PCC_HD uint32_t
irsqrt_seed_3r(int idx)
{
#if defined(__CUDA_ARCH__)
  return kIrsqrt3R_dev[idx];
#else
  return kIrsqrt3R_host[idx];
#endif
}

PCC_HD uint32_t
irsqrt_seed_r3(int idx)
{
#if defined(__CUDA_ARCH__)
  return kIrsqrtR3_dev[idx];
#else
  return kIrsqrtR3_host[idx];
#endif
}

PCC_HD uint64_t
irsqrt64(uint64_t a64)
{
  if (!a64)
    return 0;

  // bring the argument into [2^30, 2^32) by an even shift; the result is
  // de-normalised by half of it
  int lz = clz64(a64);            // 0..63
  int msb = 63 - lz;              // position of the top bit
  int shift;                      // reference's `shift`, -3 at msb in {30,31}
  uint32_t a;
  if (msb >= 32) {
    int k = (msb - 30) >> 1;      // number of >>2 steps until the value fits
    a = uint32_t(a64 >> (2 * k));
    shift = -3 - k;
  } else {
    int k = (31 - msb) >> 1;      // number of <<2 steps
    a = uint32_t(a64) << (2 * k);
    shift = -3 + k;
  }

  int idx = int(a >> 25) - 32;
  uint64_t r = uint64_t(irsqrt_seed_3r(idx))
    - ((uint64_t(irsqrt_seed_r3(idx)) * a) >> 32);
  uint64_t ar = (r * a) >> 32;
  uint64_t s = 0x30000000u - ((r * ar) >> 32);
  r = (r * s) >> 32;
  return shift > 0 ? r << shift : r >> -shift;
}

PCC_HD uint32_t
isqrt64(uint64_t x)
{
  if (x <= (uint64_t(1) << 46))
    return uint32_t(1 + ((x * irsqrt64(x)) >> 40));
  uint64_t x0 = (x + 65536) >> 16;
  return uint32_t(1 + ((x0 * irsqrt64(x0)) >> 32));
}

//----------------------------------------------------------------------------
// quantisation

struct Quantizer {
  int32_t step;
  int32_t recip;

  PCC_HD int64_t quantize(int64_t x) const
  {
    constexpr int fracBits = 18 + kAttrShift;
    constexpr int64_t offset = (int64_t(1) << fracBits) / 3;
    return x >= 0 ? (x * recip + offset) >> fracBits
                  : -((offset - x * recip) >> fracBits);
  }
  PCC_HD int64_t scale(int64_t x) const { return x * step; }
};

// kQpStep / kQpStepRecip, tmc3/tables.cpp:478-481
PCC_TABLE(int32_t, kQpStep, 6, {161, 181, 203, 228, 256, 287})
PCC_TABLE(int32_t, kQpStepRecip, 6, {416825, 370767, 330586, 294337, 262144, 233829})

PCC_HD Quantizer
make_quantizer(int qp)
{
  qp = qp < 4 ? 4 : qp;
  int sh = qp / 6;
  int r = qp - 6 * sh;
  Quantizer q;
  q.step = kQpStep(r) << sh;
  q.recip = kQpStepRecip(r) >> sh;
  return q;
}

// the part of QpSet the kernels need for one transform layer
struct LayerQp {
  int32_t luma;
  int32_t chromaOffset;
  int32_t maxQp;
  int32_t fixedPointQpOffset;
};

PCC_HD void
make_quantizers(const LayerQp& l, int off0, int off1, Quantizer q[2])
{
  int qp0 = l.luma + off0;
  qp0 = qp0 < 4 ? 4 : (qp0 > l.maxQp ? l.maxQp : qp0);
  int qp1 = l.chromaOffset + off1 + qp0;
  qp1 = qp1 < 4 ? 4 : (qp1 > l.maxQp ? l.maxQp : qp1);
  q[0] = make_quantizer(qp0 + l.fixedPointQpOffset);
  q[1] = make_quantizer(qp1 + l.fixedPointQpOffset);
}

//----------------------------------------------------------------------------
// Morton arithmetic

PCC_HD uint64_t
morton3d_add(uint64_t a, uint64_t b)
{
  uint64_t mask = 0x9249249249249249ull;
  uint64_t val = 0;
  for (int i = 0; i < 3; i++) {
    val |= ((a | ~mask) + (b & mask)) & mask;
    mask <<= 1;
  }
  return val;
}

// spread the low 21 bits of v so that bit i lands at bit 3i
PCC_HD uint64_t
spread3(uint32_t v)
{
  uint64_t x = v & 0x1fffffu;
  x = (x | (x << 32)) & 0x001f00000000ffffull;
  x = (x | (x << 16)) & 0x001f0000ff0000ffull;
  x = (x | (x << 8)) & 0x100f00f00f00f00full;
  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

// x -> bit 2, y -> bit 1, z -> bit 0 of every triple.  The reference's LUT
// version consumes 24 bits per axis and lets everything above bit 63 fall
// off; that leaves 21 bits of x and y and 22 bits of z.
PCC_HD int64_t
morton_addr(int32_t x, int32_t y, int32_t z)
{
  uint64_t r = (spread3(uint32_t(x)) << 2) | (spread3(uint32_t(y)) << 1)
    | spread3(uint32_t(z));
  r |= uint64_t((uint32_t(z) >> 21) & 1) << 63;
  return int64_t(r);
}

//----------------------------------------------------------------------------
// divApprox: kDivApproxDivisor[i] + 1 == round(65536 / (i + 1))

PCC_HD int64_t
div_approx(int64_t a, uint64_t b, int log2Scale)
{
  int n = ilog2_u64(b) + 1 - 8;
  n = n < 0 ? 0 : n;
  uint32_t index = uint32_t((b + ((uint64_t(1) << n) >> 1)) >> n);
  int64_t invB = int64_t((65536u + index / 2) / index);
  return (invB * a) >> (n + 16 - log2Scale);
}

}  // namespace pccb200
