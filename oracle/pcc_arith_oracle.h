/* pcc_arith_oracle.h — TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Plain-C restatement of the integer / fixed-point helpers the TMC13
 * attribute-transform path is built on.  Every function names the reference
 * lines it follows (paths relative to the TMC13 tree).  Parity is PINNED:
 * tests/test_oracle_vs_reference.py checks each helper against the compiled
 * reference (oracle/_ref/libtmc13_ref.so) and against the committed golden
 * vectors in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * use anything under oracle/.
 */
#ifndef PCC_ARITH_ORACLE_H
#define PCC_ARITH_ORACLE_H

#include <stdint.h>

#define ORC_FRAC_BITS 15                 /* FixedPoint::kFracBits, FixedPoint.h:47 */
#define ORC_ONE_HALF (1 << (ORC_FRAC_BITS - 1))
#define ORC_ATTR_SHIFT 8                 /* kFixedPointAttributeShift, constants.h:46 */

/* FixedPoint::operator*= — Q.15 product, round half away from zero
 * (FixedPoint.h:113-122). */
static inline int64_t
orc_fxmul(int64_t a, int64_t b)
{
  int64_t v = a * b;
  if (v < 0)
    return -((ORC_ONE_HALF - v) >> ORC_FRAC_BITS);
  return (ORC_ONE_HALF + v) >> ORC_FRAC_BITS;
}

/* FixedPoint::round (FixedPoint.h:76-82). */
static inline int64_t
orc_fxround(int64_t v)
{
  if (v > 0)
    return (ORC_ONE_HALF + v) >> ORC_FRAC_BITS;
  return -((ORC_ONE_HALF - v) >> ORC_FRAC_BITS);
}

/* FixedPoint::operator=(int64) (FixedPoint.h:86-93): sign-magnitude shift. */
static inline int64_t
orc_fxfromint(int64_t v)
{
  if (v > 0)
    return v << ORC_FRAC_BITS;
  return -((-v) << ORC_FRAC_BITS);
}

/* divExp2RoundHalfUp (PCCMath.h:650-658). */
static inline int64_t
orc_div_exp2_half_up(int64_t x, int shift)
{
  if (!shift)
    return x;
  return (x + (1ll << (shift - 1))) >> shift;
}

/* divExp2RoundHalfInf, signed and unsigned (PCCMath.h:664-685). */
static inline int64_t
orc_div_exp2_half_inf(int64_t x, int shift)
{
  if (!shift)
    return x;
  int64_t s0 = 1ll << (shift - 1);
  return x >= 0 ? (s0 + x) >> shift : -((s0 - x) >> shift);
}
static inline uint64_t
orc_div_exp2_half_inf_u(uint64_t x, int shift)
{
  if (!shift)
    return x;
  return ((1ull << (shift - 1)) + x) >> shift;
}

/* ilog2 (PCCMisc.h:149-165): floor(log2 x), ilog2(0) = -1. */
static inline int
orc_ilog2(uint64_t x)
{
  int r = -1;
  while (x) {
    r++;
    x >>= 1;
  }
  return r;
}

/* Inverse square root, 96-entry seed + two Newton steps (misc.cpp:140-225).
 * The two seed tables are normative constants of the G-PCC specification. */
static const uint64_t orc_k3timesR[96] = {
  3196059648u, 3145728000u, 3107979264u, 3057647616u, 3019898880u, 2969567232u,
  2931818496u, 2894069760u, 2868903936u, 2831155200u, 2793406464u, 2768240640u,
  2730491904u, 2705326080u, 2667577344u, 2642411520u, 2617245696u, 2592079872u,
  2566914048u, 2541748224u, 2516582400u, 2491416576u, 2466250752u, 2441084928u,
  2428502016u, 2403336192u, 2378170368u, 2365587456u, 2340421632u, 2327838720u,
  2302672896u, 2290089984u, 2264924160u, 2252341248u, 2239758336u, 2214592512u,
  2202009600u, 2189426688u, 2164260864u, 2151677952u, 2139095040u, 2126512128u,
  2113929216u, 2101346304u, 2088763392u, 2076180480u, 2051014656u, 2038431744u,
  2025848832u, 2013265920u, 2000683008u, 2000683008u, 1988100096u, 1962934272u,
  1962934272u, 1950351360u, 1937768448u, 1925185536u, 1912602624u, 1900019712u,
  1900019712u, 1887436800u, 1874853888u, 1862270976u, 1849688064u, 1849688064u,
  1837105152u, 1824522240u, 1811939328u, 1811939328u, 1799356416u, 1786773504u,
  1786773504u, 1774190592u, 1761607680u, 1761607680u, 1749024768u, 1736441856u,
  1736441856u, 1723858944u, 1723858944u, 1711276032u, 1698693120u, 1698693120u,
  1686110208u, 1686110208u, 1673527296u, 1660944384u, 1660944384u, 1648361472u,
  1648361472u, 1635778560u, 1635778560u, 1623195648u, 1623195648u, 1610612736u};

static const uint64_t orc_kRcubed[96] = {
  4195081216u, 3999986688u, 3857709056u, 3673323520u, 3538940928u, 3364924416u,
  3238224896u, 3114735616u, 3034196992u, 2915990528u, 2800922624u, 2725880832u,
  2615890944u, 2544223232u, 2439185408u, 2370818048u, 2303728640u, 2237913088u,
  2173355008u, 2110061568u, 2048008192u, 1987165184u, 1927563264u, 1869150208u,
  1840392192u, 1783783424u, 1728321536u, 1701024768u, 1647311872u, 1620883456u,
  1568898048u, 1543306240u, 1492993024u, 1468236800u, 1443762176u, 1395656704u,
  1372007424u, 1348605952u, 1302626304u, 1280060416u, 1257736192u, 1235650560u,
  1213861888u, 1192294400u, 1171008512u, 1149979648u, 1108673536u, 1088379904u,
  1068352512u, 1048567808u, 1029031936u, 1029036032u, 1009729536u, 971888640u,
  971882496u,  953319424u,  934993920u,  916897792u,  899011584u,  881389568u,
  881392640u,  864009216u,  846846976u,  829900800u,  813182976u,  813201408u,
  796721152u,  780459008u,  764412928u,  764417024u,  748601344u,  732995584u,
  733017088u,  717624320u,  702468096u,  702466048u,  687520768u,  672786432u,
  672787456u,  658258944u,  658256896u,  643947520u,  629854208u,  629862400u,
  615976960u,  615952384u,  602276864u,  588779520u,  588804096u,  575512576u,
  575526912u,  562433024u,  562439168u,  549556224u,  549564416u,  536876032u};

static inline uint64_t
orc_irsqrt(uint64_t a64)
{
  if (!a64)
    return 0;

  /* normalise into a 32-bit mantissa with its top two bits not both clear */
  int shift = -3;
  while (a64 >> 32) {
    a64 >>= 2;
    shift--;
  }
  uint32_t a = (uint32_t)a64;
  while (!(a >> 30)) {
    a <<= 2;
    shift++;
  }

  int idx = (int)(a >> 25) - 32;
  uint64_t r = orc_k3timesR[idx] - ((orc_kRcubed[idx] * a) >> 32);
  uint64_t ar = (r * a) >> 32;
  uint64_t s = 0x30000000u - ((r * ar) >> 32);
  r = (r * s) >> 32;
  return shift > 0 ? r << shift : r >> -shift;
}

/* isqrt (misc.cpp:138-147). */
static inline uint32_t
orc_isqrt(uint64_t x)
{
  if (x <= ((uint64_t)1 << 46))
    return (uint32_t)(1 + ((x * orc_irsqrt(x)) >> 40));
  uint64_t x0 = (x + 65536) >> 16;
  return (uint32_t)(1 + ((x0 * orc_irsqrt(x0)) >> 32));
}

/* Quantizer (quantization.cpp:46-52, quantization.h:79-102). */
typedef struct {
  int step;
  int recip;
} orc_quantizer;

static const int orc_kQpStep[6] = {161, 181, 203, 228, 256, 287};            /* tables.cpp:478 */
static const int orc_kQpStepRecip[6] = {416825, 370767, 330586, 294337, 262144, 233829}; /* tables.cpp:480 */

static inline orc_quantizer
orc_mkquant(int qp)
{
  if (qp < 4)
    qp = 4;
  int sh = qp / 6;
  orc_quantizer q;
  q.step = orc_kQpStep[qp % 6] << sh;
  q.recip = orc_kQpStepRecip[qp % 6] >> sh;
  return q;
}

static inline int64_t
orc_quantize(orc_quantizer q, int64_t x)
{
  const int fracBits = 18 + ORC_ATTR_SHIFT;
  const int64_t offset = (1ll << fracBits) / 3;
  if (x >= 0)
    return (x * q.recip + offset) >> fracBits;
  return -((offset - x * q.recip) >> fracBits);
}

static inline int64_t
orc_scale(orc_quantizer q, int64_t x)
{
  return x * q.step;
}

/* morton3dAdd (PCCMisc.h:244-256): per-axis add of two interleaved triples. */
static inline uint64_t
orc_morton3d_add(uint64_t a, uint64_t b)
{
  uint64_t mask = 0x9249249249249249ull;
  uint64_t val = 0;
  for (int i = 0; i < 3; i++) {
    val |= ((a | ~mask) + (b & mask)) & mask;
    mask <<= 1;
  }
  return val;
}

/* mortonAddr (PCCMath.h:605-626): x -> bit 2, y -> bit 1, z -> bit 0 of each
 * triple; 24 bits per axis are consumed (LUT version, restated as a loop). */
static inline int64_t
orc_morton_addr(int32_t x, int32_t y, int32_t z)
{
  uint64_t r = 0;
  for (int b = 0; b < 24; b++) {
    /* bits shifted beyond bit 63 fall off, as in the reference's << 24 */
    if (3 * b + 2 < 64)
      r |= (uint64_t)((x >> b) & 1) << (3 * b + 2);
    if (3 * b + 1 < 64)
      r |= (uint64_t)((y >> b) & 1) << (3 * b + 1);
    if (3 * b < 64)
      r |= (uint64_t)((z >> b) & 1) << (3 * b);
  }
  return (int64_t)r;
}

/* divApprox (PCCMath.h:714-736); kDivApproxDivisor[i] + 1 == round(65536 / (i + 1))
 * (misc.cpp:313-335, checked entry by entry in the tests). */
static inline int64_t
orc_div_approx(int64_t a, uint64_t b, int log2Scale)
{
  int n = orc_ilog2(b) + 1 - 8;
  if (n < 0)
    n = 0;
  uint64_t index = (b + ((1ull << n) >> 1)) >> n;
  int log2InvScale = n + 16;
  int64_t invB = (int64_t)((65536 + index / 2) / index);
  return (invB * a) >> (log2InvScale - log2Scale);
}

#endif
