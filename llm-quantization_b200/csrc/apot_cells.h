// Cell table for the APOT nearest-level search (pot_apot_quantizer.py:294-297).
//
// The reference picks, for every element x = w / s_b, the level with the smallest fp32 distance
// |x - level_l| (first minimum).  For sorted levels that index is a STEP FUNCTION of x: level i+1
// takes over from level i at one fp32 value thr[i] (the first x whose rounded distance to level
// i+1 is strictly smaller), so   idx(x) = #{ i : thr[i] <= x }.
// Instead of bisecting 31 thresholds per element and candidate scale, the kernel cuts the level
// range into kApotCells uniform cells; a cell remembers how many thresholds lie in lower cells and
// the one threshold inside it (level sets with two thresholds in one cell are not eligible and
// keep the bisecting kernel).  cell_of() is monotone in x, which is all the argument needs:
//   cell_of(thr) < cell_of(x)  =>  thr < x,      cell_of(thr) > cell_of(x)  =>  thr > x.
// One shared-memory read then yields the threshold and the two candidate levels.
//
// Plain C++: nvcc compiles it into levels.cu (host table builder + device lookup), g++ compiles it
// into tests/native/apot_cells_check.cpp, which checks the lookup against the literal argmin.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define APOT_HD __host__ __device__ __forceinline__
#else
#define APOT_HD inline
#endif

namespace b200q {

constexpr int kApotCells = 512;
constexpr int kApotMaxThr = 31;                                  // <= 32 levels
constexpr float kApotMagic = 12582912.f + kApotCells / 2;        // 1.5 * 2^23 + C / 2

struct ApotCells {
  float thr[kApotMaxThr];     // ascending; entries >= n_thr unused
  int32_t cell[kApotMaxThr];  // cell_of(thr[i]), strictly increasing; INT32_MAX beyond n_thr
  float R;                    // max |level|
  float scale;                // (C/2 - 1) / R
  int32_t n_thr;
};

APOT_HD uint32_t apot_float_bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t b;
  memcpy(&b, &x, 4);
  return b;
#endif
}

// Monotone non-decreasing in x (clamp, one correctly rounded fma, integer read off the mantissa);
// NaN lands in the lowest cell, where the comparison against any threshold is false (index 0,
// like the reference's argmin over NaN distances).  Result in [1, kApotCells - 1].
APOT_HD int apot_cell_of(float x, float R, float scale) {
#if defined(__CUDA_ARCH__)
  const float t = fminf(fmaxf(x, -R), R);                  // FMNMX: any NaN -> -R
#else
  const float t = !(x > -R) ? -R : (x < R ? x : R);        // the same, whatever libm does with sNaN
#endif
  const float u = fmaf(t, scale, kApotMagic);
  return (int)(apot_float_bits(u) & (uint32_t)(kApotCells - 1));
}

// what cell c stores: base = thresholds in lower cells, thr = the threshold inside (NaN: none)
APOT_HD void apot_cell_entry(const ApotCells& t, int c, int& base, float& thr) {
  base = 0;
  thr = NAN;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int i = 0; i < kApotMaxThr; ++i) {
    base += (t.cell[i] < c) ? 1 : 0;
    if (t.cell[i] == c) thr = t.thr[i];
  }
}

// index of the nearest level through the table (what the kernel evaluates per element)
APOT_HD int apot_lookup(const ApotCells& t, float x) {
  int base;
  float thr;
  apot_cell_entry(t, apot_cell_of(x, t.R, t.scale), base, thr);
  return base + ((x >= thr) ? 1 : 0);
}

// -------------------------------------------------------------------------------------------------
// host only: thresholds by bisection over the ordered fp32 bit patterns
// -------------------------------------------------------------------------------------------------
inline int32_t apot_key(float x) {          // order-preserving; -0.0 and +0.0 share key 0
  const uint32_t b = apot_float_bits(x);
  return (b & 0x80000000u) ? -(int32_t)(b & 0x7fffffffu) : (int32_t)b;
}
inline float apot_unkey(int32_t k) {
  const uint32_t b = k >= 0 ? (uint32_t)k : (0x80000000u | (uint32_t)(-k));
  float x;
  memcpy(&x, &b, 4);
  return x;
}
// the upper bracketing level wins iff its rounded distance is STRICTLY smaller (argmin keeps the
// first minimum); both distances are single fp32 subtractions, monotone in x
inline bool apot_upper_wins(float x, float lo, float hi) {
  const volatile float d0 = x - lo, d1 = x - hi;
  return fabsf(d1) < fabsf(d0);
}

// false: the level set is not eligible (unsorted, too close, two thresholds in one cell, ...)
inline bool apot_build_cells(const float* lv, int n_levels, ApotCells& t) {
  memset(&t, 0, sizeof(t));
  for (int i = 0; i < kApotMaxThr; ++i) t.cell[i] = INT32_MAX;
  if (n_levels < 1 || n_levels > kApotMaxThr + 1) return false;
  float R = 0.f;
  for (int i = 0; i < n_levels; ++i) {
    if (!isfinite(lv[i])) return false;
    if (i > 0 && !(lv[i] > lv[i - 1])) return false;
    R = fmaxf(R, fabsf(lv[i]));
  }
  if (!(R > 1e-30f) || !(R < 1e30f)) return false;
  t.R = R;
  t.scale = (float)(kApotCells / 2 - 1) / R;
  t.n_thr = n_levels - 1;
  for (int i = 0; i + 1 < n_levels; ++i) {
    const float lo = lv[i], hi = lv[i + 1];
    if (apot_upper_wins(lo, lo, hi) || !apot_upper_wins(hi, lo, hi)) return false;
    int64_t a = apot_key(lo), b = apot_key(hi);          // pred(a) false, pred(b) true
    while (b - a > 1) {
      const int64_t m = a + (b - a) / 2;
      if (apot_upper_wins(apot_unkey((int32_t)m), lo, hi)) b = m; else a = m;
    }
    t.thr[i] = apot_unkey((int32_t)b);
    t.cell[i] = apot_cell_of(t.thr[i], t.R, t.scale);
    if (i > 0 && !(t.cell[i] > t.cell[i - 1])) return false;
  }
  return true;
}

}  // namespace b200q
