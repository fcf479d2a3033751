// POT / APOT nearest-level rounding with the per-group scale grid search, fp32 / fp16 / bf16.
//   ref: pot_apot_quantizer.py:25-115 (POT), :192-351 (APOT)
//
// The reference evaluates n_grid candidate scales per group, each a full pass of div / log2 / round
// / pow (POT) or a 31-way nearest search (APOT) plus a row sum of squared errors, and keeps the
// first candidate with the strictly smallest error.  Here one 8-lane team owns one group, keeps it
// in registers and runs ALL candidates without touching memory again: HBM traffic is the
// algorithmic 2 x sizeof(T) bytes per element, the rest is FP32 issue slots.  The APOT nearest-level
// search inside the candidate loop is a table lookup (apot_cells.h) instead of a bisection.
//
// Bit-exactness with torch's CPU kernels needs three things, all reproduced literally:
//   1. rne(log2(r)) / floor(log2(m)) are evaluated as step functions whose step positions come
//      from host tables (core.cu; one set per dtype, because for fp16/bf16 torch rounds log2's
//      result to the tensor type first) -> no dependence on CUDA's log2f;
//   2. every elementwise op is a separately rounded IEEE op in fp32, then rounded to the tensor
//      type as torch does (-fmad=false, exact division, ST<T>::rnd);
//   3. the row sum ((w - wq)**2).sum(dim=1) follows ATen's vectorised inner-sum order: 8-lane fp32
//      vectors, 4 interleaved accumulators, cascade levels, lanes added 0..7.  A vector is 8 floats
//      for fp32 input and 16 halves (low 8 + high 8, added lane-wise on load) for fp16/bf16 input;
//      rows shorter than one vector use 4 interleaved scalar accumulators.
//      tests/test_torch_semantics.py pins these orders against torch itself.
#include <cstdlib>
#include <mutex>

#include "apot_cells.h"
#include "common.cuh"

namespace b200q {

struct GridParam {
  float b[256];
};
struct LevelParam {
  float lv[32];
};

__constant__ uint32_t c_round_thr[3][255];  // [dtype][e+127]
__constant__ uint32_t c_floor_thr[3][277];  // [dtype][e+149]

static int ensure_tables_uploaded() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return fail(B200Q_ECUDA, "cudaGetDevice failed");
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return B200Q_OK;
  for (int dt = 0; dt < 3; ++dt) {
    cudaError_t e = cudaMemcpyToSymbol(c_round_thr, log2_round_thresholds(dt), sizeof(uint32_t) * 255,
                                       sizeof(uint32_t) * 255 * dt);
    if (e == cudaSuccess)
      e = cudaMemcpyToSymbol(c_floor_thr, log2_floor_thresholds(dt), sizeof(uint32_t) * 277,
                             sizeof(uint32_t) * 277 * dt);
    if (e != cudaSuccess)
      return fail(B200Q_ECUDA, std::string("table upload: ") + cudaGetErrorString(e));
  }
  done[dev] = true;
  return B200Q_OK;
}

// per-dtype constants of the level kernels
template <typename T>
struct LT;
template <>
struct LT<float> {
  static constexpr int DT = B200Q_F32;
  static constexpr int EPV = 8;                    // elements per ATen load vector
  static __device__ __forceinline__ float tiny() { return 1.17549435e-38f; }
};
template <>
struct LT<__half> {
  static constexpr int DT = B200Q_F16;
  static constexpr int EPV = 16;
  static __device__ __forceinline__ float tiny() { return 6.103515625e-05f; }
};
template <>
struct LT<__nv_bfloat16> {
  static constexpr int DT = B200Q_BF16;
  static constexpr int EPV = 16;
  static __device__ __forceinline__ float tiny() { return 1.17549435e-38f; }
};

// floor(log2(m)) as torch evaluates it on a tensor of type T; m > 0, finite, T-representable
template <typename T>
__device__ __forceinline__ int floor_log2_torch(float m) {
  const uint32_t bits = __float_as_uint(m);
  int e = (int)(bits >> 23) - 127;              // true floor(log2 m) for normal m
  if (e < -126) e = -149 + (31 - __clz(bits));  // fp32 subnormal
  // log2 (rounded to fp32, then to T) reaches e+1 for the last few values below 2^(e+1)
  if (e + 1 <= 127 && bits >= c_floor_thr[LT<T>::DT][e + 1 + 149]) e += 1;
  return e;
}

// torch.pow(2, e) on a tensor of type T, e integral
template <typename T>
__device__ __forceinline__ float pow2i(int e) {
  float p;
  if (e > 127) p = INFINITY;
  else if (e >= -126) p = __uint_as_float((uint32_t)(e + 127) << 23);
  else if (e >= -149) p = __uint_as_float(1u << (e + 149));
  else p = 0.f;
  return ST<T>::rnd(p);
}

// -------------------------------------------------------------------------------------------------
// ATen's row sum over one contiguous row of G values (SumKernel.cpp: vectorized_inner_sum ->
// row_sum -> multi_row_sum), executed by one warp.  Lane = (k, l): k = interleaved accumulator
// (ilp_factor 4), l = lane of the 8-float vector.  f(i) yields the i-th addend (fp32).
// EPV = 8: a vector is 8 consecutive elements.  EPV = 16 (fp16/bf16 input): a vector is the
// lane-wise sum of elements [16v, 16v+8) and [16v+8, 16v+16).  All lanes return the total.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ceil_log2_i(int x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }

template <int EPV, typename F>
__device__ __forceinline__ float warp_torch_rowsum(int G, int lane, F f) {
  if (G < EPV) {
    // shorter than one vector: ATen's scalar_inner_sum (4 interleaved scalar accumulators)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int t = 0;
    for (; t + 3 < G; t += 4) { a0 += f(t); a1 += f(t + 1); a2 += f(t + 2); a3 += f(t + 3); }
    for (; t < G; ++t) a0 += f(t);
    return ((a0 + a1) + a2) + a3;
  }
  auto vec = [&](int vi, int l) -> float {
    if constexpr (EPV == 8) return f(vi * 8 + l);
    else return f(vi * 16 + l) + f(vi * 16 + 8 + l);
  };
  const int vec_size = G / EPV;
  const int size_ilp = vec_size >> 2;
  const int k = lane >> 3, l = lane & 7;
  const int level_power = max(4, ceil_log2_i(size_ilp) / 4);
  const int level_step = 1 << level_power;
  const int level_mask = level_step - 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = 0;
  while (i + level_step <= size_ilp) {
    for (int j = 0; j < level_step; ++j, ++i) a0 += vec((i << 2) + k, l);
    a1 += a0; a0 = 0.f;
    if ((i & (level_mask << level_power)) != 0) continue;
    a2 += a1; a1 = 0.f;
    if ((i & (level_mask << (2 * level_power))) != 0) continue;
    a3 += a2; a2 = 0.f;
  }
  for (; i < size_ilp; ++i) a0 += vec((i << 2) + k, l);
  a0 += a1; a0 += a2; a0 += a3;
  // vectors left over after the 4-way interleave go to accumulator 0
  for (int vi = size_ilp << 2; vi < vec_size; ++vi)
    if (k == 0) a0 += vec(vi, l);
  // partial_sums[0] += partial_sums[1..3]
  float p = a0;
  p += __shfl_sync(0xffffffffu, a0, l + 8);
  p += __shfl_sync(0xffffffffu, a0, l + 16);
  p += __shfl_sync(0xffffffffu, a0, l + 24);
  // scalar tail first, then the 8 lanes of the vector accumulator in order
  float fin = 0.f;
  for (int t = vec_size * EPV; t < G; ++t) fin += f(t);
#pragma unroll
  for (int q = 0; q < 8; ++q) fin += __shfl_sync(0xffffffffu, p, q);
  return fin;
}

// The same order for G == 128 with an 8-lane team: lane l holds sq[v] = addend of element v*8+l.
template <typename T>
__device__ __forceinline__ float team128_rowsum(const float (&sq)[16]) {
  float a[4];
  if constexpr (LT<T>::EPV == 8) {
    // 16 vectors, accumulator k takes vectors k, k+4, k+8, k+12 in that order
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = ((sq[k] + sq[k + 4]) + sq[k + 8]) + sq[k + 12];
  } else {
    // 8 vectors of 16 elements: vector j = elements (2j)*8+l and (2j+1)*8+l added on load
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = (sq[2 * k] + sq[2 * k + 1]) + (sq[2 * (k + 4)] + sq[2 * (k + 4) + 1]);
  }
  const float p = ((a[0] + a[1]) + a[2]) + a[3];
  float err = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) err += __shfl_sync(0xffffffffu, p, q, 8);
  return ST<T>::rnd(err);   // the sum of a 16-bit tensor is returned in its dtype
}

// -------------------------------------------------------------------------------------------------
// POT element evaluation
// -------------------------------------------------------------------------------------------------
struct PotConsts {
  int emax_idx;      // E_max_idx = 2^(b-1) - 1
  float bmin, bmax;  // smallest / largest grid multiplier (fast-path range check)
};

// general path.  E = clamp(round(log2(clamp(|w| / s, 1e-10))), 0, E_max_idx)   pot:87-88
template <typename T>
__device__ __forceinline__ int pot_exponent(float aw, float s, int emax_idx,
                                            const uint32_t* __restrict__ thr /* smem [emax+1] */) {
  const float r = fmaxf(ST<T>::rnd(__fdiv_rn(aw, s)), ST<T>::rnd(1e-10f));
  const uint32_t bits = __float_as_uint(r);
  int e = (int)(bits >> 23) - 127;
  e = min(max(e, 0), emax_idx);
  // thr[e] = first value whose rne(log2) is e+1; thr[emax] = 0xffffffff
  return e + (bits >= thr[e] ? 1 : 0);
}

// w_q = s * sign(w) * 2^E                                            pot_apot_quantizer.py:91
template <typename T>
__device__ __forceinline__ float pot_value(float w, float s, int E) {
  const float sg = (w > 0.f) ? s : ((w < 0.f) ? -s : 0.f * s);
  return ST<T>::rnd(sg * pow2i<T>(E));
}

template <typename T>
__device__ __forceinline__ float pot_base_scale(float amax, int emax_idx) {
  // e_min = floor(log2(clamp(max, 1e-12))) - E_max_idx ; s_0 = clamp(2^e_min, tiny)   :62-71
  const float msafe = fmaxf(amax, ST<T>::rnd(1e-12f));
  if (!(msafe > 0.f)) return LT<T>::tiny();     // log2(0) = -inf -> 2^-inf = 0 -> clamp
  if (!(msafe < INFINITY)) return INFINITY;
  const int emin = floor_log2_torch<T>(msafe) - emax_idx;
  return fmaxf(pow2i<T>(emin), LT<T>::tiny());
}

// shared setup of the two exponent tables
template <typename T>
__device__ __forceinline__ void pot_tables(uint32_t* thr, uint2* lut, int emax_idx) {
  for (int i = threadIdx.x; i <= emax_idx; i += blockDim.x)
    thr[i] = (i < emax_idx) ? c_round_thr[LT<T>::DT][i + 127] : 0xffffffffu;
  if (lut != nullptr) {
    // indexed by the BIASED fp32 exponent of the ratio r = |w| / s:
    //   .x = first bit pattern of that binade whose rne(log2) rounds up (0xffffffff: never)
    //   .y = bit pattern of 2^E_low; 2^E = as_float(.y + (bits(r) >= .x ? 1 << 23 : 0))
    // fp32: entry 0 (r == 0 <=> w == 0, guaranteed by the fast-path guard) carries multiplier 0
    // because sign(0) = 0 makes w_q = 0; 16-bit kernels test w == 0 explicitly instead (their
    // ratio can underflow to 0 for w != 0, which torch maps to E = 0).
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      const int e = i - 127;
      uint2 t;
      if (i == 0 && LT<T>::DT == B200Q_F32) t = make_uint2(0xffffffffu, 0u);
      else if (e < 0) t = make_uint2(0xffffffffu, 0x3f800000u);
      else if (e >= emax_idx) t = make_uint2(0xffffffffu, (uint32_t)(emax_idx + 127) << 23);
      else t = make_uint2(c_round_thr[LT<T>::DT][e + 127], (uint32_t)(e + 127) << 23);
      lut[i] = t;
    }
  }
  __syncthreads();
}

// G == 128: 8 lanes per group, lane l owns elements v*8+l (v = 0..15), i.e. exactly the lane of
// ATen's 8-float vector, so the sum order needs only an 8-lane shuffle chain per candidate.
template <typename T>
__global__ void __launch_bounds__(256)
pot128_kernel(const T* __restrict__ w, T* __restrict__ out, uint8_t* __restrict__ exps,
              float* __restrict__ best_scale_out, int32_t* __restrict__ best_idx_out,
              int64_t n_groups, PotConsts c, GridParam grid, int n_grid) {
  __shared__ uint32_t thr[128];
  __shared__ uint2 lut[256];
  pot_tables<T>(thr, lut, c.emax_idx);

  const int lane = threadIdx.x & 31;
  const int l = lane & 7;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = g < n_groups;
  const int64_t gg = valid ? g : 0;
  const T* wp = w + gg * 128 + l;
  float x[16];
#pragma unroll
  for (int v = 0; v < 16; ++v) x[v] = to_f(wp[v * 8]);

  float amax = 0.f, amin_nz = INFINITY;
#pragma unroll
  for (int v = 0; v < 16; ++v) {
    const float a = fabsf(x[v]);
    amax = fmaxf(amax, a);
    amin_nz = fminf(amin_nz, a == 0.f ? INFINITY : a);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    amin_nz = fminf(amin_nz, __shfl_xor_sync(0xffffffffu, amin_nz, o));
  }
  const float s0 = pot_base_scale<T>(amax, c.emax_idx);
  const float tiny = LT<T>::tiny();

  float best_err = INFINITY;
  float best_scale = s0;
  int best_idx = -1;
  // The LUT path needs the ratio computed by the exact reused-divisor division: every non-zero
  // |w| and every candidate scale inside [1e-18, 1e18] (any sane weight group; NaN fails the
  // comparisons).  Other groups take the general path.
  const bool fast = (amax < 1e18f) && (amin_nz > 1e-18f) && (s0 * c.bmin > 1e-18f) &&
                    (s0 * c.bmax < 1e18f) && (c.bmin > 0.f);
  for (int ci = 0; ci < n_grid; ++ci) {
    const float s = fmaxf(ST<T>::rnd(s0 * grid.b[ci]), tiny);              // :81-82
    float sq[16];
    if (fast) {
      const Divisor sd(s);
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const float aw = fabsf(x[v]);
        const uint32_t rb = __float_as_uint(ST<T>::rnd(sd.div_core(aw)));  // r = |w| / s_b    :87
        const uint2 t = lut[rb >> 23];
        float p2 = __uint_as_float(t.y + (rb >= t.x ? 0x00800000u : 0u));   // 2^E             :88
        float m;
        if constexpr (LT<T>::DT == B200Q_F32) {
          m = s * p2;
        } else {
          p2 = ST<T>::rnd(p2);                       // torch.pow(2.0, E) lives in the tensor dtype
          m = (aw == 0.f) ? 0.f : ST<T>::rnd(s * p2);
        }
        const float d = ST<T>::rnd(aw - m);          // |w - s sign(w) 2^E| = ||w| - s 2^E|  :91,94
        sq[v] = ST<T>::rnd(d * d);
      }
    } else {
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const int E = pot_exponent<T>(fabsf(x[v]), s, c.emax_idx, thr);
        const float d = ST<T>::rnd(x[v] - pot_value<T>(x[v], s, E));
        sq[v] = ST<T>::rnd(d * d);
      }
    }
    const float err = team128_rowsum<T>(sq);
    if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }   // :97-99
  }
  // final pass with the best scale                                     :103-107
  best_scale = fmaxf(best_scale, tiny);
  if (valid) {
    T* op = out + g * 128 + l;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int E = pot_exponent<T>(fabsf(x[v]), best_scale, c.emax_idx, thr);
      op[v * 8] = from_f<T>(pot_value<T>(x[v], best_scale, E));
      if (exps != nullptr) exps[g * 128 + v * 8 + l] = (uint8_t)E;
    }
    if (l == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

// any group length: one warp per group, elements re-read through L1/L2 per candidate.
template <typename T>
__global__ void __launch_bounds__(256)
pot_generic_kernel(const T* __restrict__ w, T* __restrict__ out, uint8_t* __restrict__ exps,
                   float* __restrict__ best_scale_out, int32_t* __restrict__ best_idx_out,
                   int64_t n_groups, int G, PotConsts c, GridParam grid, int n_grid) {
  __shared__ uint32_t thr[128];
  pot_tables<T>(thr, nullptr, c.emax_idx);
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float tiny = LT<T>::tiny();
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const T* wp = w + g * (int64_t)G;
    float amax = 0.f;
    for (int i = lane; i < G; i += 32) amax = fmaxf(amax, fabsf(to_f(wp[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float s0 = pot_base_scale<T>(amax, c.emax_idx);
    float best_err = INFINITY, best_scale = s0;
    int best_idx = -1;
    for (int ci = 0; ci < n_grid; ++ci) {
      const float s = fmaxf(ST<T>::rnd(s0 * grid.b[ci]), tiny);
      const float err = ST<T>::rnd(warp_torch_rowsum<LT<T>::EPV>(G, lane, [&](int i) {
        const float xv = to_f(wp[i]);
        const float d = ST<T>::rnd(xv - pot_value<T>(xv, s, pot_exponent<T>(fabsf(xv), s, c.emax_idx, thr)));
        return ST<T>::rnd(d * d);
      }));
      if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }
    }
    best_scale = fmaxf(best_scale, tiny);
    for (int i = lane; i < G; i += 32) {
      const float xv = to_f(wp[i]);
      const int E = pot_exponent<T>(fabsf(xv), best_scale, c.emax_idx, thr);
      out[g * (int64_t)G + i] = from_f<T>(pot_value<T>(xv, best_scale, E));
      if (exps != nullptr) exps[g * (int64_t)G + i] = (uint8_t)E;
    }
    if (lane == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// APOT element evaluation
// -------------------------------------------------------------------------------------------------
// closest_idx = argmin_l |x - level_l| (first minimum)                 pot_apot_quantizer.py:294-297
// The distances are fp32 in every dtype (the level table is fp32 and promotes x).  Levels are
// sorted ascending, so the minimum is one of the two levels bracketing x; the literal fp32
// distances to those two decide, ties keep the lower index exactly like torch.argmin.  (A farther
// level can only tie a bracketing one if the level spacing is below one ulp of |x| <= 101; the
// host rejects such level sets for this kernel and uses the exhaustive variant.)
template <bool EXHAUSTIVE>
__device__ __forceinline__ int apot_nearest(float x, const float* __restrict__ lv, int L) {
  if constexpr (EXHAUSTIVE) {
    int best = 0;
    float bd = fabsf(x - lv[0]);
    for (int i = 1; i < L; ++i) {
      const float d = fabsf(x - lv[i]);
      if (d < bd) { bd = d; best = i; }
    }
    return best;
  } else {
    if (!(x >= lv[0])) return 0;      // also NaN: torch.argmin returns the first NaN distance
    int lo = 0, hi = L - 1;           // largest i with lv[i] <= x, by bisection over <= 32 entries
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo + hi + 1) >> 1;
      if (lv[mid] <= x) lo = mid; else hi = mid - 1;
    }
    const int up = min(lo + 1, L - 1);
    const float d0 = fabsf(x - lv[lo]);
    const float d1 = fabsf(x - lv[up]);
    return (d1 < d0) ? up : lo;
  }
}

// one element under scale s: index of the nearest level and the squared error, torch op by op
template <typename T, bool EXHAUSTIVE, bool FAST_DIV>
__device__ __forceinline__ float apot_eval(float wv, float s, const Divisor& sd,
                                           const float* __restrict__ lv, int L, int& idx) {
  const float xn = ST<T>::rnd(FAST_DIV ? sd.div_core(wv) : __fdiv_rn(wv, s));   // w / s_b     :284
  idx = apot_nearest<EXHAUSTIVE>(xn, lv, L);                                     // :294-297
  const float wq = ST<T>::rnd(s * ST<T>::rnd(lv[idx]));                          // :298,304
  const float d = ST<T>::rnd(wv - wq);
  return ST<T>::rnd(d * d);                                                      // :307
}

// The same through the cell table (apot_cells.h): one shared-memory read gives the threshold inside
// x's cell and the two levels it separates, already rounded to the tensor type.
template <typename T, bool FAST_DIV>
__device__ __forceinline__ float apot_eval_cells(float wv, float s, const Divisor& sd,
                                                 const float4* __restrict__ lut, float R, float scale) {
  const float xn = ST<T>::rnd(FAST_DIV ? sd.div_core(wv) : __fdiv_rn(wv, s));   // w / s_b     :284
  const float4 e = lut[apot_cell_of(xn, R, scale)];
  const float q = (xn >= e.x) ? e.z : e.y;                                       // :294-298
  const float wq = ST<T>::rnd(s * q);                                            // :304
  const float d = ST<T>::rnd(wv - wq);
  return ST<T>::rnd(d * d);                                                      // :307
}

template <typename T, bool EXHAUSTIVE, bool CELLS>
__global__ void __launch_bounds__(256)
apot128_kernel(const T* __restrict__ w, T* __restrict__ out, uint8_t* __restrict__ lidx,
               float* __restrict__ best_scale_out, int32_t* __restrict__ best_idx_out,
               int64_t n_groups, LevelParam levels, int L, GridParam grid, int n_grid, float bmin,
               float bmax, ApotCells cells) {
  __shared__ float lv[32];
  __shared__ float4 lut[CELLS ? kApotCells : 1];
  if (threadIdx.x < 32) lv[threadIdx.x] = levels.lv[min((int)threadIdx.x, L - 1)];
  __syncthreads();
  if constexpr (CELLS) {
    for (int c = threadIdx.x; c < kApotCells; c += blockDim.x) {
      int base;
      float thr;
      apot_cell_entry(cells, c, base, thr);
      lut[c] = make_float4(thr, ST<T>::rnd(lv[base]), ST<T>::rnd(lv[min(base + 1, L - 1)]), 0.f);
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int l = lane & 7;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = g < n_groups;
  const int64_t gg = valid ? g : 0;
  const T* wp = w + gg * 128 + l;
  float x[16];
#pragma unroll
  for (int v = 0; v < 16; ++v) x[v] = to_f(wp[v * 8]);
  float amax = 0.f, amin_nz = INFINITY;
#pragma unroll
  for (int v = 0; v < 16; ++v) {
    const float a = fabsf(x[v]);
    amax = fmaxf(amax, a);
    amin_nz = fminf(amin_nz, a == 0.f ? INFINITY : a);
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    amin_nz = fminf(amin_nz, __shfl_xor_sync(0xffffffffu, amin_nz, o));
  }
  // s_0 = clamp(max|w|, 1e-5)                                          :250-251
  const float s0 = fmaxf(amax, ST<T>::rnd(1e-5f));
  const bool fast = (amax < 1e18f) && (amin_nz > 1e-18f) && (s0 * bmin > 1e-18f) &&
                    (s0 * bmax < 1e18f) && (bmin > 0.f);
  float best_err = INFINITY, best_scale = s0;
  int best_idx = -1;
  for (int ci = 0; ci < n_grid; ++ci) {
    const float s = ST<T>::rnd(s0 * grid.b[ci]);                         // :281
    const Divisor sd(s);
    float sq[16];
    int idx;
    if constexpr (CELLS) {
      if (fast) {
#pragma unroll
        for (int v = 0; v < 16; ++v)
          sq[v] = apot_eval_cells<T, true>(x[v], s, sd, lut, cells.R, cells.scale);
      } else {
#pragma unroll
        for (int v = 0; v < 16; ++v)
          sq[v] = apot_eval_cells<T, false>(x[v], s, sd, lut, cells.R, cells.scale);
      }
    } else if (fast) {
#pragma unroll
      for (int v = 0; v < 16; ++v) sq[v] = apot_eval<T, EXHAUSTIVE, true>(x[v], s, sd, lv, L, idx);
    } else {
#pragma unroll
      for (int v = 0; v < 16; ++v) sq[v] = apot_eval<T, EXHAUSTIVE, false>(x[v], s, sd, lv, L, idx);
    }
    const float err = team128_rowsum<T>(sq);
    if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }  // :310-312
  }
  if (valid) {
    T* op = out + g * 128 + l;
    const Divisor sd(best_scale);
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      int idx;
      apot_eval<T, EXHAUSTIVE, false>(x[v], best_scale, sd, lv, L, idx);            // :323-335
      op[v * 8] = from_f<T>(ST<T>::rnd(best_scale * ST<T>::rnd(lv[idx])));            // :340
      if (lidx != nullptr) lidx[g * 128 + v * 8 + l] = (uint8_t)idx;
    }
    if (l == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

template <typename T, bool EXHAUSTIVE>
__global__ void __launch_bounds__(256)
apot_generic_kernel(const T* __restrict__ w, T* __restrict__ out, uint8_t* __restrict__ lidx,
                    float* __restrict__ best_scale_out, int32_t* __restrict__ best_idx_out,
                    int64_t n_groups, int G, LevelParam levels, int L, GridParam grid, int n_grid) {
  __shared__ float lv[32];
  if (threadIdx.x < 32) lv[threadIdx.x] = levels.lv[min((int)threadIdx.x, L - 1)];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const T* wp = w + g * (int64_t)G;
    float amax = 0.f;
    for (int i = lane; i < G; i += 32) amax = fmaxf(amax, fabsf(to_f(wp[i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float s0 = fmaxf(amax, ST<T>::rnd(1e-5f));
    float best_err = INFINITY, best_scale = s0;
    int best_idx = -1;
    const Divisor unused;
    for (int ci = 0; ci < n_grid; ++ci) {
      const float s = ST<T>::rnd(s0 * grid.b[ci]);
      const float err = ST<T>::rnd(warp_torch_rowsum<LT<T>::EPV>(G, lane, [&](int i) {
        int idx;
        return apot_eval<T, EXHAUSTIVE, false>(to_f(wp[i]), s, unused, lv, L, idx);
      }));
      if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }
    }
    for (int i = lane; i < G; i += 32) {
      int idx;
      apot_eval<T, EXHAUSTIVE, false>(to_f(wp[i]), best_scale, unused, lv, L, idx);
      out[g * (int64_t)G + i] = from_f<T>(ST<T>::rnd(best_scale * ST<T>::rnd(lv[idx])));
      if (lidx != nullptr) lidx[g * (int64_t)G + i] = (uint8_t)idx;
    }
    if (lane == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

static void grid_param(GridParam& gp, const float* grid_host, int n_grid, float& bmin, float& bmax) {
  bmin = bmax = grid_host[0];
  for (int i = 0; i < 256; ++i) {
    gp.b[i] = i < n_grid ? grid_host[i] : 0.f;
    if (i < n_grid) { bmin = fminf(bmin, grid_host[i]); bmax = fmaxf(bmax, grid_host[i]); }
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int b200q_pot_quant(const void* w, void* out, uint8_t* exps, float* best_scale, int32_t* best_idx,
                    int64_t n_groups, int64_t group, int n_bit, const float* grid_host, int n_grid,
                    int dtype, void* stream) {
  B200Q_REQUIRE(w && out && grid_host, "pot_quant: null pointer");
  B200Q_REQUIRE(n_groups >= 0 && group > 0 && group < (1ll << 30), "pot_quant: bad shape");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 8, "pot_quant: n_bit must be in [1,8]");
  B200Q_REQUIRE(n_grid >= 1 && n_grid <= 256, "pot_quant: n_grid must be in [1,256]");
  if (n_groups == 0) return B200Q_OK;
  int rc = ensure_tables_uploaded();
  if (rc != B200Q_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("pot_quant", 2.0 * n_groups * group * elem_size(dtype), 0, st);
  PotConsts c;
  c.emax_idx = (1 << (n_bit - 1)) - 1;
  GridParam gp;
  grid_param(gp, grid_host, n_grid, c.bmin, c.bmax);
  B200Q_DISPATCH_DTYPE(dtype, T, {
    const T* wt = static_cast<const T*>(w);
    T* ot = static_cast<T*>(out);
    if (group == 128) {
      const int64_t blocks = (n_groups + 31) / 32;  // 256 threads = 32 groups
      pot128_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(wt, ot, exps, best_scale, best_idx,
                                                         n_groups, c, gp, n_grid);
    } else {
      const int64_t blocks = std::min<int64_t>((n_groups + 7) / 8, (int64_t)kNumSMs * 32);
      pot_generic_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(wt, ot, exps, best_scale, best_idx,
                                                              n_groups, (int)group, c, gp, n_grid);
    }
  });
  count_launch();
  return check_launch("pot_quant");
}

int b200q_apot_quant(const void* w, void* out, uint8_t* level_idx, float* best_scale,
                     int32_t* best_idx, int64_t n_groups, int64_t group, const float* levels_host,
                     int n_levels, const float* grid_host, int n_grid, int dtype, void* stream) {
  B200Q_REQUIRE(w && out && grid_host && levels_host, "apot_quant: null pointer");
  B200Q_REQUIRE(n_groups >= 0 && group > 0 && group < (1ll << 30), "apot_quant: bad shape");
  B200Q_REQUIRE(n_levels >= 1 && n_levels <= 32, "apot_quant: n_levels must be in [1,32]");
  B200Q_REQUIRE(n_grid >= 1 && n_grid <= 256, "apot_quant: n_grid must be in [1,256]");
  if (n_groups == 0) return B200Q_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("apot_quant", 2.0 * n_groups * group * elem_size(dtype), 0, st);
  LevelParam lp;
  bool sorted = true;
  float min_gap = INFINITY;
  for (int i = 0; i < 32; ++i) lp.lv[i] = levels_host[i < n_levels ? i : n_levels - 1];
  for (int i = 1; i < n_levels; ++i) {
    const float gap = levels_host[i] - levels_host[i - 1];
    if (!(gap > 0.f)) sorted = false;
    min_gap = fminf(min_gap, gap);
  }
  // bracket search is exact only for strictly increasing levels spaced well above ulp(101)
  const bool exhaustive = !sorted || (n_levels > 1 && min_gap < 6.2e-5f);
  // G == 128: the candidate loop finds the nearest level through the cell table when the level
  // set allows it (apot_cells.h; every reference level set with k <= 2 does); B200Q_APOT_CELLS=0
  // keeps the bisecting kernel (A/B timing)
  const char* cells_env = getenv("B200Q_APOT_CELLS");       // read per call: A/B inside one process
  const bool cells_enabled = !(cells_env && cells_env[0] == '0');
  ApotCells cells;
  const bool use_cells = apot_build_cells(levels_host, n_levels, cells) && !exhaustive &&
                         cells_enabled && n_levels > 1;
  GridParam gp;
  float bmin, bmax;
  grid_param(gp, grid_host, n_grid, bmin, bmax);
  B200Q_DISPATCH_DTYPE(dtype, T, {
    const T* wt = static_cast<const T*>(w);
    T* ot = static_cast<T*>(out);
    if (group == 128) {
      const int64_t blocks = (n_groups + 31) / 32;
      if (exhaustive)
        apot128_kernel<T, true, false><<<(unsigned)blocks, 256, 0, st>>>(
            wt, ot, level_idx, best_scale, best_idx, n_groups, lp, n_levels, gp, n_grid, bmin, bmax,
            cells);
      else if (use_cells)
        apot128_kernel<T, false, true><<<(unsigned)blocks, 256, 0, st>>>(
            wt, ot, level_idx, best_scale, best_idx, n_groups, lp, n_levels, gp, n_grid, bmin, bmax,
            cells);
      else
        apot128_kernel<T, false, false><<<(unsigned)blocks, 256, 0, st>>>(
            wt, ot, level_idx, best_scale, best_idx, n_groups, lp, n_levels, gp, n_grid, bmin, bmax,
            cells);
    } else {
      const int64_t blocks = std::min<int64_t>((n_groups + 7) / 8, (int64_t)kNumSMs * 32);
      if (exhaustive)
        apot_generic_kernel<T, true><<<(unsigned)blocks, 256, 0, st>>>(
            wt, ot, level_idx, best_scale, best_idx, n_groups, (int)group, lp, n_levels, gp, n_grid);
      else
        apot_generic_kernel<T, false><<<(unsigned)blocks, 256, 0, st>>>(
            wt, ot, level_idx, best_scale, best_idx, n_groups, (int)group, lp, n_levels, gp, n_grid);
    }
  });
  count_launch();
  return check_launch("apot_quant");
}

}  // extern "C"
