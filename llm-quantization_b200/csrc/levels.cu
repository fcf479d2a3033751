// POT / APOT nearest-level rounding with the per-group scale grid search.
//   ref: pot_apot_quantizer.py:25-115 (POT), :192-351 (APOT)
//
// The reference evaluates n_grid candidate scales per group, each a full pass of div / log2 / round
// / pow (POT) or a 31-way nearest search (APOT) plus a row sum of squared errors, and keeps the
// first candidate with the strictly smallest error.  Here one 8-lane team owns one group, keeps it
// in registers and runs ALL candidates without touching memory again: HBM traffic is the
// algorithmic 2 x sizeof(T) bytes per element, the rest is FP32 issue slots.
//
// Bit-exactness with torch's CPU kernels needs three things, all reproduced literally:
//   1. rne(log2f(r)) is evaluated as a step function whose step positions come from the host table
//      (core.cu) -> no dependence on CUDA's log2f;
//   2. every elementwise op is a separately rounded IEEE op (-fmad=false, __fdiv_rn);
//   3. the row sum ((w - wq)**2).sum(dim=1) follows ATen's vectorised inner-sum order
//      (8-float vectors, 4 interleaved accumulators, cascade levels, lanes added 0..7) — see
//      torch_rowsum below; tests/test_torch_semantics.py pins that order against torch itself.
#include <mutex>

#include "common.cuh"

namespace b200q {

struct GridParam {
  float b[256];
};
struct LevelParam {
  float lv[32];
};

__constant__ uint32_t c_round_thr[255];  // index e+127
__constant__ uint32_t c_floor_thr[277];  // index e+149

static int ensure_tables_uploaded() {
  // once per device
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return fail(B200Q_ECUDA, "cudaGetDevice failed");
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return B200Q_OK;
  cudaError_t e = cudaMemcpyToSymbol(c_round_thr, log2_round_thresholds(), sizeof(uint32_t) * 255);
  if (e == cudaSuccess)
    e = cudaMemcpyToSymbol(c_floor_thr, log2_floor_thresholds(), sizeof(uint32_t) * 277);
  if (e != cudaSuccess) return fail(B200Q_ECUDA, std::string("table upload: ") + cudaGetErrorString(e));
  done[dev] = true;
  return B200Q_OK;
}

// floor(log2f(m)) with torch-CPU rounding behaviour, m > 0 finite
__device__ __forceinline__ int floor_log2_torch(float m) {
  const uint32_t bits = __float_as_uint(m);
  int e = (int)(bits >> 23) - 127;  // true floor(log2 m) for normal m
  if (e < -126) {
    // subnormal: exponent from the leading bit
    e = -149 + (31 - __clz(bits));
  }
  // log2f rounds up to exactly e+1 for the last few floats below 2^(e+1)
  if (e + 1 <= 127 && bits >= c_floor_thr[e + 1 + 149]) e += 1;
  return e;
}

__device__ __forceinline__ float pow2i(int e) {
  // torch.pow(2.0, e) for integral e: exact, with gradual underflow
  if (e >= -126) return __uint_as_float((uint32_t)(e + 127) << 23);
  if (e >= -149) return __uint_as_float(1u << (e + 149));
  return 0.f;
}

// -------------------------------------------------------------------------------------------------
// ATen's row sum over one contiguous row of G floats (SumKernel.cpp: vectorized_inner_sum ->
// row_sum -> multi_row_sum), executed by one warp.  Lane = (k, l): k = interleaved accumulator
// (ilp_factor 4), l = position inside the 8-float vector.  f(i) yields the i-th addend.
// All lanes return the final value.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ceil_log2_i(int x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }

template <typename F>
__device__ __forceinline__ float warp_torch_rowsum(int G, int lane, F f) {
  if (G < 8) {
    // shorter than one vector: ATen's scalar_inner_sum (4 interleaved scalar accumulators)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int t = 0;
    if (G >= 4) { a0 += f(0); a1 += f(1); a2 += f(2); a3 += f(3); t = 4; }
    for (; t < G; ++t) a0 += f(t);
    return ((a0 + a1) + a2) + a3;
  }
  const int vec_size = G >> 3;
  const int size_ilp = vec_size >> 2;
  const int k = lane >> 3, l = lane & 7;
  const int level_power = max(4, ceil_log2_i(size_ilp) / 4);
  const int level_step = 1 << level_power;
  const int level_mask = level_step - 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = 0;
  while (i + level_step <= size_ilp) {
    for (int j = 0; j < level_step; ++j, ++i) a0 += f(((i << 2) + k) * 8 + l);
    a1 += a0; a0 = 0.f;
    if ((i & (level_mask << level_power)) != 0) continue;
    a2 += a1; a1 = 0.f;
    if ((i & (level_mask << (2 * level_power))) != 0) continue;
    a3 += a2; a2 = 0.f;
  }
  for (; i < size_ilp; ++i) a0 += f(((i << 2) + k) * 8 + l);
  a0 += a1; a0 += a2; a0 += a3;
  // vectors left over after the 4-way interleave go to accumulator 0
  for (int vi = size_ilp << 2; vi < vec_size; ++vi)
    if (k == 0) a0 += f(vi * 8 + l);
  // partial_sums[0] += partial_sums[1..3]
  float p = a0;
  p += __shfl_sync(0xffffffffu, a0, l + 8);
  p += __shfl_sync(0xffffffffu, a0, l + 16);
  p += __shfl_sync(0xffffffffu, a0, l + 24);
  // scalar tail first, then the 8 lanes of the vector accumulator in order
  float fin = 0.f;
  for (int t = vec_size << 3; t < G; ++t) fin += f(t);
  if (vec_size > 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) fin += __shfl_sync(0xffffffffu, p, q);
  }
  return fin;
}

// -------------------------------------------------------------------------------------------------
// POT element evaluation
// -------------------------------------------------------------------------------------------------
struct PotConsts {
  int emax_idx;     // E_max_idx = 2^(b-1) - 1
  float tiny;       // finfo(dtype).tiny
  float ratio_min;  // 1e-10 in the tensor's dtype
  float bmin, bmax; // smallest / largest positive grid multiplier (for the fast-path range check)
};

// E = clamp(round(log2(clamp(|w| / s, 1e-10))), 0, E_max_idx)      pot_apot_quantizer.py:87-88
__device__ __forceinline__ int pot_exponent(float aw, float s, const PotConsts& c,
                                            const uint32_t* __restrict__ thr /* smem, [emax+1] */) {
  const float r = fmaxf(__fdiv_rn(aw, s), c.ratio_min);
  const uint32_t bits = __float_as_uint(r);
  int e = (int)(bits >> 23) - 127;
  e = min(max(e, 0), c.emax_idx);
  // thr[e] = first float whose rne(log2f) is e+1; thr[emax] = 0xffffffff
  return e + (bits >= thr[e] ? 1 : 0);
}

// w_q = s * sign(w) * 2^E                                            pot_apot_quantizer.py:91
__device__ __forceinline__ float pot_value(float w, float s, int E) {
  const float sg = (w > 0.f) ? s : ((w < 0.f) ? -s : 0.f * s);
  return sg * __uint_as_float((uint32_t)(E + 127) << 23);
}

__device__ __forceinline__ void pot_base_scale(float amax, const PotConsts& c, float& s0) {
  // e_min = floor(log2(clamp(max,1e-12))) - E_max_idx ; s_0 = clamp(2^e_min, tiny)   :62-71
  const float msafe = fmaxf(amax, 1e-12f);
  const int emin = floor_log2_torch(msafe) - c.emax_idx;
  s0 = fmaxf(pow2i(emin), c.tiny);
}

// G == 128, fp32: 8 lanes per group, lane l owns elements v*8+l (v = 0..15), i.e. exactly the lane
// of ATen's 8-float vector, so the sum order needs only an 8-lane shuffle chain per candidate.
__global__ void __launch_bounds__(256)
pot128_f32_kernel(const float* __restrict__ w, float* __restrict__ out, uint8_t* __restrict__ exps,
                  float* __restrict__ best_scale_out, int32_t* __restrict__ best_idx_out,
                  int64_t n_groups, PotConsts c, GridParam grid, int n_grid) {
  __shared__ uint32_t thr[128];
  // Fast path table, indexed by the BIASED exponent of the ratio r = |w| / s:
  //   lut[e].x = first bit pattern in that binade whose rne(log2f) rounds up (0xffffffff: never)
  //   lut[e].y = bit pattern of 2^E_low, the level magnitude multiplier when it does not
  // so that  2^E = as_float(lut.y + (bits(r) >= lut.x ? 1 << 23 : 0)).  Entry 0 (r == 0, i.e.
  // w == 0: sign(0) = 0 makes w_q = 0) holds multiplier 0.
  __shared__ uint2 lut[256];
  for (int i = threadIdx.x; i <= c.emax_idx; i += blockDim.x)
    thr[i] = (i < c.emax_idx) ? c_round_thr[i + 127] : 0xffffffffu;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const int e = i - 127;
    uint2 t;
    if (i == 0) t = make_uint2(0xffffffffu, 0u);
    else if (e < 0) t = make_uint2(0xffffffffu, 0x3f800000u);
    else if (e >= c.emax_idx) t = make_uint2(0xffffffffu, (uint32_t)(c.emax_idx + 127) << 23);
    else t = make_uint2(c_round_thr[e + 127], (uint32_t)(e + 127) << 23);
    lut[i] = t;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int l = lane & 7;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = g < n_groups;
  const int64_t gg = valid ? g : 0;
  const float* wp = w + gg * 128 + l;
  float x[16];
#pragma unroll
  for (int v = 0; v < 16; ++v) x[v] = wp[v * 8];

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
  float s0;
  pot_base_scale(amax, c, s0);

  float best_err = INFINITY;
  float best_scale = s0;
  int best_idx = -1;
  // The LUT path needs the ratio to be a normal float computed by the exact reused-divisor
  // division: every non-zero |w| and every candidate scale inside [1e-18, 1e18] (any sane weight
  // group; NaN fails the comparisons).  Other groups take the general path below.
  const bool fast = (amax < 1e18f) && (amin_nz > 1e-18f) && (s0 * c.bmin > 1e-18f) &&
                    (s0 * c.bmax < 1e18f) && (c.bmin > 0.f);
  if (fast) {
    for (int ci = 0; ci < n_grid; ++ci) {
      const float s = fmaxf(s0 * grid.b[ci], c.tiny);                     // :81-82
      const Divisor sd(s);
      float acc[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = 0.f;
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const float aw = fabsf(x[v]);
        const uint32_t rb = __float_as_uint(sd.div_core(aw));             // r = |w| / s_b   :87
        const uint2 t = lut[rb >> 23];
        const float p2 = __uint_as_float(t.y + (rb >= t.x ? 0x00800000u : 0u));   // 2^E     :88
        const float d = aw - s * p2;         // |w - s*sign(w)*2^E| = | |w| - s*2^E |       :91,94
        acc[v & 3] += d * d;
      }
      float p = ((acc[0] + acc[1]) + acc[2]) + acc[3];
      float err = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) err += __shfl_sync(0xffffffffu, p, q, 8);
      if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }   // :97-99
    }
  } else
  for (int ci = 0; ci < n_grid; ++ci) {
    // s_b = clamp(s_0 * b, tiny)                                       :81-82
    const float s = fmaxf(s0 * grid.b[ci], c.tiny);
    float acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = 0.f;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int E = pot_exponent(fabsf(x[v]), s, c, thr);
      const float d = x[v] - pot_value(x[v], s, E);
      acc[v & 3] += d * d;   // vector v feeds accumulator v % 4, in order of v
    }
    float p = ((acc[0] + acc[1]) + acc[2]) + acc[3];
    float err = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) err += __shfl_sync(0xffffffffu, p, q, 8);
    // mask = error < best_error (strict: first minimum wins)           :97-99
    if (err < best_err) {
      best_err = err;
      best_scale = s;
      best_idx = ci;
    }
  }
  // final pass with the best scale                                     :103-107
  best_scale = fmaxf(best_scale, c.tiny);
  if (valid) {
    float* op = out + g * 128 + l;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int E = pot_exponent(fabsf(x[v]), best_scale, c, thr);
      op[v * 8] = pot_value(x[v], best_scale, E);
      if (exps != nullptr) exps[g * 128 + v * 8 + l] = (uint8_t)E;
    }
    if (l == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

// any group length, fp32: one warp per group, elements re-read through L1/L2 per candidate.
__global__ void __launch_bounds__(256)
pot_generic_f32_kernel(const float* __restrict__ w, float* __restrict__ out,
                       uint8_t* __restrict__ exps, float* __restrict__ best_scale_out,
                       int32_t* __restrict__ best_idx_out, int64_t n_groups, int G, PotConsts c,
                       GridParam grid, int n_grid) {
  __shared__ uint32_t thr[128];
  for (int i = threadIdx.x; i <= c.emax_idx; i += blockDim.x)
    thr[i] = (i < c.emax_idx) ? c_round_thr[i + 127] : 0xffffffffu;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const float* wp = w + g * (int64_t)G;
    float amax = 0.f;
    for (int i = lane; i < G; i += 32) amax = fmaxf(amax, fabsf(wp[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    float s0;
    pot_base_scale(amax, c, s0);
    float best_err = INFINITY, best_scale = s0;
    int best_idx = -1;
    for (int ci = 0; ci < n_grid; ++ci) {
      const float s = fmaxf(s0 * grid.b[ci], c.tiny);
      const float err = warp_torch_rowsum(G, lane, [&](int i) {
        const float x = wp[i];
        const float d = x - pot_value(x, s, pot_exponent(fabsf(x), s, c, thr));
        return d * d;
      });
      if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }
    }
    best_scale = fmaxf(best_scale, c.tiny);
    for (int i = lane; i < G; i += 32) {
      const float x = wp[i];
      const int E = pot_exponent(fabsf(x), best_scale, c, thr);
      out[g * (int64_t)G + i] = pot_value(x, best_scale, E);
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
// Levels are sorted ascending, so the minimum is one of the two levels bracketing x; the literal
// fp32 distances to those two decide, ties keep the lower index exactly like torch.argmin.  (A
// farther level can only tie a bracketing one if the level spacing is below one ulp of |x| <= 101;
// the host rejects such level sets for this kernel and uses the exhaustive variant.)
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
    // largest i with lv[i] <= x, by bisection over <= 32 entries
    int lo = 0, hi = L - 1;  // invariant: answer in [lo-1 .. hi]
    if (!(x >= lv[0])) return 0;
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

template <bool EXHAUSTIVE>
__global__ void __launch_bounds__(256)
apot128_f32_kernel(const float* __restrict__ w, float* __restrict__ out,
                   uint8_t* __restrict__ lidx, float* __restrict__ best_scale_out,
                   int32_t* __restrict__ best_idx_out, int64_t n_groups, LevelParam levels, int L,
                   GridParam grid, int n_grid) {
  __shared__ float lv[32];
  if (threadIdx.x < 32) lv[threadIdx.x] = levels.lv[min((int)threadIdx.x, L - 1)];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int l = lane & 7;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = g < n_groups;
  const int64_t gg = valid ? g : 0;
  const float* wp = w + gg * 128 + l;
  float x[16];
#pragma unroll
  for (int v = 0; v < 16; ++v) x[v] = wp[v * 8];
  float amax = 0.f;
#pragma unroll
  for (int v = 0; v < 16; ++v) amax = fmaxf(amax, fabsf(x[v]));
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  // s_0 = clamp(max|w|, 1e-5)                                          :250-251
  const float s0 = fmaxf(amax, 1e-5f);
  float best_err = INFINITY, best_scale = s0;
  int best_idx = -1;
  for (int ci = 0; ci < n_grid; ++ci) {
    const float s = s0 * grid.b[ci];  // :281
    float acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = 0.f;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int idx = apot_nearest<EXHAUSTIVE>(__fdiv_rn(x[v], s), lv, L);  // :284,294-297
      const float d = x[v] - s * lv[idx];                                   // :304,307
      acc[v & 3] += d * d;
    }
    float p = ((acc[0] + acc[1]) + acc[2]) + acc[3];
    float err = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) err += __shfl_sync(0xffffffffu, p, q, 8);
    if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }  // :310-312
  }
  if (valid) {
    float* op = out + g * 128 + l;
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int idx = apot_nearest<EXHAUSTIVE>(__fdiv_rn(x[v], best_scale), lv, L);  // :323-335
      op[v * 8] = best_scale * lv[idx];                                               // :340
      if (lidx != nullptr) lidx[g * 128 + v * 8 + l] = (uint8_t)idx;
    }
    if (l == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
  }
}

template <bool EXHAUSTIVE>
__global__ void __launch_bounds__(256)
apot_generic_f32_kernel(const float* __restrict__ w, float* __restrict__ out,
                        uint8_t* __restrict__ lidx, float* __restrict__ best_scale_out,
                        int32_t* __restrict__ best_idx_out, int64_t n_groups, int G,
                        LevelParam levels, int L, GridParam grid, int n_grid) {
  __shared__ float lv[32];
  if (threadIdx.x < 32) lv[threadIdx.x] = levels.lv[min((int)threadIdx.x, L - 1)];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const float* wp = w + g * (int64_t)G;
    float amax = 0.f;
    for (int i = lane; i < G; i += 32) amax = fmaxf(amax, fabsf(wp[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float s0 = fmaxf(amax, 1e-5f);
    float best_err = INFINITY, best_scale = s0;
    int best_idx = -1;
    for (int ci = 0; ci < n_grid; ++ci) {
      const float s = s0 * grid.b[ci];
      const float err = warp_torch_rowsum(G, lane, [&](int i) {
        const float x = wp[i];
        const float d = x - s * lv[apot_nearest<EXHAUSTIVE>(__fdiv_rn(x, s), lv, L)];
        return d * d;
      });
      if (err < best_err) { best_err = err; best_scale = s; best_idx = ci; }
    }
    for (int i = lane; i < G; i += 32) {
      const float x = wp[i];
      const int idx = apot_nearest<EXHAUSTIVE>(__fdiv_rn(x, best_scale), lv, L);
      out[g * (int64_t)G + i] = best_scale * lv[idx];
      if (lidx != nullptr) lidx[g * (int64_t)G + i] = (uint8_t)idx;
    }
    if (lane == 0) {
      if (best_scale_out != nullptr) best_scale_out[g] = best_scale;
      if (best_idx_out != nullptr) best_idx_out[g] = best_idx;
    }
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
  if (dtype != B200Q_F32)
    return fail(B200Q_EUNSUPPORTED, "pot_quant: only fp32 weights are implemented");
  if (n_groups == 0) return B200Q_OK;
  int rc = ensure_tables_uploaded();
  if (rc != B200Q_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("pot_quant", 2.0 * n_groups * group * elem_size(dtype), 0, st);
  PotConsts c;
  c.emax_idx = (1 << (n_bit - 1)) - 1;
  c.tiny = 1.17549435e-38f;
  c.ratio_min = 1e-10f;
  c.bmin = grid_host[0];
  c.bmax = grid_host[0];
  for (int i = 1; i < n_grid; ++i) {
    c.bmin = fminf(c.bmin, grid_host[i]);
    c.bmax = fmaxf(c.bmax, grid_host[i]);
  }
  GridParam gp;
  for (int i = 0; i < 256; ++i) gp.b[i] = i < n_grid ? grid_host[i] : 0.f;
  const float* wf = static_cast<const float*>(w);
  float* of = static_cast<float*>(out);
  if (group == 128 && (reinterpret_cast<uintptr_t>(w) % 4 == 0)) {
    const int64_t blocks = (n_groups + 31) / 32;  // 256 threads = 32 groups
    pot128_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(wf, of, exps, best_scale, best_idx,
                                                        n_groups, c, gp, n_grid);
  } else {
    int64_t blocks = std::min<int64_t>((n_groups + 7) / 8, (int64_t)kNumSMs * 32);
    pot_generic_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(wf, of, exps, best_scale, best_idx,
                                                             n_groups, (int)group, c, gp, n_grid);
  }
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
  if (dtype != B200Q_F32)
    return fail(B200Q_EUNSUPPORTED, "apot_quant: only fp32 weights are implemented");
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
  GridParam gp;
  for (int i = 0; i < 256; ++i) gp.b[i] = i < n_grid ? grid_host[i] : 0.f;
  const float* wf = static_cast<const float*>(w);
  float* of = static_cast<float*>(out);
  if (group == 128) {
    const int64_t blocks = (n_groups + 31) / 32;
    if (exhaustive)
      apot128_f32_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(wf, of, level_idx, best_scale,
                                                                 best_idx, n_groups, lp, n_levels,
                                                                 gp, n_grid);
    else
      apot128_f32_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(wf, of, level_idx, best_scale,
                                                                  best_idx, n_groups, lp, n_levels,
                                                                  gp, n_grid);
  } else {
    int64_t blocks = std::min<int64_t>((n_groups + 7) / 8, (int64_t)kNumSMs * 32);
    if (exhaustive)
      apot_generic_f32_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(
          wf, of, level_idx, best_scale, best_idx, n_groups, (int)group, lp, n_levels, gp, n_grid);
    else
      apot_generic_f32_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(
          wf, of, level_idx, best_scale, best_idx, n_groups, (int)group, lp, n_levels, gp, n_grid);
  }
  count_launch();
  return check_launch("apot_quant");
}

}  // extern "C"
