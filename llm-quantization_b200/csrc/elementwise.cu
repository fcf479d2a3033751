// HBM-bound stages: column |max|, GPTQ parity column quantisation, uniform group fake-quant
// (pseudo_quantize_tensor / _simple_quantize_layer, with AWQ and SmoothQuant column ops fused),
// SmoothQuant scale + migration, activation statistics.
//
// Design: every kernel moves 128 bits per thread per access (ld.global.nc.L1::no_allocate /
// st.global.L1::no_allocate), keeps >=4 independent loads in flight per thread, and is launched
// with a grid that is a multiple of the 148 SMs where the shape allows.  Arithmetic follows torch's
// eager op sequence exactly (IEEE div, rint = half-to-even, no FMA contraction: this file is
// compiled with -fmad=false), so fp32 results are bit-identical to the reference's.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace b200q {

// =================================================================================================
// column abs-max  (ref: gptq_quantizer.py:182, smooth_quant_quantizer.py:156,68)
// =================================================================================================
// block = 32 column-lanes x 8 row-lanes; a warp reads 512 contiguous bytes of one row.
template <typename T, bool ABS_ONLY>
__global__ void __launch_bounds__(256)
col_absmax_kernel(const T* __restrict__ W, int64_t N, int64_t K, int64_t ld, int rows_per_block,
                  unsigned int* __restrict__ colmax_bits) {
  constexpr int VEC = ST<T>::VEC;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col0 = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(N, r0 + (int64_t)rows_per_block);
  float m[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) m[j] = 0.f;
  if (col0 < K) {
    const T* p = W + col0;
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float a[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_vec<T>(p + (r + 8 * u) * ld, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) m[j] = fmaxf(m[j], fabsf(a[u][j]));
    }
    for (; r < r1; r += 8) {
      float a[VEC];
      load_vec<T>(p + r * ld, a);
#pragma unroll
      for (int j = 0; j < VEC; ++j) m[j] = fmaxf(m[j], fabsf(a[j]));
    }
  }
  __shared__ float sm[8][32 * VEC + 1];
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[ty][tx * VEC + j] = m[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int64_t col = (int64_t)blockIdx.x * 32 * VEC + c;
    if (col < K) {
      float v = sm[0][c];
#pragma unroll
      for (int y = 1; y < 8; ++y) v = fmaxf(v, sm[y][c]);
      // non-negative floats order like their bit patterns
      atomicMax(colmax_bits + col, __float_as_uint(v));
    }
  }
}

// any alignment / leading dimension: one thread per column, rows chunked over blockIdx.y
template <typename T>
__global__ void __launch_bounds__(256)
col_absmax_scalar_kernel(const T* __restrict__ W, int64_t N, int64_t K, int64_t ld,
                         int rows_per_block, unsigned int* __restrict__ colmax_bits) {
  const int64_t col = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (col >= K) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(N, r0 + (int64_t)rows_per_block);
  float m = 0.f;
  for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(to_f(W[r * ld + col])));
  atomicMax(colmax_bits + col, __float_as_uint(m));
}

static int pick_rows_per_block(int64_t N, int64_t col_tiles) {
  // aim for ~8 CTAs per SM worth of blocks, at least 32 rows each
  const int64_t want_blocks = (int64_t)kNumSMs * 8;
  int64_t chunks = std::max<int64_t>(1, want_blocks / std::max<int64_t>(1, col_tiles));
  int64_t rpb = (N + chunks - 1) / chunks;
  rpb = std::max<int64_t>(32, (rpb + 31) / 32 * 32);
  return (int)std::min<int64_t>(rpb, 1 << 20);
}

template <typename T>
static int launch_col_absmax(const void* Wv, int64_t N, int64_t K, int64_t ld, float* colmax,
                             int accumulate, cudaStream_t st) {
  constexpr int VEC = ST<T>::VEC;
  const T* W = static_cast<const T*>(Wv);
  if (!accumulate) cudaMemsetAsync(colmax, 0, sizeof(float) * K, st);
  const bool vec_ok = aligned16(W) && (K % VEC == 0) && (ld % VEC == 0);
  if (vec_ok) {
    const int64_t col_tiles = (K + 32 * VEC - 1) / (32 * VEC);
    const int rpb = pick_rows_per_block(N, col_tiles);
    dim3 grid((unsigned)col_tiles, (unsigned)((N + rpb - 1) / rpb));
    col_absmax_kernel<T, true><<<grid, 256, 0, st>>>(W, N, K, ld, rpb,
                                                     reinterpret_cast<unsigned int*>(colmax));
  } else {
    const int64_t col_tiles = (K + 255) / 256;
    const int rpb = pick_rows_per_block(N, col_tiles);
    dim3 grid((unsigned)col_tiles, (unsigned)((N + rpb - 1) / rpb));
    col_absmax_scalar_kernel<T><<<grid, 256, 0, st>>>(W, N, K, ld, rpb,
                                                      reinterpret_cast<unsigned int*>(colmax));
  }
  count_launch();
  return check_launch("col_absmax");
}

// =================================================================================================
// GPTQ column stage, reference-parity semantics (ref: gptq_quantizer.py:167-206)
// =================================================================================================
template <typename T>
__device__ __forceinline__ float clamp_min_1e5() {
  return ST<T>::rnd(1e-5f);  // torch casts the python scalar to the tensor's dtype
}

template <typename T, bool VECTOR>
__global__ void __launch_bounds__(256)
gptq_parity_kernel(const T* __restrict__ W, T* __restrict__ out, int8_t* __restrict__ codes,
                   const float* __restrict__ colmax, float* __restrict__ scales, int64_t N,
                   int64_t K, int64_t ld, float maxint, int rows_per_block) {
  constexpr int VEC = VECTOR ? ST<T>::VEC : 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col0 = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  if (col0 >= K) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(N, r0 + (int64_t)rows_per_block);
  float s[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    // scale = clamp(max_val / max_int, min=1e-5)      gptq_quantizer.py:182-184
    s[j] = fmaxf(ST<T>::rnd(__fdiv_rn(colmax[col0 + j], maxint)), clamp_min_1e5<T>());
  }
  if (scales != nullptr && blockIdx.y == 0 && ty == 0) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) scales[col0 + j] = s[j];
  }
  const float lo = -maxint - 1.f, hi = maxint;
  Divisor d[VEC];                       // the column scale divides every row: y = 1/s once
#pragma unroll
  for (int j = 0; j < VEC; ++j) d[j] = Divisor(s[j]);
  if constexpr (VECTOR) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float a[4][ST<T>::VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_vec<T>(W + (r + 8 * u) * ld + col0, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float q[ST<T>::VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          // clamp(round(w / scale), -max_int-1, max_int) * scale    gptq_quantizer.py:187-188
          q[j] = clampf(rint_then_clamped(ST<T>::rnd(d[j].div(a[u][j]))), lo, hi);
          a[u][j] = ST<T>::rnd(q[j] * s[j]);
        }
        store_vec<T>(out + (r + 8 * u) * K + col0, a[u]);
        if (codes != nullptr) {
          int8_t* c = codes + (r + 8 * u) * K + col0;
#pragma unroll
          for (int j = 0; j < VEC; ++j) c[j] = (int8_t)q[j];
        }
      }
    }
    for (; r < r1; r += 8) {
      float a[ST<T>::VEC], q[ST<T>::VEC];
      load_vec<T>(W + r * ld + col0, a);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        q[j] = clampf(rint_then_clamped(ST<T>::rnd(d[j].div(a[j]))), lo, hi);
        a[j] = ST<T>::rnd(q[j] * s[j]);
      }
      store_vec<T>(out + r * K + col0, a);
      if (codes != nullptr) {
        int8_t* c = codes + r * K + col0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) c[j] = (int8_t)q[j];
      }
    }
  } else {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float w = to_f(W[r * ld + col0]);
      const float q = clampf(rint_then_clamped(ST<T>::rnd(d[0].div(w))), lo, hi);
      out[r * K + col0] = from_f<T>(q * s[0]);
      if (codes != nullptr) codes[r * K + col0] = (int8_t)q;
    }
  }
}

// =================================================================================================
// uniform group fake-quant
//   asym: ref quantization_utils.py:390-405       sym: ref gptq_quantizer.py:94-100
// =================================================================================================
struct GroupQuantArgs {
  float maxint;
  int64_t n_groups;
  int64_t G;   // elements per group
  int64_t K;   // row length (for column ops)
  const float* colvec;
  const float* colrcp;   // RN(1 / colvec[k]) when the caller has it (else computed per element)
  int binary;            // colvec holds only {1, bin_factor}: reciprocal = select(1, bin_rcp)
  float bin_rcp;
  void* codes;
  float* scales;
  float* zeros;
};

template <typename T, int COLOP>
__device__ __forceinline__ float pre_op(float w, const Divisor& cv) {
  if constexpr (COLOP == B200Q_COLOP_MUL_DIV) return ST<T>::rnd(w * cv.b);  // awq_quantizer.py:70
  if constexpr (COLOP == B200Q_COLOP_DIV) return ST<T>::rnd(cv.div(w));     // smooth:170
  return w;
}
template <typename T, int COLOP>
__device__ __forceinline__ float post_op(float o, const Divisor& cv) {
  if constexpr (COLOP == B200Q_COLOP_MUL_DIV) return ST<T>::rnd(cv.div(o));  // awq:81
  return o;
}

// per-group parameters from the group's min/max (asym) or |max| (sym)
template <typename T, bool SYM>
__device__ __forceinline__ void group_params(float mx, float mn, float maxint, float& scale,
                                             float& zp) {
  if constexpr (SYM) {
    // scales = clamp(max_val / max_int, min=1e-5)                 gptq_quantizer.py:94-97
    scale = fmaxf(ST<T>::rnd(__fdiv_rn(mx, maxint)), clamp_min_1e5<T>());
    zp = 0.f;
  } else {
    // scales = (max - min).clamp(min=1e-5) / max_int               quantization_utils.py:395
    scale = ST<T>::rnd(__fdiv_rn(fmaxf(ST<T>::rnd(mx - mn), clamp_min_1e5<T>()), maxint));
    // zeros = (-round(min / scales)).clamp_(0, max_int)            quantization_utils.py:396
    zp = clampf(-rint_then_clamped(ST<T>::rnd(__fdiv_rn(mn, scale))), 0.f, maxint);
  }
}
template <typename T, bool SYM>
__device__ __forceinline__ float quant_one(float x, const Divisor& scale, float zp, float maxint,
                                           float& code) {
  if constexpr (SYM) {
    code = clampf(rint_then_clamped(ST<T>::rnd(scale.div(x))), -maxint - 1.f, maxint);
    return ST<T>::rnd(code * scale.b);
  } else {
    // w_q = clamp(round(w / scales) + zeros, 0, max_int); w = (w_q - zeros) * scales   :402-405
    code = clampf(ST<T>::rnd(rint_then_clamped(ST<T>::rnd(scale.div(x))) + zp), 0.f, maxint);
    return ST<T>::rnd(ST<T>::rnd(code - zp) * scale.b);
  }
}

// G == 128: EIGHT lanes own one group, 16 elements per lane as 4 (fp32) or 2 (16-bit) 128-bit
// accesses; a warp works on four consecutive groups.  Compared with one-warp-per-group this cuts
// the per-group scalar work (min/max shuffles, scale / zero-point divisions, reciprocal) that
// every lane repeats from 1/4 to 1/16 of an element's cost, and the per-element division is the
// 5-op reused-divisor form — together ~20 thread-instructions per element, below the ~44 the SMs
// can issue per element at HBM speed.  Groups whose values leave [1e-18, 1e18] (where the
// unguarded division core is not proven exact) are redone with plain IEEE divisions.
template <typename T, bool SYM, int COLOP, bool EXACT_SLOW>
__device__ __forceinline__ void quantize_group16(float (&x)[16], const float (&cv)[16],
                                                 const float (&cr)[16], float mx, float mn,
                                                 const GroupQuantArgs& a, float& scale, float& zp,
                                                 float (&code)[16]) {
  group_params<T, SYM>(mx, mn, a.maxint, scale, zp);
  const Divisor sd(scale);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float q;
    if constexpr (EXACT_SLOW) q = __fdiv_rn(x[e], scale);
    else q = sd.div_core(x[e]);
    float o;
    if constexpr (SYM) {
      code[e] = clampf(rint_then_clamped(ST<T>::rnd(q)), -a.maxint - 1.f, a.maxint);
      o = ST<T>::rnd(code[e] * scale);
    } else {
      code[e] = clampf(ST<T>::rnd(rint_then_clamped(ST<T>::rnd(q)) + zp), 0.f, a.maxint);
      o = ST<T>::rnd(ST<T>::rnd(code[e] - zp) * scale);
    }
    if constexpr (COLOP == B200Q_COLOP_MUL_DIV) {
      if constexpr (EXACT_SLOW) o = ST<T>::rnd(__fdiv_rn(o, cv[e]));
      else o = ST<T>::rnd(Divisor(cv[e], cr[e]).div_core(o));
    }
    x[e] = o;
  }
}

template <typename T, bool SYM, int COLOP>
__global__ void __launch_bounds__(256)
group128_kernel(const T* __restrict__ W, T* __restrict__ out, GroupQuantArgs a) {
  constexpr int VEC = ST<T>::VEC;   // elements per 128-bit access
  constexpr int NV = 16 / VEC;      // accesses per lane
  const int lane = threadIdx.x & 31;
  const int l8 = lane & 7, sub = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups_per_row = a.K / 128;
  for (int64_t base = warp * 4; base < a.n_groups; base += nwarps * 4) {
    const int64_t g = base + sub;
    const bool valid = g < a.n_groups;
    const int64_t gg = valid ? g : a.n_groups - 1;
    const T* wp = W + gg * 128 + l8 * VEC;
    float x[16], cv[16], cr[16];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t[VEC];
      load_vec<T>(wp + i * 8 * VEC, t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) x[i * VEC + j] = t[j];
    }
    if constexpr (COLOP != B200Q_COLOP_NONE) {
      const int64_t c0 = (gg % groups_per_row) * 128 + l8 * VEC;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(a.colvec + c0 + i * 8 * VEC + j));
          cv[i * VEC + j] = f.x; cv[i * VEC + j + 1] = f.y;
          cv[i * VEC + j + 2] = f.z; cv[i * VEC + j + 3] = f.w;
        }
      if (a.binary) {
        // colvec is {1, factor}: the reciprocal is a select, not a second vector
#pragma unroll
        for (int e = 0; e < 16; ++e) cr[e] = (cv[e] == 1.f) ? 1.f : a.bin_rcp;
      } else if (a.colrcp != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int j = 0; j < VEC; j += 4) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(a.colrcp + c0 + i * 8 * VEC + j));
            cr[i * VEC + j] = f.x; cr[i * VEC + j + 1] = f.y;
            cr[i * VEC + j + 2] = f.z; cr[i * VEC + j + 3] = f.w;
          }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) cr[e] = __frcp_rn(cv[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) { cv[e] = 1.f; cr[e] = 1.f; }
    }
    // column pre-op.  The unguarded division core is exact for operands in [1e-18, 1e18]; results
    // outside that (or NaN) send the whole group through the IEEE path below.
    bool cfast = true;
    if constexpr (COLOP == B200Q_COLOP_MUL_DIV) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        x[e] = ST<T>::rnd(x[e] * cv[e]);                                   // awq_quantizer.py:70
        cfast = cfast && (fabsf(cv[e]) > 1e-18f) && (fabsf(cv[e]) < 1e18f);
      }
    } else if constexpr (COLOP == B200Q_COLOP_DIV) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        cfast = cfast && (fabsf(cv[e]) > 1e-18f) && (fabsf(cv[e]) < 1e18f);
        x[e] = ST<T>::rnd(Divisor(cv[e], cr[e]).div_core(x[e]));          // smooth_quant:170
      }
    }
    float mx, mn;
    if constexpr (SYM) {
      mx = fabsf(x[0]);
#pragma unroll
      for (int e = 1; e < 16; ++e) mx = fmaxf(mx, fabsf(x[e]));
      mn = 0.f;
    } else {
      mx = x[0]; mn = x[0];
#pragma unroll
      for (int e = 1; e < 16; ++e) { mx = fmaxf(mx, x[e]); mn = fminf(mn, x[e]); }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if constexpr (!SYM) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const float gmax = fmaxf(fabsf(mx), fabsf(mn));
    // group-uniform decision (all 8 lanes see the same gmax; cfast is combined across them)
    unsigned ok = __ballot_sync(0xffffffffu, cfast);
    const bool fast = (((ok >> (sub * 8)) & 0xffu) == 0xffu) && (gmax < 1e18f);
    float scale, zp, code[16];
    if (fast) {
      quantize_group16<T, SYM, COLOP, false>(x, cv, cr, mx, mn, a, scale, zp, code);
    } else {
      // rare: reload and redo everything with IEEE divisions
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float t[VEC];
        load_vec<T>(wp + i * 8 * VEC, t);
#pragma unroll
        for (int j = 0; j < VEC; ++j) x[i * VEC + j] = t[j];
      }
      if constexpr (COLOP == B200Q_COLOP_MUL_DIV) {
#pragma unroll
        for (int e = 0; e < 16; ++e) x[e] = ST<T>::rnd(x[e] * cv[e]);
      } else if constexpr (COLOP == B200Q_COLOP_DIV) {
#pragma unroll
        for (int e = 0; e < 16; ++e) x[e] = ST<T>::rnd(__fdiv_rn(x[e], cv[e]));
      }
      if constexpr (SYM) {
        mx = fabsf(x[0]);
#pragma unroll
        for (int e = 1; e < 16; ++e) mx = fmaxf(mx, fabsf(x[e]));
      } else {
        mx = x[0]; mn = x[0];
#pragma unroll
        for (int e = 1; e < 16; ++e) { mx = fmaxf(mx, x[e]); mn = fminf(mn, x[e]); }
      }
      // this branch is taken by whole 8-lane teams: name the team explicitly (under independent
      // thread scheduling __activemask() need not contain all eight lanes at this point)
      const unsigned team = 0xffu << (sub * 8);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(team, mx, o, 8));
        if constexpr (!SYM) mn = fminf(mn, __shfl_xor_sync(team, mn, o, 8));
      }
      quantize_group16<T, SYM, COLOP, true>(x, cv, cr, mx, mn, a, scale, zp, code);
    }
    if (valid) {
      T* op = out + g * 128 + l8 * VEC;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float t[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) t[j] = x[i * VEC + j];
        store_vec<T>(op + i * 8 * VEC, t);
      }
      if (a.codes != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const int64_t at = g * 128 + i * 8 * VEC + l8 * VEC + j;
            if constexpr (SYM) static_cast<int8_t*>(a.codes)[at] = (int8_t)code[i * VEC + j];
            else static_cast<uint8_t*>(a.codes)[at] = (uint8_t)code[i * VEC + j];
          }
      }
      if (l8 == 0) {
        if (a.scales != nullptr) a.scales[g] = scale;
        if (a.zeros != nullptr) a.zeros[g] = zp;
      }
    }
  }
}

// =================================================================================================
// AWQ scale search, stage 1: dW_c = Q_c(W) - W for every candidate scale factor, as bf16
// (ref: awq_quantizer.py:116-119 "for each scale factor, quantize and measure reconstruction error";
//  Q_c is exactly awq_quantize_model_weight's arithmetic with scale_factor = sf_c, :70-81).
// One read of W serves all candidates: a group stays in registers while the candidates are
// quantised one after the other; candidate c's deltas go to D + c * rows_pad * K.
// =================================================================================================
struct CandParam {
  float sf[32];
  float sf_rcp[32];
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256)
awq_delta128_kernel(const T* __restrict__ W, __nv_bfloat16* __restrict__ D,
                    const uint8_t* __restrict__ salient, int64_t n_groups, int64_t K,
                    int64_t cand_stride, float maxint, CandParam cp) {
  constexpr int VEC = ST<T>::VEC;
  constexpr int NV = 16 / VEC;
  const int lane = threadIdx.x & 31;
  const int l8 = lane & 7, sub = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups_per_row = K / 128;
  GroupQuantArgs a;
  a.maxint = maxint;
  for (int64_t base = warp * 4; base < n_groups; base += nwarps * 4) {
    const int64_t g = base + sub;
    const bool valid = g < n_groups;
    const int64_t gg = valid ? g : n_groups - 1;
    const T* wp = W + gg * 128 + l8 * VEC;
    float w[16];
    uint32_t smask = 0;     // bit e: element e of this lane sits on a salient column
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t[VEC];
      load_vec<T>(wp + i * 8 * VEC, t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) w[i * VEC + j] = t[j];
    }
    const int64_t c0 = (gg % groups_per_row) * 128 + l8 * VEC;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        smask |= (salient[c0 + i * 8 * VEC + j] ? 1u : 0u) << (i * VEC + j);
    // What changes from one candidate to the next is only the value of the SALIENT elements (1 % of
    // the columns): x = w * sf_c there, x = w everywhere else.  The group's scale and zero point
    // follow from (max, min) over both kinds, and they often do not move at all between candidates
    // (the scaled salient value is rarely the group's extreme).  So: (max, min) of the non-salient
    // elements once per group; per candidate only the salient elements are re-scaled, and when the
    // resulting (scale, zero point) are bit-identical to the previous candidate's, the deltas of
    // the non-salient elements are REUSED -- same arithmetic, same bits, ~1/10 of the instructions.
    float mx0 = -INFINITY, mn0 = INFINITY;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      if (!((smask >> e) & 1u)) { mx0 = fmaxf(mx0, w[e]); mn0 = fminf(mn0, w[e]); }
    float d[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) d[e] = 0.f;
    float prev_scale = __int_as_float(0x7fc00000), prev_zp = 0.f;      // NaN: never equal
    for (int c = 0; c < cp.n; ++c) {
      const float sf = cp.sf[c];
      float mx = mx0, mn = mn0;
      if (smask != 0u) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if ((smask >> e) & 1u) {
            const float xs = ST<T>::rnd(w[e] * sf);                        // awq_quantizer.py:70
            mx = fmaxf(mx, xs); mn = fminf(mn, xs);
          }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float scale, zp;
      group_params<T, false>(mx, mn, maxint, scale, zp);
      const Divisor sd(scale);
      // the search only ranks candidates: the (exact for |w| < 1e18) fast division is enough here
      auto quant = [&](float x) {
        const float q = sd.div_core(x);
        const float code = clampf(ST<T>::rnd(rint_then_clamped(ST<T>::rnd(q)) + zp), 0.f, maxint);
        return ST<T>::rnd(ST<T>::rnd(code - zp) * scale);
      };
      const bool same = (scale == prev_scale) && (zp == prev_zp);          // uniform per 8-lane team
      if (!same) {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (!((smask >> e) & 1u)) d[e] = quant(w[e]) - w[e];
        prev_scale = scale; prev_zp = zp;
      }
      if (smask != 0u) {
        const Divisor sfd(sf, cp.sf_rcp[c]);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if ((smask >> e) & 1u) {
            const float o = quant(ST<T>::rnd(w[e] * sf));
            d[e] = ST<T>::rnd(sfd.div_core(o)) - w[e];                       // awq_quantizer.py:81
          }
      }
      if (valid) {
        __nv_bfloat16* dp = D + c * cand_stride + g * 128 + l8 * VEC;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float t[VEC];
#pragma unroll
          for (int j = 0; j < VEC; ++j) t[j] = d[i * VEC + j];
          if constexpr (VEC == 8) {
            store_vec<__nv_bfloat16>(dp + i * 64, t);
          } else {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(t[0], t[1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(t[2], t[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0);
            pk.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(dp + i * 32) = pk;
          }
        }
      }
    }
  }
}

// Any group length (32, 64, 256, per-row ...): one warp per group, and per candidate a min/max pass
// and a quantize pass over the group (re-read through L1/L2; W itself comes from HBM once).  IEEE
// divisions throughout -- this is the boundary-completeness path, not the measured one.
template <typename T>
__global__ void __launch_bounds__(256)
awq_delta_generic_kernel(const T* __restrict__ W, __nv_bfloat16* __restrict__ D,
                         const uint8_t* __restrict__ salient, int64_t n_groups, int64_t G, int64_t K,
                         int64_t cand_stride, float maxint, CandParam cp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const int64_t off = g * G;
    for (int c = 0; c < cp.n; ++c) {
      const float sf = cp.sf[c];
      float mx = -INFINITY, mn = INFINITY;
      for (int64_t i = lane; i < G; i += 32) {
        const float w = to_f(W[off + i]);
        const float x = salient[(off + i) % K] ? ST<T>::rnd(w * sf) : w;     // awq_quantizer.py:70
        mx = fmaxf(mx, x); mn = fminf(mn, x);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float scale, zp;
      group_params<T, false>(mx, mn, maxint, scale, zp);
      for (int64_t i = lane; i < G; i += 32) {
        const float w = to_f(W[off + i]);
        const bool s = salient[(off + i) % K] != 0;
        const float x = s ? ST<T>::rnd(w * sf) : w;
        const float code = clampf(ST<T>::rnd(rintf(ST<T>::rnd(__fdiv_rn(x, scale))) + zp), 0.f, maxint);
        float o = ST<T>::rnd(ST<T>::rnd(code - zp) * scale);
        if (s) o = ST<T>::rnd(__fdiv_rn(o, sf));                               // awq_quantizer.py:81
        D[c * cand_stride + off + i] = __float2bfloat16_rn(o - w);
      }
    }
  }
}

int launch_awq_delta(const void* W, void* D, const uint8_t* salient, int64_t N, int64_t K,
                     int64_t group, int64_t cand_stride, int n_bit, const float* sf_host, int n_cand,
                     int dtype, cudaStream_t st) {
  const int64_t G = (group > 0 && group < K) ? group : K;
  B200Q_REQUIRE(K % G == 0, "awq_search: in_features not divisible by group size");
  B200Q_REQUIRE(n_cand >= 1 && n_cand <= 32, "awq_search: 1..32 candidates per call");
  B200Q_REQUIRE(aligned16(W) && aligned16(D), "awq_search: unaligned pointer");
  CandParam cp;
  cp.n = n_cand;
  for (int i = 0; i < 32; ++i) {
    cp.sf[i] = i < n_cand ? sf_host[i] : 1.f;
    cp.sf_rcp[i] = 1.0f / cp.sf[i];
  }
  const float maxint = (float)((1 << n_bit) - 1);
  const int64_t n_groups = N * (K / G);
  if (G == 128) {
    const int64_t warps_needed = (n_groups + 3) / 4;
    int64_t blocks = std::min<int64_t>((warps_needed + 7) / 8, (int64_t)kNumSMs * 8);
    B200Q_DISPATCH_DTYPE(dtype, T,
                         (awq_delta128_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
                             static_cast<const T*>(W), static_cast<__nv_bfloat16*>(D), salient,
                             n_groups, K, cand_stride, maxint, cp)));
  } else {
    int64_t blocks = std::min<int64_t>((n_groups + 7) / 8, (int64_t)kNumSMs * 16);
    B200Q_DISPATCH_DTYPE(dtype, T,
                         (awq_delta_generic_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
                             static_cast<const T*>(W), static_cast<__nv_bfloat16*>(D), salient,
                             n_groups, G, K, cand_stride, maxint, cp)));
  }
  count_launch();
  return check_launch("awq_delta");
}

// =================================================================================================
// SmoothQuant alpha sweep (ref: smooth_quant_quantizer.py:327-371, a stub there): for every candidate
// alpha a, with s_a the smoothing scale of that alpha,
//     err[a] (+)= sum_{i,k} ( (Q(W[i,k] / s_a[k]) * s_a[k] - W[i,k]) * act[k] )^2
// i.e. the error of the smoothed-then-quantized weight mapped back to the original basis and
// weighted by the calibration activation magnitude.  One warp owns one quantization group and
// sweeps ALL alphas over it (min/max pass + quantize pass per alpha), so W comes from HBM once and
// from L1 afterwards; nothing of size N x K is written.  Per-warp partial sums, reduced in a fixed
// order by smooth_alpha_reduce_kernel (deterministic).
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
smooth_alpha_err_kernel(const T* __restrict__ W, const float* __restrict__ S,
                        const float* __restrict__ act, int64_t n_groups, int64_t G, int64_t K,
                        float maxint, int n_alpha, double* __restrict__ partial) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  __shared__ double accs[8][32];                       // per warp, per alpha (n_alpha <= 32)
  if (lane < n_alpha) accs[wib][lane] = 0.0;
  __syncwarp();
  for (int64_t g = warp; g < n_groups; g += nwarps) {
    const int64_t off = g * G;
    const int64_t k0 = off % K;                        // a group never crosses a row (G divides K)
    for (int a = 0; a < n_alpha; ++a) {                // the group stays in L1 across the alphas
      const float* sa = S + (int64_t)a * K + k0;
      float mx = -INFINITY, mn = INFINITY;
      for (int64_t i = lane; i < G; i += 32) {
        const float x = ST<T>::rnd(__fdiv_rn(to_f(W[off + i]), sa[i]));
        mx = fmaxf(mx, x); mn = fminf(mn, x);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float scale, zp;
      group_params<T, false>(mx, mn, maxint, scale, zp);
      float e2 = 0.f;
      for (int64_t i = lane; i < G; i += 32) {
        const float w = to_f(W[off + i]);
        const float sv = sa[i];
        const float x = ST<T>::rnd(__fdiv_rn(w, sv));
        const float code = clampf(ST<T>::rnd(rint_then_clamped(ST<T>::rnd(__fdiv_rn(x, scale))) + zp),
                                  0.f, maxint);
        const float deq = ST<T>::rnd(ST<T>::rnd(code - zp) * scale);
        const float e = (ST<T>::rnd(deq * sv) - w) * act[k0 + i];
        e2 = fmaf(e, e, e2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
      if (lane == 0) accs[wib][a] += (double)e2;
    }
  }
  __syncwarp();
  if (lane < n_alpha) partial[(int64_t)lane * nwarps + warp] = accs[wib][lane];
}

// G == 128: the register form of group128_kernel -- eight lanes own a group (16 elements per lane,
// loaded once with 128-bit accesses and kept in registers across all alphas), a warp works on four
// consecutive groups.  Same arithmetic per element as the generic kernel above.
template <typename T>
__global__ void __launch_bounds__(256)
smooth_alpha_err128_kernel(const T* __restrict__ W, const float* __restrict__ S,
                           const float* __restrict__ act, int64_t n_groups, int64_t K, float maxint,
                           int n_alpha, double* __restrict__ partial) {
  constexpr int VEC = ST<T>::VEC;
  constexpr int NV = 16 / VEC;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int l8 = lane & 7, sub = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups_per_row = K / 128;
  __shared__ double accs[8][32];
  if (lane < n_alpha) accs[wib][lane] = 0.0;
  __syncwarp();
  for (int64_t base = warp * 4; base < n_groups; base += nwarps * 4) {
    const int64_t g = base + sub;
    const bool valid = g < n_groups;
    const int64_t gg = valid ? g : n_groups - 1;
    const T* wp = W + gg * 128 + l8 * VEC;
    const int64_t c0 = (gg % groups_per_row) * 128 + l8 * VEC;
    float w[16], aw[16];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t[VEC];
      load_vec<T>(wp + i * 8 * VEC, t);
#pragma unroll
      for (int j = 0; j < VEC; ++j) w[i * VEC + j] = t[j];
#pragma unroll
      for (int j = 0; j < VEC; j += 4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(act + c0 + i * 8 * VEC + j));
        aw[i * VEC + j] = f.x; aw[i * VEC + j + 1] = f.y;
        aw[i * VEC + j + 2] = f.z; aw[i * VEC + j + 3] = f.w;
      }
    }
    for (int a = 0; a < n_alpha; ++a) {
      const float* sa = S + (int64_t)a * K + c0;
      float sv[16], x[16];
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(sa + i * 8 * VEC + j));
          sv[i * VEC + j] = f.x; sv[i * VEC + j + 1] = f.y;
          sv[i * VEC + j + 2] = f.z; sv[i * VEC + j + 3] = f.w;
        }
      float mx = -INFINITY, mn = INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        x[e] = ST<T>::rnd(__fdiv_rn(w[e], sv[e]));
        mx = fmaxf(mx, x[e]); mn = fminf(mn, x[e]);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      float scale, zp;
      group_params<T, false>(mx, mn, maxint, scale, zp);
      const Divisor sd(scale);
      float e2 = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float code = clampf(ST<T>::rnd(rint_then_clamped(ST<T>::rnd(sd.div(x[e]))) + zp), 0.f, maxint);
        const float deq = ST<T>::rnd(ST<T>::rnd(code - zp) * scale);
        const float err = (ST<T>::rnd(deq * sv[e]) - w[e]) * aw[e];
        e2 = fmaf(err, err, e2);
      }
      if (!valid) e2 = 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
      if (lane == 0) accs[wib][a] += (double)e2;
    }
  }
  __syncwarp();
  if (lane < n_alpha) partial[(int64_t)lane * nwarps + warp] = accs[wib][lane];
}

__global__ void smooth_alpha_reduce_kernel(const double* __restrict__ partial, int64_t nwarps,
                                           int n_alpha, double* __restrict__ err, int accumulate) {
  const int a = blockIdx.x;
  if (a >= n_alpha) return;
  __shared__ double sm[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < nwarps; i += blockDim.x) s += partial[(int64_t)a * nwarps + i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) err[a] = (accumulate ? err[a] : 0.0) + sm[0];
}

// any group length: one warp per group, two passes (the second re-reads through L2).
template <typename T, bool SYM, int COLOP>
__global__ void __launch_bounds__(256)
group_generic_kernel(const T* __restrict__ W, T* __restrict__ out, GroupQuantArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = warp; g < a.n_groups; g += nwarps) {
    const int64_t off = g * a.G;
    float mx = SYM ? 0.f : -INFINITY, mn = INFINITY;
    for (int64_t i = lane; i < a.G; i += 32) {
      Divisor cv;
      if constexpr (COLOP != B200Q_COLOP_NONE) cv = Divisor(a.colvec[(off + i) % a.K]);
      const float x = pre_op<T, COLOP>(to_f(W[off + i]), cv);
      if constexpr (SYM) mx = fmaxf(mx, fabsf(x));
      else { mx = fmaxf(mx, x); mn = fminf(mn, x); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    float scale, zp;
    group_params<T, SYM>(mx, mn, a.maxint, scale, zp);
    const Divisor sd(scale);
    for (int64_t i = lane; i < a.G; i += 32) {
      Divisor cv;
      if constexpr (COLOP != B200Q_COLOP_NONE) cv = Divisor(a.colvec[(off + i) % a.K]);
      const float x = pre_op<T, COLOP>(to_f(W[off + i]), cv);
      float code;
      const float o = post_op<T, COLOP>(quant_one<T, SYM>(x, sd, zp, a.maxint, code), cv);
      out[off + i] = from_f<T>(o);
      if (a.codes != nullptr) {
        if constexpr (SYM) static_cast<int8_t*>(a.codes)[off + i] = (int8_t)code;
        else static_cast<uint8_t*>(a.codes)[off + i] = (uint8_t)code;
      }
    }
    if (lane == 0) {
      if (a.scales != nullptr) a.scales[g] = scale;
      if (a.zeros != nullptr) a.zeros[g] = zp;
    }
  }
}

template <typename T, bool SYM, int COLOP>
static int launch_group_quant(const void* W, void* out, const GroupQuantArgs& a, bool fast128,
                              cudaStream_t st) {
  if (a.n_groups == 0) return B200Q_OK;
  if (fast128) {
    const int64_t warps_needed = (a.n_groups + 3) / 4;
    int64_t blocks = (warps_needed + 7) / 8;
    blocks = std::min<int64_t>(blocks, (int64_t)kNumSMs * 8 * 4);  // grid-stride beyond 4 waves
    if (blocks > kNumSMs) blocks = blocks / kNumSMs * kNumSMs;      // whole waves
    group128_kernel<T, SYM, COLOP><<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const T*>(W), static_cast<T*>(out), a);
  } else {
    int64_t blocks = (a.n_groups + 7) / 8;
    blocks = std::min<int64_t>(blocks, (int64_t)kNumSMs * 8 * 4);
    group_generic_kernel<T, SYM, COLOP><<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const T*>(W), static_cast<T*>(out), a);
  }
  count_launch();
  return check_launch("group_fakequant");
}

// =================================================================================================
// SmoothQuant
// =================================================================================================
// torch.pow(tensor, python_float) has exact special cases (sqrt, x*x, ...): replicate them so the
// default alpha = 0.5 is bit-exact; a general exponent goes through powf (<= 2 ulp from SLEEF's).
__device__ __forceinline__ float torch_pow_scalar(float x, float e) {
  if (e == 0.f) return 1.f;
  if (e == 1.f) return x;
  if (e == 0.5f) return __fsqrt_rn(x);
  if (e == 2.f) return x * x;
  if (e == 3.f) return x * x * x;
  if (e == -0.5f) return __fdiv_rn(1.f, __fsqrt_rn(x));
  if (e == -1.f) return __fdiv_rn(1.f, x);
  if (e == -2.f) return __fdiv_rn(1.f, x * x);
  return powf(x, e);
}

__global__ void smooth_scale_kernel(const float* __restrict__ act, const float* __restrict__ wmax,
                                    float* __restrict__ s, float* __restrict__ s_rcp, int64_t K,
                                    float alpha, float one_m_alpha, int act_dtype, int w_dtype,
                                    int res_dtype) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  // act_scale = clamp(act_scale, 1e-5); weight_scale = clamp(weight_scale, 1e-5)   smooth:159-160
  const float a = fmaxf(act[k], rnd_rt(1e-5f, act_dtype));
  const float w = fmaxf(wmax[k], rnd_rt(1e-5f, w_dtype));
  // s = pow(a, alpha) / pow(w, 1 - alpha); s = clamp(s, 1e-5)                       smooth:165-166
  const float pa = rnd_rt(torch_pow_scalar(a, alpha), act_dtype);
  const float pw = rnd_rt(torch_pow_scalar(w, one_m_alpha), w_dtype);
  const float sk = fmaxf(rnd_rt(__fdiv_rn(pa, pw), res_dtype), rnd_rt(1e-5f, res_dtype));
  s[k] = sk;
  if (s_rcp != nullptr) s_rcp[k] = __frcp_rn(sk);
}

template <typename T, bool MUL, bool VECTOR>
__global__ void __launch_bounds__(256)
col_scale_kernel(const T* __restrict__ W, T* __restrict__ out, const float* __restrict__ s,
                 int64_t N, int64_t K, int rows_per_block) {
  constexpr int VEC = VECTOR ? ST<T>::VEC : 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col0 = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  if (col0 >= K) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(N, r0 + (int64_t)rows_per_block);
  float sv[VEC];
  Divisor sd[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { sv[j] = s[col0 + j]; sd[j] = Divisor(sv[j]); }
  if constexpr (VECTOR) {
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float a[4][ST<T>::VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_vec<T>(W + (r + 8 * u) * K + col0, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          a[u][j] = ST<T>::rnd(MUL ? a[u][j] * sv[j] : sd[j].div(a[u][j]));
        store_vec<T>(out + (r + 8 * u) * K + col0, a[u]);
      }
    }
    for (; r < r1; r += 8) {
      float a[ST<T>::VEC];
      load_vec<T>(W + r * K + col0, a);
#pragma unroll
      for (int j = 0; j < VEC; ++j) a[j] = ST<T>::rnd(MUL ? a[j] * sv[j] : sd[j].div(a[j]));
      store_vec<T>(out + r * K + col0, a);
    }
  } else {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float w = to_f(W[r * K + col0]);
      out[r * K + col0] = from_f<T>(MUL ? w * sv[0] : sd[0].div(w));
    }
  }
}

// =================================================================================================
// activation statistics (ref: quantization_utils.py:231, smooth_quant_quantizer.py:68)
// =================================================================================================
// mean|x| over tokens: stage 1 writes one fp32 partial per (row chunk, column) in a fixed order,
// stage 2 adds the chunks in order and divides by T -> deterministic, no float atomics.
template <typename T>
__global__ void __launch_bounds__(256)
act_abssum_partial_kernel(const T* __restrict__ X, int64_t Trows, int64_t K, int rows_per_block,
                          float* __restrict__ partial) {
  constexpr int VEC = ST<T>::VEC;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col0 = ((int64_t)blockIdx.x * 32 + tx) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(Trows, r0 + (int64_t)rows_per_block);
  // blockIdx.z = sample of a batch of equal-length samples stacked along the rows
  X += (int64_t)blockIdx.z * Trows * K;
  partial += (int64_t)blockIdx.z * gridDim.y * K;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (col0 < K) {
    const T* p = X + col0;
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      float a[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u) load_vec<T>(p + (r + 8 * u) * K, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += fabsf(a[u][j]);
    }
    for (; r < r1; r += 8) {
      float a[VEC];
      load_vec<T>(p + r * K, a);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += fabsf(a[j]);
    }
  }
  __shared__ float sm[8][32 * VEC + 1];
#pragma unroll
  for (int j = 0; j < VEC; ++j) sm[ty][tx * VEC + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int64_t col = (int64_t)blockIdx.x * 32 * VEC + c;
    if (col < K) {
      float v = sm[0][c];
#pragma unroll
      for (int y = 1; y < 8; ++y) v += sm[y][c];
      partial[(int64_t)blockIdx.y * K + col] = v;
    }
  }
}

template <typename T>
__global__ void act_mean_finish_kernel(const float* __restrict__ partial, int chunks, int64_t K,
                                       float count, float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  partial += (int64_t)blockIdx.y * chunks * K;     // blockIdx.y = sample
  float v = 0.f;
  for (int c = 0; c < chunks; ++c) v += partial[(int64_t)c * K + k];
  // torch: sum (rounded to the tensor dtype) then div_ by the row count
  out[(int64_t)blockIdx.y * K + k] = ST<T>::rnd(__fdiv_rn(ST<T>::rnd(v), count));
}

template <typename T>
__global__ void seq_sum_rows_kernel(const T* __restrict__ V, int64_t n, int64_t K,
                                    float* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float v = 0.f;  // Python's sum() starts from int 0; every partial sum is a tensor of V's dtype
  for (int64_t i = 0; i < n; ++i) v = ST<T>::rnd(v + to_f(V[i * K + k]));
  out[k] = v;
}

// =================================================================================================
// top-k channel mask (ref: awq_quantizer.py:60-61  torch.topk(importance, n_protect))
// =================================================================================================
// One CTA: 4-pass MSB radix select over the order-preserving uint image of the floats finds the
// k-th largest value; everything above it is salient, ties at the threshold are taken in index
// order.  Writes colmul[i] = (salient ? factor : 1) and, optionally, a uint8 mask.
__device__ __forceinline__ uint32_t ordered_bits(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024)
topk_colmul_kernel(const float* __restrict__ v, int K, int k, float factor, float factor_rcp,
                   float* __restrict__ colmul, float* __restrict__ colrcp,
                   uint8_t* __restrict__ mask) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need, s_eq_base[1024];
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_need = (unsigned)k; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < K; i += blockDim.x) {
      const uint32_t u = ordered_bits(v[i]);
      if ((u & pmask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned need = s_need;
      int b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= need) break;
        need -= hist[b];
      }
      s_need = need;                       // how many still to take inside bin b
      s_prefix = prefix | ((uint32_t)b << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;           // exact bits of the k-th largest value
  const unsigned need_eq = s_need;         // number of elements == thr to accept (index order)
  // contiguous chunk per thread so that "index order" is a prefix sum over threads
  const int per = (K + blockDim.x - 1) / blockDim.x;
  const int i0 = tid * per, i1 = min(K, i0 + per);
  unsigned eq = 0;
  for (int i = i0; i < i1; ++i) eq += (ordered_bits(v[i]) == thr);
  s_eq_base[tid] = eq;
  __syncthreads();
  if (tid == 0) {
    unsigned run = 0;
    for (int t = 0; t < blockDim.x; ++t) { const unsigned c = s_eq_base[t]; s_eq_base[t] = run; run += c; }
  }
  __syncthreads();
  unsigned seen = s_eq_base[tid];
  for (int i = i0; i < i1; ++i) {
    const uint32_t u = ordered_bits(v[i]);
    bool take = u > thr;
    if (u == thr) { take = seen < need_eq; ++seen; }
    colmul[i] = take ? factor : 1.f;
    if (colrcp != nullptr) colrcp[i] = take ? factor_rcp : 1.f;
    if (mask != nullptr) mask[i] = take ? 1 : 0;
  }
}

}  // namespace b200q

// =================================================================================================
// C ABI
// =================================================================================================
using namespace b200q;

// The AWQ search runs these kernels on a side stream WHILE a tcgen05 GEMM (197 KB of dynamic
// shared memory per CTA, i.e. the maximum shared-memory carveout) occupies every SM.  Blocks of two
// kernels only share an SM if both run under the same L1 / shared-memory split, so the kernels
// that are meant to fill the GEMM's idle issue slots ask for the same (maximum-shared) carveout.
template <typename K>
static void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
}
static void side_stream_kernels_share_the_gemm_carveout() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64 || done[dev]) return;
  done[dev] = true;
  const char* e = std::getenv("B200Q_SIDE_CARVEOUT");
  if (e == nullptr || e[0] != '1') return;
  prefer_max_shared(act_abssum_partial_kernel<float>);
  prefer_max_shared(act_abssum_partial_kernel<__half>);
  prefer_max_shared(act_abssum_partial_kernel<__nv_bfloat16>);
  prefer_max_shared(act_mean_finish_kernel<float>);
  prefer_max_shared(act_mean_finish_kernel<__half>);
  prefer_max_shared(act_mean_finish_kernel<__nv_bfloat16>);
  prefer_max_shared(seq_sum_rows_kernel<float>);
  prefer_max_shared(seq_sum_rows_kernel<__half>);
  prefer_max_shared(seq_sum_rows_kernel<__nv_bfloat16>);
  prefer_max_shared(topk_colmul_kernel);
  prefer_max_shared(awq_delta128_kernel<float>);
  prefer_max_shared(awq_delta128_kernel<__half>);
  prefer_max_shared(awq_delta128_kernel<__nv_bfloat16>);
  cudaGetLastError();
  done[dev] = true;
}

extern "C" {

int b200q_col_absmax(const void* W, int64_t N, int64_t K, int64_t ld, int dtype, float* colmax,
                     int accumulate, void* stream) {
  B200Q_REQUIRE(W != nullptr && colmax != nullptr, "col_absmax: null pointer");
  B200Q_REQUIRE(N >= 0 && K > 0 && ld >= K, "col_absmax: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N == 0) {
    if (!accumulate) cudaMemsetAsync(colmax, 0, sizeof(float) * K, st);
    return B200Q_OK;
  }
  KernelScope scope("col_absmax", (double)N * K * elem_size(dtype), 0, st);
  B200Q_DISPATCH_DTYPE(dtype, T, return launch_col_absmax<T>(W, N, K, ld, colmax, accumulate, st));
  return B200Q_OK;
}

int b200q_gptq_parity_quant(const void* W, void* out, int8_t* codes, const float* colmax,
                            float* scales, int64_t N, int64_t K, int64_t ld, int n_bit, int dtype,
                            void* stream) {
  B200Q_REQUIRE(W && out && colmax, "gptq_parity_quant: null pointer");
  B200Q_REQUIRE(N >= 0 && K > 0 && ld >= K, "gptq_parity_quant: bad shape");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 16, "gptq_parity_quant: n_bit must be in [1,16]");
  B200Q_REQUIRE(codes == nullptr || n_bit <= 7, "gptq_parity_quant: int8 codes need n_bit <= 7");
  if (N == 0) return B200Q_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float maxint = (float)((1 << n_bit) - 1);
  KernelScope scope("gptq_parity_quant", 2.0 * N * K * elem_size(dtype), 0, st);
  B200Q_DISPATCH_DTYPE(dtype, T, {
    constexpr int VEC = ST<T>::VEC;
    const bool vec_ok = aligned16(W) && aligned16(out) && (K % VEC == 0) && (ld % VEC == 0);
    const int64_t cols_per_block = vec_ok ? 32 * VEC : 32;
    const int64_t col_tiles = (K + cols_per_block - 1) / cols_per_block;
    const int rpb = pick_rows_per_block(N, col_tiles);
    dim3 grid((unsigned)col_tiles, (unsigned)((N + rpb - 1) / rpb));
    if (vec_ok)
      gptq_parity_kernel<T, true><<<grid, 256, 0, st>>>(static_cast<const T*>(W),
                                                        static_cast<T*>(out), codes, colmax, scales,
                                                        N, K, ld, maxint, rpb);
    else
      gptq_parity_kernel<T, false><<<grid, 256, 0, st>>>(static_cast<const T*>(W),
                                                         static_cast<T*>(out), codes, colmax,
                                                         scales, N, K, ld, maxint, rpb);
  });
  count_launch();
  return check_launch("gptq_parity_quant");
}

static int group_fakequant_impl(const void* W, void* out, void* codes, float* scales, float* zeros,
                               int64_t N, int64_t K, int64_t group, int n_bit, int symmetric,
                               int colop, const float* colvec, const float* colrcp,
                               float binary_rcp, int dtype, void* stream) {
  B200Q_REQUIRE(W && out, "group_fakequant: null pointer");
  B200Q_REQUIRE(N >= 0 && K > 0, "group_fakequant: bad shape");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 16, "group_fakequant: n_bit must be in [1,16]");
  const int64_t G = group > 0 ? group : K;
  // ref: assert org_w_shape[-1] % q_group_size == 0   (quantization_utils.py:384)
  B200Q_REQUIRE(K % G == 0, "group_fakequant: in_features not divisible by group size");
  B200Q_REQUIRE(colop == B200Q_COLOP_NONE || colvec != nullptr, "group_fakequant: colvec missing");
  B200Q_REQUIRE(colop >= 0 && colop <= 2, "group_fakequant: bad colop");
  B200Q_REQUIRE(codes == nullptr || n_bit <= (symmetric ? 7 : 8),
                "group_fakequant: codes do not fit 8 bits");
  if (N == 0) return B200Q_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GroupQuantArgs a;
  a.maxint = (float)((1 << n_bit) - 1);
  a.n_groups = N * (K / G);
  a.G = G;
  a.K = K;
  a.colvec = colvec;
  a.colrcp = colrcp;
  a.binary = binary_rcp != 0.f;      // colvec is {1, factor}; binary_rcp = RN(1 / factor)
  a.bin_rcp = binary_rcp;
  a.codes = codes;
  a.scales = scales;
  a.zeros = zeros;
  const bool fast = (G == 128) && aligned16(W) && aligned16(out) &&
                    (colop == B200Q_COLOP_NONE ||
                     (aligned16(colvec) && (colrcp == nullptr || aligned16(colrcp))));
  KernelScope scope("group_fakequant", 2.0 * N * K * elem_size(dtype), 0, st);
#define B200Q_GQ(SYM, OP) return launch_group_quant<T, SYM, OP>(W, out, a, fast, st)
  B200Q_DISPATCH_DTYPE(dtype, T, {
    if (symmetric) {
      if (colop == B200Q_COLOP_NONE) B200Q_GQ(true, B200Q_COLOP_NONE);
      if (colop == B200Q_COLOP_MUL_DIV) B200Q_GQ(true, B200Q_COLOP_MUL_DIV);
      B200Q_GQ(true, B200Q_COLOP_DIV);
    } else {
      if (colop == B200Q_COLOP_NONE) B200Q_GQ(false, B200Q_COLOP_NONE);
      if (colop == B200Q_COLOP_MUL_DIV) B200Q_GQ(false, B200Q_COLOP_MUL_DIV);
      B200Q_GQ(false, B200Q_COLOP_DIV);
    }
  });
#undef B200Q_GQ
  return B200Q_OK;
}

int b200q_group_fakequant(const void* W, void* out, void* codes, float* scales, float* zeros,
                          int64_t N, int64_t K, int64_t group, int n_bit, int symmetric, int colop,
                          const float* colvec, int dtype, void* stream) {
  return group_fakequant_impl(W, out, codes, scales, zeros, N, K, group, n_bit, symmetric, colop,
                              colvec, nullptr, 0.f, dtype, stream);
}

static int promote_dtype(int a, int b) {
  if (a == b) return a;
  return B200Q_F32;
}

int b200q_smooth_scale(const float* act_scale, const float* wmax, float* s, int64_t K, float alpha,
                       int act_dtype, int w_dtype, void* stream) {
  B200Q_REQUIRE(act_scale && wmax && s && K > 0, "smooth_scale: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the reference evaluates 1.0 - alpha in Python doubles before torch narrows it
  const float one_m_alpha = (float)(1.0 - (double)alpha);
  smooth_scale_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(
      act_scale, wmax, s, nullptr, K, alpha, one_m_alpha, act_dtype, w_dtype,
      promote_dtype(act_dtype, w_dtype));
  count_launch();
  return check_launch("smooth_scale");
}

int b200q_col_scale(const void* W, void* out, const float* s, int64_t N, int64_t K, int mul,
                    int dtype, void* stream) {
  B200Q_REQUIRE(W && out && s && N >= 0 && K > 0, "col_scale: bad argument");
  if (N == 0) return B200Q_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("col_scale", 2.0 * N * K * elem_size(dtype), 0, st);
  B200Q_DISPATCH_DTYPE(dtype, T, {
    constexpr int VEC = ST<T>::VEC;
    const bool vec_ok = aligned16(W) && aligned16(out) && (K % VEC == 0);
    const int64_t cols_per_block = vec_ok ? 32 * VEC : 32;
    const int64_t col_tiles = (K + cols_per_block - 1) / cols_per_block;
    const int rpb = pick_rows_per_block(N, col_tiles);
    dim3 grid((unsigned)col_tiles, (unsigned)((N + rpb - 1) / rpb));
    const T* Wt = static_cast<const T*>(W);
    T* Ot = static_cast<T*>(out);
    if (vec_ok) {
      if (mul) col_scale_kernel<T, true, true><<<grid, 256, 0, st>>>(Wt, Ot, s, N, K, rpb);
      else col_scale_kernel<T, false, true><<<grid, 256, 0, st>>>(Wt, Ot, s, N, K, rpb);
    } else {
      if (mul) col_scale_kernel<T, true, false><<<grid, 256, 0, st>>>(Wt, Ot, s, N, K, rpb);
      else col_scale_kernel<T, false, false><<<grid, 256, 0, st>>>(Wt, Ot, s, N, K, rpb);
    }
  });
  count_launch();
  return check_launch("col_scale");
}

static int act_chunks(int64_t T, int64_t K, int n_samples, int vec, int* rpb_out) {
  const int64_t col_tiles = (K + 32 * vec - 1) / (32 * vec);
  int rpb = pick_rows_per_block(T, col_tiles * n_samples);
  *rpb_out = rpb;
  return (int)((T + rpb - 1) / rpb);
}

int64_t b200q_act_stat_workspace(int64_t T, int64_t K) {
  // one fp32 partial row per (sample, row chunk): callers of the batched form multiply by n_samples
  int rpb;
  const int chunks = act_chunks(T, K, 1, 4, &rpb);
  return (int64_t)sizeof(float) * K * (chunks + 1);
}

// out[s, k] = mean over the rows_per_sample rows of sample s of |X|; X is [n_samples *
// rows_per_sample, K].  One launch pair for the whole batch.
int b200q_act_meanabs_batched(const void* X, int n_samples, int64_t rows_per_sample, int64_t K,
                              int dtype, float* out, void* work, void* stream) {
  B200Q_REQUIRE(X && out && work && n_samples > 0 && rows_per_sample > 0 && K > 0,
                "act_meanabs: bad argument");
  B200Q_REQUIRE(n_samples <= 65535, "act_meanabs: too many samples");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t T = rows_per_sample;
  side_stream_kernels_share_the_gemm_carveout();
  KernelScope scope("act_meanabs", (double)n_samples * T * K * elem_size(dtype), 0, st);
  B200Q_DISPATCH_DTYPE(dtype, Tt, {
    constexpr int VEC = ST<Tt>::VEC;
    B200Q_REQUIRE(aligned16(X) && K % VEC == 0, "act_meanabs: K must be a multiple of 16 bytes");
    const int64_t col_tiles = (K + 32 * VEC - 1) / (32 * VEC);
    int rpb;
    // chunk count from the single-sample rule so that the workspace formula holds per sample
    const int chunks = act_chunks(T, K, 1, 4, &rpb);
    dim3 grid((unsigned)col_tiles, (unsigned)chunks, (unsigned)n_samples);
    act_abssum_partial_kernel<Tt><<<grid, 256, 0, st>>>(static_cast<const Tt*>(X), T, K, rpb,
                                                        static_cast<float*>(work));
    dim3 g2((unsigned)((K + 255) / 256), (unsigned)n_samples);
    act_mean_finish_kernel<Tt><<<g2, 256, 0, st>>>(static_cast<const float*>(work), chunks, K,
                                                   (float)T, out);
  });
  count_launch(2);
  return check_launch("act_meanabs");
}

int b200q_act_meanabs(const void* X, int64_t T, int64_t K, int dtype, float* out, void* work,
                      void* stream) {
  return b200q_act_meanabs_batched(X, 1, T, K, dtype, out, work, stream);
}

int b200q_act_maxabs(const void* X, int64_t T, int64_t K, int dtype, float* out, int accumulate,
                     void* stream) {
  // max|x| over tokens is the same reduction as the weight column |max|
  return b200q_col_absmax(X, T, K, K, dtype, out, accumulate, stream);
}

int b200q_seq_sum_rows(const void* V, int64_t n, int64_t K, int dtype, float* out, void* stream) {
  B200Q_REQUIRE(V && out && n >= 0 && K > 0, "seq_sum_rows: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200Q_DISPATCH_DTYPE(dtype, T,
                       (seq_sum_rows_kernel<T><<<(unsigned)((K + 255) / 256), 256, 0, st>>>(
                           static_cast<const T*>(V), n, K, out)));
  count_launch();
  return check_launch("seq_sum_rows");
}

int b200q_topk_colmul(const float* importance, int64_t K, int64_t k, float factor, float* colmul,
                      uint8_t* mask, void* stream) {
  B200Q_REQUIRE(importance && colmul && K > 0 && k >= 0 && k <= K && K < (1ll << 30),
                "topk_colmul: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  topk_colmul_kernel<<<1, 1024, 0, st>>>(importance, (int)K, (int)k, factor, 1.0f / factor, colmul,
                                         nullptr, mask);
  count_launch();
  return check_launch("topk_colmul");
}

// ---- whole-layer entry points: every launch of one Linear behind ONE host call ----------------
int b200q_awq_layer(const void* W, void* out, int64_t N, int64_t K, int64_t group, int n_bit,
                    const void* feats, int64_t n_feats, int feat_dtype, int64_t n_protect,
                    float scale_factor, float* work, uint8_t* salient_mask, int dtype,
                    void* stream) {
  B200Q_REQUIRE(feats && work && n_feats > 0, "awq_layer: bad argument");
  B200Q_REQUIRE(K > 0 && n_protect >= 0 && n_protect <= K && K < (1ll << 30), "awq_layer: bad K");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* importance = work;          // [K]
  float* colmul = work + K;          // [K]   scale_factor on the salient columns, 1 elsewhere
  float* colrcp = work + 2 * K;      // [K]   RN(1 / colmul)
  int rc = b200q_seq_sum_rows(feats, n_feats, K, feat_dtype, importance, stream);
  if (rc != B200Q_OK) return rc;
  topk_colmul_kernel<<<1, 1024, 0, st>>>(importance, (int)K, (int)n_protect, scale_factor,
                                         1.0f / scale_factor, colmul, colrcp, salient_mask);
  count_launch();
  rc = check_launch("awq_layer/topk");
  if (rc != B200Q_OK) return rc;
  return group_fakequant_impl(W, out, nullptr, nullptr, nullptr, N, K, group, n_bit, 0,
                              B200Q_COLOP_MUL_DIV, colmul, colrcp, 1.0f / scale_factor, dtype,
                              stream);
}

int b200q_gptq_parity_layer(const void* W, void* out, int64_t N, int64_t K, int n_bit,
                            float* colmax, int dtype, void* stream) {
  int rc = b200q_col_absmax(W, N, K, K, dtype, colmax, 0, stream);
  if (rc != B200Q_OK) return rc;
  return b200q_gptq_parity_quant(W, out, nullptr, colmax, nullptr, N, K, K, n_bit, dtype, stream);
}

int b200q_smoothquant_layer(const void* W, void* out, int64_t N, int64_t K, int64_t group,
                            int n_bit, const float* act_scale, float alpha, int act_dtype,
                            float* s, float* work, int dtype, void* stream) {
  B200Q_REQUIRE(act_scale && s && work && K > 0, "smoothquant_layer: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* colmax = work;              // [K]
  float* s_rcp = work + K;           // [K]   RN(1 / s)
  int rc = b200q_col_absmax(W, N, K, K, dtype, colmax, 0, stream);
  if (rc != B200Q_OK) return rc;
  smooth_scale_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(
      act_scale, colmax, s, s_rcp, K, alpha, (float)(1.0 - (double)alpha), act_dtype, dtype,
      promote_dtype(act_dtype, dtype));
  count_launch();
  rc = check_launch("smoothquant_layer/scale");
  if (rc != B200Q_OK) return rc;
  return group_fakequant_impl(W, out, nullptr, nullptr, nullptr, N, K, group, n_bit, 0,
                              B200Q_COLOP_DIV, s, s_rcp, 0.f, dtype, stream);
}

// ---- self test: Divisor::div against __fdiv_rn ------------------------------------------------
}  // extern "C"

namespace b200q {
__global__ void selftest_div_kernel(uint64_t seed, int64_t n_per_thread, unsigned long long* bad,
                                    float* first_a, float* first_b) {
  uint64_t x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1);
  auto next = [&]() {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return x;
  };
  for (int64_t it = 0; it < n_per_thread; ++it) {
    const uint64_t r = next();
    // b: random mantissa, exponent in [-30, 30]; reused for 8 numerators like a group scale
    const uint32_t bb = ((uint32_t)(r & 0x7fffff)) | ((uint32_t)(127 - 30 + (r >> 23) % 61) << 23);
    const float b = __uint_as_float(bb);
    const Divisor d(b);
    for (int k = 0; k < 8; ++k) {
      const uint64_t q = next();
      float a;
      if (k < 4) {
        // random numerator, exponent in [-40, 40], random sign
        const uint32_t ab = ((uint32_t)(q & 0x7fffff)) | ((uint32_t)(127 - 40 + (q >> 23) % 81) << 23) |
                            ((uint32_t)(q >> 63) << 31);
        a = __uint_as_float(ab);
      } else {
        // adversarial: a ~ b * (m + 0.5) +- a few ulps -> quotients next to rounding ties
        const float m = (float)((q >> 8) % 64) + 0.5f;
        a = __uint_as_float(__float_as_uint(b * m) + (int)(q & 7) - 3);
      }
      const float want = __fdiv_rn(a, b);
      const float got = d.div(a);
      if (__float_as_uint(want) != __float_as_uint(got)) {
        if (atomicAdd(bad, 1ull) == 0) { *first_a = a; *first_b = b; }
      }
    }
  }
}
}  // namespace b200q

extern "C" {

int b200q_selftest_div(int64_t n_quotients, uint64_t seed, int64_t* mismatches, void* stream) {
  B200Q_REQUIRE(mismatches != nullptr && n_quotients > 0, "selftest_div: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* d_bad = nullptr;
  float* d_first = nullptr;
  if (cudaMalloc(&d_bad, sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&d_first, 2 * sizeof(float)) != cudaSuccess)
    return fail(B200Q_ECUDA, "selftest_div: cudaMalloc");
  cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st);
  const int blocks = kNumSMs * 8, threads = 256;
  const int64_t per_thread = (n_quotients / 8 + (int64_t)blocks * threads - 1) / ((int64_t)blocks * threads);
  b200q::selftest_div_kernel<<<blocks, threads, 0, st>>>(seed, per_thread, d_bad, d_first, d_first + 1);
  count_launch();
  unsigned long long h_bad = 0;
  float h_first[2] = {0, 0};
  cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(h_first, d_first, sizeof(h_first), cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d_bad);
  cudaFree(d_first);
  if (e != cudaSuccess) return fail(B200Q_ECUDA, std::string("selftest_div: ") + cudaGetErrorString(e));
  *mismatches = (int64_t)h_bad;
  if (h_bad != 0) {
    char buf[128];
    snprintf(buf, sizeof(buf), "selftest_div: first mismatch a=%.9g b=%.9g", h_first[0], h_first[1]);
    set_error(buf);
  }
  return B200Q_OK;
}

static int64_t smooth_alpha_warps(int64_t n_groups) {
  return std::min<int64_t>(n_groups, (int64_t)kNumSMs * 8 * 8);
}

int64_t b200q_smooth_alpha_workspace(int64_t N, int64_t K, int64_t group, int n_alpha) {
  if (N <= 0 || K <= 0 || n_alpha <= 0) return 0;
  const int64_t G = group > 0 ? group : K;
  if (K % G != 0) return 0;
  const int64_t warps = (smooth_alpha_warps(N * (K / G)) + 7) / 8 * 8;
  return (int64_t)sizeof(double) * warps * n_alpha + 256;
}

int b200q_smooth_alpha_errors(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                              const float* S, int n_alpha, const float* act_weight, int dtype,
                              void* work, double* err, int accumulate, void* stream) {
  B200Q_REQUIRE(W && S && act_weight && work && err, "smooth_alpha_errors: null pointer");
  B200Q_REQUIRE(N > 0 && K > 0 && n_alpha > 0 && n_alpha <= 32, "smooth_alpha_errors: bad shape (1..32 alphas)");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 16, "smooth_alpha_errors: n_bit must be in [1,16]");
  const int64_t G = group > 0 ? group : K;
  B200Q_REQUIRE(K % G == 0, "smooth_alpha_errors: in_features not divisible by group size");
  B200Q_REQUIRE((reinterpret_cast<uintptr_t>(work) & 7u) == 0, "smooth_alpha_errors: unaligned workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n_groups = N * (K / G);
  const int blocks = (int)((smooth_alpha_warps(n_groups) + 7) / 8);
  const int64_t nwarps = (int64_t)blocks * 8;
  double* partial = static_cast<double*>(work);
  KernelScope scope("smooth_alpha_errors", (double)N * K * elem_size(dtype), 0, st);
  const float maxint = (float)((1 << n_bit) - 1);
  const bool fast128 = G == 128 && aligned16(W) && aligned16(S) && aligned16(act_weight) && K % 4 == 0;
  if (fast128) {
    B200Q_DISPATCH_DTYPE(dtype, T,
                         (smooth_alpha_err128_kernel<T><<<blocks, 256, 0, st>>>(
                             static_cast<const T*>(W), S, act_weight, n_groups, K, maxint, n_alpha,
                             partial)));
  } else {
    B200Q_DISPATCH_DTYPE(dtype, T,
                         (smooth_alpha_err_kernel<T><<<blocks, 256, 0, st>>>(
                             static_cast<const T*>(W), S, act_weight, n_groups, G, K, maxint, n_alpha,
                             partial)));
  }
  smooth_alpha_reduce_kernel<<<n_alpha, 256, 0, st>>>(partial, nwarps, n_alpha, err, accumulate);
  count_launch(2);
  return check_launch("smooth_alpha_errors");
}

}  // extern "C"
