// Shared device/host helpers for libb200quant (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

#include "../../include/b200quant.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200quant is written for sm_100a (B200) only"
#endif

namespace b200q {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void count_launch(int n = 1);
int check_launch(const char* what);
int64_t launches_so_far();
// While a thread records launches into a CUDA graph (linalg.cu) the event profiler stays out of the
// way: events recorded during capture would become graph nodes.
void set_stream_capture(bool on);
bool in_stream_capture();

// Brackets the launches of one C-ABI call with CUDA events when profiling is on (core.cu).
class KernelScope {
 public:
  KernelScope(const char* name, double bytes, double flops, cudaStream_t st);
  ~KernelScope();
  KernelScope(const KernelScope&) = delete;
  KernelScope& operator=(const KernelScope&) = delete;

 private:
  const char* name_;
  double bytes_, flops_;
  cudaStream_t st_;
  cudaEvent_t e0_, e1_;
};

#define B200Q_REQUIRE(cond, msg)                                 \
  do {                                                           \
    if (!(cond)) return ::b200q::fail(B200Q_EINVAL, (msg));      \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- storage types ---------------------------------------------------------------------------
// torch evaluates an elementwise op on fp16/bf16 tensors as: widen to fp32, operate, round to the
// storage type.  rnd<T>() is that final rounding; for fp32 it is the identity, which makes the
// fp32 kernels bit-identical to torch's.
template <typename T>
struct ST;
template <>
struct ST<float> {
  static constexpr int VEC = 4;
  static __device__ __forceinline__ float rnd(float x) { return x; }
};
template <>
struct ST<__half> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float rnd(float x) { return __half2float(__float2half_rn(x)); }
};
template <>
struct ST<__nv_bfloat16> {
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float rnd(float x) {
    return __bfloat162float(__float2bfloat16_rn(x));
  }
};

__device__ __forceinline__ float rnd_rt(float x, int dtype) {
  if (dtype == B200Q_F16) return ST<__half>::rnd(x);
  if (dtype == B200Q_BF16) return ST<__nv_bfloat16>::rnd(x);
  return x;
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ float from_f<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}

// ---- 128-bit streaming loads / stores ------------------------------------------------------
// Weights are read once and written once: keep them out of L1.
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T>
__device__ __forceinline__ void unpack16(uint4 raw, float (&v)[ST<T>::VEC]);
template <>
__device__ __forceinline__ void unpack16<float>(uint4 raw, float (&v)[4]) {
  v[0] = __uint_as_float(raw.x);
  v[1] = __uint_as_float(raw.y);
  v[2] = __uint_as_float(raw.z);
  v[3] = __uint_as_float(raw.w);
}
template <>
__device__ __forceinline__ void unpack16<__half>(uint4 raw, float (&v)[8]) {
  const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = *reinterpret_cast<const __half2*>(&u[i]);
    float2 f = __half22float2(h);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void unpack16<__nv_bfloat16>(uint4 raw, float (&v)[8]) {
  const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

template <typename T>
__device__ __forceinline__ uint4 pack16(const float (&v)[ST<T>::VEC]);
template <>
__device__ __forceinline__ uint4 pack16<float>(const float (&v)[4]) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                    __float_as_uint(v[3]));
}
template <>
__device__ __forceinline__ uint4 pack16<__half>(const float (&v)[8]) {
  uint32_t u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    u[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}
template <>
__device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float (&v)[8]) {
  uint32_t u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    u[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[ST<T>::VEC]) {
  unpack16<T>(ld_stream16(p), v);
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[ST<T>::VEC]) {
  st_stream16(p, pack16<T>(v));
}

// clamp(x, lo, hi) as torch evaluates it: min(max(x, lo), hi)
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  return fminf(fmaxf(x, lo), hi);
}

// ---- exact division by a reused divisor ---------------------------------------------------------
// a / b correctly rounded (bit-identical to IEEE / torch's CPU division) when the same b divides
// many values: y = RN(1/b) once, then q0 = a*y and two Markstein corrections
//     r = fma(-b, q, a);  q = fma(r, y, q)
// The first makes q faithful, the second (Markstein's theorem: y correctly rounded, q faithful, r
// exact) makes it the correctly rounded quotient — 5 FP32-pipe ops against the ~15-instruction
// __fdiv_rn sequence, which is what moves the fake-quant kernels from issue-bound to HBM-bound.
// Operands outside [1e-18, 1e18] (where r could underflow) take __fdiv_rn.  b200q_selftest_div
// checks the identity against __fdiv_rn on the GPU.
struct Divisor {
  float b, y;
  bool fast;
  __device__ __forceinline__ Divisor() : b(1.f), y(1.f), fast(true) {}
  __device__ __forceinline__ explicit Divisor(float b_) {
    b = b_;
    y = __frcp_rn(b_);
    const float ab = fabsf(b_);
    fast = ab > 1e-18f && ab < 1e18f;
  }
  __device__ __forceinline__ Divisor(float b_, float y_) : b(b_), y(y_) {
    const float ab = fabsf(b_);
    fast = ab > 1e-18f && ab < 1e18f;
  }
  // the unguarded core: exact for |a| and |b| in [1e-18, 1e18]; for smaller |a| the quotient may
  // be 1-2 ulp off (fine where it is rounded to an integer code and |a / b| < 1/4 anyway)
  __device__ __forceinline__ float div_core(float a) const {
    float q = a * y;
    float r = fmaf(-b, q, a);
    q = fmaf(r, y, q);
    r = fmaf(-b, q, a);
    return fmaf(r, y, q);
  }
  __device__ __forceinline__ float div(float a) const {
    const float aa = fabsf(a);
    if (fast && ((aa > 1e-18f && aa < 1e18f) || aa == 0.f)) {
      float q = a * y;
      float r = fmaf(-b, q, a);
      q = fmaf(r, y, q);
      r = fmaf(-b, q, a);
      return fmaf(r, y, q);
    }
    return __fdiv_rn(a, b);
  }
};

// round-half-to-even on the FP32 pipe (rintf is a quarter-rate XU-pipe conversion): two FADDs.
// Exact for |x| < 2^22.  Larger |x| come back within 2 of x, i.e. still far beyond any code range
// (<= 2^16), so this is ONLY for values that are clamped to the code range right after.
__device__ __forceinline__ float rint_then_clamped(float x) {
  const float magic = 12582912.f;  // 1.5 * 2^23
  return (x + magic) - magic;
}

// dtype dispatch on the host
#define B200Q_DISPATCH_DTYPE(dtype, T, ...)                            \
  switch (dtype) {                                                     \
    case B200Q_F32: {                                                  \
      using T = float;                                                 \
      __VA_ARGS__;                                                     \
      break;                                                           \
    }                                                                  \
    case B200Q_F16: {                                                  \
      using T = __half;                                                \
      __VA_ARGS__;                                                     \
      break;                                                           \
    }                                                                  \
    case B200Q_BF16: {                                                 \
      using T = __nv_bfloat16;                                         \
      __VA_ARGS__;                                                     \
      break;                                                           \
    }                                                                  \
    default:                                                           \
      return ::b200q::fail(B200Q_EINVAL, "unknown dtype");             \
  }

inline int elem_size(int dtype) { return dtype == B200Q_F32 ? 4 : 2; }

// elementwise.cu: dW_c = Q_c(W) - W (bf16) for all AWQ search candidates; D + c * cand_stride
int launch_awq_delta(const void* W, void* D, const uint8_t* salient, int64_t N, int64_t K,
                     int64_t group, int64_t cand_stride, int n_bit, const float* sf_host, int n_cand,
                     int dtype, cudaStream_t st);

// host tables of torch-CPU log2 semantics (log2_tables.cpp)
const uint32_t* log2_round_thresholds(int dtype);  // index e+127, e in [-127,127]
const uint32_t* log2_floor_thresholds(int dtype);  // index e+149, e in [-149,127]

}  // namespace b200q
