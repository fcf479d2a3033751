// W4A16 GEMM with the dequantisation fused into the operand pipeline:
//     Y[t, n] = sum_k X[t, k] * ((q[n, k] - zero[n, g(k)]) * scale[n, g(k)])
// X: 16-bit activations [M tokens, K]; q: the 4-bit codes b200q.export packs (eight per int32,
// lowest nibble first) with one fp32 scale / zero point per group -- the "uniform_asym" records
// that pseudo_quantize_tensor / AWQ / SmoothQuant produce (quantization_utils.py:395-407).  This
// is the Linear of the reference's perplexity loop (quantization_utils.py:269-322) evaluated on
// the PACKED weight: 0.5 byte per weight from HBM instead of the 2 bytes of a fake-quantised fp16
// copy.
//
// One CTA per 128 x 128 output tile, 4-stage ring over 64-wide k-blocks:
//   warp 0      TMA producer: X tile (128 tokens x 64 k) -> swizzled smem            (A, K-major)
//   warps 4-7   dequant producers: thread r owns weight row n0 + r of the tile; per k-block it
//               reads the row's 32 bytes of codes, turns them into 64 16-bit weights and writes
//               them into shared memory IN THE 128-BYTE-SWIZZLED K-MAJOR LAYOUT the UMMA
//               descriptor expects (16-byte chunk c of row r lands at chunk c ^ (r & 7)), then
//               fence.proxy.async + one mbarrier arrival per warp                    (B, K-major)
//   warp 1      MMA issuer: tcgen05.mma (128 x 128 x 16, fp32 accumulator in TMEM) once both
//               halves of a stage are full; tcgen05.commit frees the stage
//   warps 4-7   epilogue after the last k-block: TMEM -> registers -> Y.
// int4 -> fp16/bf16 without I2F: (w >> 4j) & 0x000f000f | magic puts two codes into the mantissas
// of 1024 + q (fp16) / 128 + q (bf16); q - zero is then one exact HSUB2.  For 16-bit records the
// product with the 16-bit scale is a single HMUL2 -- bit for bit what export.dequantize computes
// ((q - z) * s, every op rounded to the weight's dtype); for fp32 records the product is formed in
// fp32 and rounded once to the activation dtype.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"

namespace b200q {
using namespace sm100;

namespace qg {
constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;     // 16 KiB
constexpr int B_BYTES = BN * BK * 2;     // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 128;
}  // namespace qg

template <typename T>
struct Pair16;
template <>
struct Pair16<__half> {
  using V2 = __half2;
  static constexpr uint32_t MAGIC = 0x64006400u;            // fp16 1024.0 twice
  static __device__ __forceinline__ V2 bias(float z) { return __float2half2_rn(1024.f + z); }
  static __device__ __forceinline__ V2 from_float(float s) { return __float2half2_rn(s); }
  static __device__ __forceinline__ float2 to_float2(V2 v) { return __half22float2(v); }
  static __device__ __forceinline__ V2 from_float2(float a, float b) { return __floats2half2_rn(a, b); }
};
template <>
struct Pair16<__nv_bfloat16> {
  using V2 = __nv_bfloat162;
  static constexpr uint32_t MAGIC = 0x43004300u;            // bf16 128.0 twice
  static __device__ __forceinline__ V2 bias(float z) { return __float2bfloat162_rn(128.f + z); }
  static __device__ __forceinline__ V2 from_float(float s) { return __float2bfloat162_rn(s); }
  static __device__ __forceinline__ float2 to_float2(V2 v) { return __bfloat1622float2(v); }
  static __device__ __forceinline__ V2 from_float2(float a, float b) { return __floats2bfloat162_rn(a, b); }
};

template <typename V2>
__device__ __forceinline__ uint32_t as_u32(V2 v) { return *reinterpret_cast<uint32_t*>(&v); }
template <typename V2>
__device__ __forceinline__ V2 as_v2(uint32_t u) { return *reinterpret_cast<V2*>(&u); }

// eight 4-bit codes of one packed word -> eight 16-bit weights (four packed pairs, element order)
template <typename T, bool REC_F32>
__device__ __forceinline__ void dequant_word(uint32_t w, float scale, float zero, uint32_t (&out)[4]) {
  using P = Pair16<T>;
  using V2 = typename P::V2;
  const V2 bias = P::bias(zero);                 // 1024 + z (exact: z is an integer <= 15)
  V2 h[4];                                       // h[j] = (code j, code j + 4) - z
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t bits = ((w >> (4 * j)) & 0x000f000fu) | P::MAGIC;
    h[j] = __hsub2(as_v2<V2>(bits), bias);
  }
  if constexpr (REC_F32) {
    // fp32 record: (q - z) * s in fp32, one rounding to the activation dtype
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = P::to_float2(h[j]);
      h[j] = P::from_float2(f.x * scale, f.y * scale);
    }
  } else {
    // 16-bit record: the scale IS a value of that dtype; (q - z) * s rounded to it (export.dequantize)
    const V2 s2 = P::from_float(scale);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __hmul2(h[j], s2);
  }
  // (c0,c4) (c1,c5) (c2,c6) (c3,c7)  ->  (c0,c1) (c2,c3) (c4,c5) (c6,c7)
  out[0] = __byte_perm(as_u32(h[0]), as_u32(h[1]), 0x5410);
  out[1] = __byte_perm(as_u32(h[2]), as_u32(h[3]), 0x5410);
  out[2] = __byte_perm(as_u32(h[0]), as_u32(h[1]), 0x7632);
  out[3] = __byte_perm(as_u32(h[2]), as_u32(h[3]), 0x7632);
}

template <typename T, bool REC_F32, bool OUT_F32>
__global__ void __launch_bounds__(qg::THREADS, 1)
w4a16_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint32_t* __restrict__ qweight,
                  const float* __restrict__ scales, const float* __restrict__ zeros,
                  void* __restrict__ Y, int64_t M, int64_t N, int64_t K, int64_t G) {
  using namespace qg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_b = full_a + STAGES;
  uint64_t* empty_bar = full_b + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int num_kb = (int)((K + BK - 1) / BK);
  const int64_t words_per_row = K / 8;
  const int64_t groups_per_row = K / G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_a[s], 1);
      mbar_init(&full_b[s], 4);                   // one arrival per dequant warp
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_a[stage], A_BYTES);
        tma_load_2d(smem + stage * STAGE_BYTES, &tmap_x, &full_a[stage], kb * BK, (int32_t)m0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, kBf16, /*a_mn=*/false, /*b_mn=*/false);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_a[stage], phase);
        mbar_wait(&full_b[stage], phase);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2, 16, 1024);
          mma_f16_ss(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===== dequant producers: thread r <-> weight row n0 + r =====
    const int r = (warp - 4) * 32 + lane;
    const int64_t n = n0 + r;
    const bool row_ok = n < N;
    const uint32_t* qrow = qweight + (row_ok ? n : 0) * words_per_row;
    const float* srow = scales + (row_ok ? n : 0) * groups_per_row;
    const float* zrow = zeros + (row_ok ? n : 0) * groups_per_row;
    int stage = 0; uint32_t phase = 0;
    int64_t cur_g = -1;
    float sc = 0.f, zp = 0.f;
    for (int kb = 0; kb < num_kb; ++kb) {
      // this row's 8 words of the k-block (two 16-byte loads), fetched before the stage is free
      uint32_t w[8];
      const int64_t w0 = (int64_t)kb * (BK / 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = 0u;
      if (row_ok) {
        if (w0 + 8 <= words_per_row) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(qrow + w0));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(qrow + w0 + 4));
          w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (w0 + i < words_per_row) w[i] = __ldg(qrow + w0 + i);
        }
      }
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* brow = smem + stage * STAGE_BYTES + A_BYTES + r * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t k = (int64_t)kb * BK + i * 8;
        uint32_t o[4] = {0u, 0u, 0u, 0u};
        if (row_ok && k < K) {
          const int64_t g = k / G;
          if (g != cur_g) { cur_g = g; sc = __ldg(srow + g); zp = __ldg(zrow + g); }
          dequant_word<T, REC_F32>(w[i], sc, zp, o);
        }
        // 128-byte swizzle: 16-byte chunk i of row r sits at chunk i ^ (r & 7)
        *reinterpret_cast<uint4*>(brow + ((i ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async_smem();                     // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_b[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // ===== epilogue =====
    const int q = warp & 3;
    const int64_t row = m0 + q * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      const int64_t col0 = n0 + c * 32;
      if (row >= M || col0 >= N) continue;
      if constexpr (OUT_F32) {
        float* dst = static_cast<float*>(Y) + row * N + col0;
        if (col0 + 32 <= N && (N % 4 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = __uint_as_float(v[j]);
        }
      } else {
        T* dst = static_cast<T*>(Y) + row * N + col0;
        if (col0 + 32 <= N && (N % 8 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(v[j + u]);
            *reinterpret_cast<uint4*>(dst + j) = pack16<T>(f);
          }
        } else {
          for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = from_f<T>(__uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

static EncodeTiledFn qg_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace b200q

using namespace b200q;

extern "C" {

// Y[M, N] = X[M, K] * dequant(qweight)[N, K]^T.   ref: the nn.Linear forward inside
// quantization_utils.py:269-322 (perplexity loop), on a packed "uniform_asym" record.
//   X        16-bit activations (act_dtype = B200Q_F16 / B200Q_BF16), row-major, K % 8 == 0
//   qweight  [N, K/8] uint32, eight 4-bit codes per word (b200q_pack_codes layout)
//   scales, zeros  fp32 [N, K/group]; rec_dtype = dtype the record was quantised in
//   Y        [M, N] in act_dtype, or fp32 when out_f32 != 0
int b200q_w4a16_gemm(const void* X, int64_t M, int64_t K, int act_dtype, const uint32_t* qweight,
                     const float* scales, const float* zeros, int64_t N, int64_t group,
                     int rec_dtype, void* Y, int out_f32, void* stream) {
  B200Q_REQUIRE(X && qweight && scales && zeros && Y, "w4a16_gemm: null pointer");
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0, "w4a16_gemm: bad shape");
  B200Q_REQUIRE(act_dtype == B200Q_F16 || act_dtype == B200Q_BF16, "w4a16_gemm: activations must be fp16 or bf16");
  B200Q_REQUIRE(K % 8 == 0, "w4a16_gemm: in_features must be a multiple of 8");
  const int64_t G = (group > 0 && group < K) ? group : K;
  B200Q_REQUIRE(K % G == 0 && G % 8 == 0, "w4a16_gemm: group must divide in_features and be a multiple of 8");
  B200Q_REQUIRE(aligned16(X) && aligned16(qweight) && aligned16(Y), "w4a16_gemm: unaligned pointer");
  B200Q_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "w4a16_gemm: dimension too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeTiledFn enc = qg_encode_fn();
  if (enc == nullptr) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmap;
  {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
    cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)qg::BK, (cuuint32_t)qg::BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, act_dtype == B200Q_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                     2, const_cast<void*>(X), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200Q_ECUDA, "w4a16_gemm: cuTensorMapEncodeTiled failed");
  }
  // algorithmic traffic: activations + packed codes + group parameters + output
  KernelScope scope("w4a16_gemm", 2.0 * M * K + 0.5 * N * K + 8.0 * N * (K / G) + (out_f32 ? 4.0 : 2.0) * M * N,
                    2.0 * M * (double)N * K, st);
  dim3 grid((unsigned)((N + qg::BN - 1) / qg::BN), (unsigned)((M + qg::BM - 1) / qg::BM));
  // a record quantised in another dtype than the activations' (fp32, or fp16 vs bf16): the product
  // is formed in fp32 and rounded once
  const bool rec32 = rec_dtype != act_dtype;
#define B200Q_QG_LAUNCH(T, R, O)                                                                       \
  do {                                                                                                 \
    cudaFuncSetAttribute(w4a16_gemm_kernel<T, R, O>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                         qg::SMEM_BYTES);                                                              \
    w4a16_gemm_kernel<T, R, O><<<grid, qg::THREADS, qg::SMEM_BYTES, st>>>(tmap, qweight, scales, zeros, \
                                                                          Y, M, N, K, G);              \
  } while (0)
  if (act_dtype == B200Q_F16) {
    if (rec32) { if (out_f32) B200Q_QG_LAUNCH(__half, true, true); else B200Q_QG_LAUNCH(__half, true, false); }
    else { if (out_f32) B200Q_QG_LAUNCH(__half, false, true); else B200Q_QG_LAUNCH(__half, false, false); }
  } else {
    if (rec32) { if (out_f32) B200Q_QG_LAUNCH(__nv_bfloat16, true, true); else B200Q_QG_LAUNCH(__nv_bfloat16, true, false); }
    else { if (out_f32) B200Q_QG_LAUNCH(__nv_bfloat16, false, true); else B200Q_QG_LAUNCH(__nv_bfloat16, false, false); }
  }
#undef B200Q_QG_LAUNCH
  count_launch();
  return check_launch("w4a16_gemm");
}

}  // extern "C"
