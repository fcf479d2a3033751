// fp32-accurate GEMM on the tcgen05 tensor cores for the blocked Cholesky inverse (linalg.cu).
//
// The factorisation of the GPTQ Hessian needs fp32 mantissas (the factor later propagates
// quantisation errors), which no tensor-core input format carries.  Each fp32 operand is therefore
// split into two fp16 planes under a per-operand power-of-two scale 2^s chosen so the largest
// magnitude lands in [2^13, 2^14):
//        x * 2^s = h + l + eps,   h = fp16(x 2^s),  l = fp16(x 2^s - h),   |eps| <= 2^-22 |x 2^s|
// (fp16 subnormals keep the ABSOLUTE error of l below 2^-25, i.e. 2^-38 of the operand's largest
// entry, so small entries lose nothing that matters norm-wise) and the product is formed as
//        A B^T  ~=  Ah Bh^T + Ah Bl^T + Al Bh^T        (the dropped Al Bl^T term is 2^-22 relative)
// with every partial product exact in the fp32 accumulator.  That is three k-blocks of plain fp16
// MMAs per k-block of the product: the kernel below is the same TMA -> 4-stage mbarrier ring ->
// tcgen05.mma -> TMEM pipeline as the AWQ search GEMM (tensorcore.cu), with the producer cycling
// through the (h,h), (h,l), (l,h) plane pairs.  Both operands are K-major [rows, k]; the split
// pass writes an operand transposed when the caller needs op(X) = X^T, which costs nothing extra
// since the pass has to run anyway.
//
// With THREE planes (x 2^s = h + m + l, 33 bits) and the six pairs (h,h) (h,m) (m,h) (m,m) (h,l)
// (l,h) the representation error drops to 2^-33 and only the fp32 accumulation error remains, as
// in a SIMT fp32 GEMM; the two-plane form is about ten times less accurate and twice as fast.
//
// Triangular structure is expressed as k-ranges per output tile (KB_* / KE_* flags) and by
// skipping tiles above the diagonal; SYMM writes the lower triangle and mirrors it.
#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "sm100.cuh"
#include "splitgemm.cuh"

namespace b200q {
using namespace sm100;

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
region_absmax_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols,
                     unsigned* __restrict__ max_bits) {
  float m = 0.f;
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    m = fmaxf(m, fabsf(src[r * ld + c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));
}

__device__ __forceinline__ float split_scale(unsigned max_bits) {
  const float big = __uint_as_float(max_bits);
  int e = 0;
  if (big > 0.f && isfinite(big)) {
    frexpf(big, &e);                       // big = m 2^e, m in [0.5, 1)
    e = 14 - e;
  }
  e = max(-100, min(100, e));
  return ldexpf(1.f, e);
}

__device__ __forceinline__ void store_split(__half* dst, int64_t plane_stride, int planes, float v) {
  const __half h = __float2half_rn(v);
  dst[0] = h;
  v -= __half2float(h);                                   // exact
  const __half m = __float2half_rn(v);
  dst[plane_stride] = m;
  if (planes > 2) dst[2 * plane_stride] = __float2half_rn(v - __half2float(m));
}

// planes[0] = h, planes[1] = l of  op(src) * 2^s;  op = transpose ? src^T : src.
// src region is [rows, cols] (ld); the planes are [rows, cols] or [cols, rows] with row stride ld16.
template <bool TRANSPOSE>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols,
                    const unsigned* __restrict__ max_bits, __half* __restrict__ p0,
                    int64_t plane_stride, int planes, int64_t ld16, float* __restrict__ unscale) {
  const float sc = split_scale(*max_bits);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *unscale = 1.f / sc;
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if constexpr (!TRANSPOSE) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      if (r < rows && c < cols)
        store_split(p0 + (int64_t)r * ld16 + c, plane_stride, planes, src[(int64_t)r * ld + c] * sc);
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < rows && c < cols) ? src[(int64_t)r * ld + c] * sc : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;          // output row = source column
      if (r < rows && c < cols)
        store_split(p0 + (int64_t)c * ld16 + r, plane_stride, planes, tile[tx][i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------
namespace sg {
constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;     // [0, 256) chunk accumulator, [256, 512) running total
constexpr int CHUNK_KB = 4;             // k-blocks accumulated inside the tensor core at a time
}  // namespace sg

__global__ void __launch_bounds__(sg::THREADS, 1)
split_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b0,
                  const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                  int pairs, int M, int N, int Kd, float alpha, const float* __restrict__ unscale_a,
                  const float* __restrict__ unscale_b, float beta, float* __restrict__ C, int64_t ldc,
                  int flags) {
  using namespace sg;
  const int m_blk = blockIdx.y, n_blk = blockIdx.x;
  const int m0 = m_blk * BM, n0 = n_blk * BN;
  // tiles strictly above the diagonal of a lower-triangular / symmetric result do not run
  if ((flags & (SG_LOWER | SG_SYMM)) && n0 > m0 + BM - 1) return;
  int k_lo = 0, k_hi = Kd;
  if (flags & SG_KB_M) k_lo = max(k_lo, m0);
  if (flags & SG_KB_N) k_lo = max(k_lo, n0);
  if (flags & SG_KE_M) k_hi = min(k_hi, m0 + BM);
  if (flags & SG_KE_N) k_hi = min(k_hi, n0 + BN);
  const int kb0 = k_lo / BK;
  const int kb1 = (k_hi + BK - 1) / BK;
  const int num_it = max(0, kb1 - kb0) * pairs;       // 3 (two planes) or 6 (three) pairs per k-block
  // The tensor core adds into its fp32 accumulator with TRUNCATION, a bias of half an ulp per MMA
  // that grows linearly with the k extent (measured: 6e-5 relative at k = 4096, the same for two
  // and three planes).  So the MMAs only accumulate CHUNK_KB k-blocks at a time; each finished
  // chunk is added to a running total by the epilogue warps on the FP32 pipe (round to nearest)
  // and the total lives in the other half of TMEM.
  const int chunk_it = CHUNK_KB * pairs;
  const int num_chunks = (num_it + chunk_it - 1) / chunk_it;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* chunk_full_bar = empty_bar + STAGES;       // MMA -> epilogue: chunk accumulator complete
  uint64_t* chunk_free_bar = chunk_full_bar + 1;       // epilogue -> MMA: chunk accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(chunk_free_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b0); tma_prefetch_desc(&map_b1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(chunk_full_bar, 1);
    mbar_init(chunk_free_bar, 4);                      // one arrival per epilogue warp
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
      for (int it = 0; it < num_it; ++it) {
        // plane pairs in order of significance: (0,0) (0,1) (1,0) | (1,1) (0,2) (2,0)
        const int kb = kb0 + it / pairs, pair = it % pairs;
        const int sa = (0x201100 >> (4 * pair)) & 15, sb = (0x021010 >> (4 * pair)) & 15;
        const CUtensorMap* ma = sa == 0 ? &map_a0 : (sa == 1 ? &map_a1 : &map_a2);
        const CUtensorMap* mb = sb == 0 ? &map_b0 : (sb == 1 ? &map_b1 : &map_b2);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        tma_load_2d(a_dst, ma, &full_bar[stage], kb * BK, m0);
        tma_load_2d(a_dst + A_BYTES, mb, &full_bar[stage], kb * BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, /*bf16=*/false, /*a_mn=*/false, /*b_mn=*/false);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int ch = 0; ch < num_chunks; ++ch) {
        if (ch > 0) {                                   // the previous chunk has been read out
          mbar_wait(chunk_free_bar, (uint32_t)((ch - 1) & 1));
          tc_fence_after_sync();
        }
        const int it_end = min(num_it, it + chunk_it);
        for (bool first = true; it < it_end; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2, 16, 1024);
            mma_f16_ss(tmem_base, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          mma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(chunk_full_bar);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const float a_eff = alpha * (*unscale_a) * (*unscale_b);
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // all chunks but the last: total (+)= chunk, kept in TMEM columns [256, 512)
    for (int ch = 0; ch + 1 < num_chunks; ++ch) {
      mbar_wait(chunk_full_bar, (uint32_t)(ch & 1));
      tc_fence_after_sync();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32], t[32];
        tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
        if (ch > 0) {
          tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t[j]));
        } else {
          tmem_ld_wait();
        }
        tmem_st_32x32(lane_base + (uint32_t)(BN + c * 32), v);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(chunk_free_bar);
    }
    if (num_chunks > 0) {
      mbar_wait(chunk_full_bar, (uint32_t)((num_chunks - 1) & 1));
      tc_fence_after_sync();
    }
    const bool symm = (flags & SG_SYMM) != 0;
    const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int col0 = n0 + c * 32;
      if (col0 >= N) break;
      uint32_t v[32];
      if (num_chunks > 0) {
        tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
        if (num_chunks > 1) {
          uint32_t t[32];
          tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t[j]));
        } else {
          tmem_ld_wait();
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      // symmetric result: chunks entirely right of this warp's last row hold nothing to write
      if (symm && col0 > m0 + q * 32 + 31) continue;
      if (row < M) {
        float* dst = C + (int64_t)row * ldc + col0;
        const bool interior = col0 + 32 <= N && (!symm || col0 + 31 <= row);
        if (interior && vec_ok) {
          // the lane's 32 columns are one 128-byte line: 16-byte accesses
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(a_eff * __uint_as_float(v[j]), a_eff * __uint_as_float(v[j + 1]),
                                   a_eff * __uint_as_float(v[j + 2]), a_eff * __uint_as_float(v[j + 3]));
            if (beta != 0.f) {
              const float4 c4 = *reinterpret_cast<const float4*>(dst + j);
              o.x = fmaf(beta, c4.x, o.x); o.y = fmaf(beta, c4.y, o.y);
              o.z = fmaf(beta, c4.z, o.z); o.w = fmaf(beta, c4.w, o.w);
            }
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col >= N || (symm && col > row)) continue;
            const float r = a_eff * __uint_as_float(v[j]);
            dst[j] = (beta == 0.f) ? r : fmaf(beta, dst[j], r);
          }
        }
      }
      if (symm) {
        // mirror image (col, row) for col < row: for a fixed column the lanes hold 32 consecutive
        // rows, so each store instruction writes one 128-byte line
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (row < M && col < N && col < row) {
            float* d2 = C + (int64_t)col * ldc + row;
            const float r = a_eff * __uint_as_float(v[j]);
            *d2 = (beta == 0.f) ? r : fmaf(beta, *d2, r);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static EncodeTiledFn encode_fn() {
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

// fp16 [rows, cols] with row stride ld16 elements; box = box_rows x 64 columns, 128-byte swizzle
static int plane_map(CUtensorMap* map, const __half* base, int rows, int cols, int64_t ld16,
                     int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld16 * 2};
  cuuint32_t box[2] = {(cuuint32_t)sg::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled (planes) failed");
  return B200Q_OK;
}

int64_t split_operand_bytes(int rows, int cols) {
  const int64_t ld16 = ((int64_t)cols + 7) / 8 * 8;
  return ((kSplitMaxPlanes * (int64_t)rows * ld16 * 2 + 255) / 256 * 256) + 256;
}

int split_operand(cudaStream_t st, const float* src, int64_t ld, int src_rows, int src_cols,
                  bool transpose, int planes, void* buf, SplitOperand* out) {
  if (planes < 2 || planes > kSplitMaxPlanes) return fail(B200Q_EINVAL, "split_operand: planes");
  // layout of buf: [0,4) max bits, [4,8) unscale, [256, ...) plane h then plane l
  uint8_t* base = static_cast<uint8_t*>(buf);
  unsigned* max_bits = reinterpret_cast<unsigned*>(base);
  float* unscale = reinterpret_cast<float*>(base + 4);
  const int rows = transpose ? src_cols : src_rows;
  const int cols = transpose ? src_rows : src_cols;
  const int64_t ld16 = ((int64_t)cols + 7) / 8 * 8;
  __half* ph = reinterpret_cast<__half*>(base + 256);
  const int64_t plane_stride = (int64_t)rows * ld16;
  KernelScope scope("inv_split", (4.0 + 2.0 * planes) * src_rows * src_cols, 0, st);
  cudaMemsetAsync(max_bits, 0, 8, st);
  const int64_t n = (int64_t)src_rows * src_cols;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
  region_absmax_kernel<<<blocks, 256, 0, st>>>(src, ld, src_rows, src_cols, max_bits);
  dim3 grid((unsigned)((src_cols + 31) / 32), (unsigned)((src_rows + 31) / 32));
  if (transpose)
    split_planes_kernel<true><<<grid, 256, 0, st>>>(src, ld, src_rows, src_cols, max_bits, ph,
                                                   plane_stride, planes, ld16, unscale);
  else
    split_planes_kernel<false><<<grid, 256, 0, st>>>(src, ld, src_rows, src_cols, max_bits, ph,
                                                    plane_stride, planes, ld16, unscale);
  count_launch(2);
  for (int i = 0; i < 3; ++i) out->p[i] = ph + std::min(i, planes - 1) * plane_stride;
  out->planes = planes;
  out->ld16 = ld16; out->rows = rows; out->cols = cols; out->unscale = unscale;
  return check_launch("split_operand");
}

int split_gemm_prepare() {
  // (per-device attribute; the call is cheap and idempotent)
  if (cudaFuncSetAttribute(split_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           sg::SMEM_BYTES) != cudaSuccess)
    return fail(B200Q_ECUDA, "split_gemm: cannot raise shared memory");
  return B200Q_OK;
}

int split_gemm(cudaStream_t st, const SplitOperand& A, const SplitOperand& B, float alpha, float beta,
               float* C, int64_t ldc, int flags) {
  const int M = A.rows, N = B.rows, Kd = A.cols;
  if (M <= 0 || N <= 0) return B200Q_OK;
  if (B.cols != Kd) return fail(B200Q_EINVAL, "split_gemm: inner dimensions differ");
  if (A.planes != B.planes) return fail(B200Q_EINVAL, "split_gemm: plane counts differ");
  CUtensorMap ma[3], mb[3];
  int rc;
  for (int i = 0; i < 3; ++i) {
    if ((rc = plane_map(&ma[i], A.p[i], M, Kd, A.ld16, sg::BM)) != B200Q_OK) return rc;
    if ((rc = plane_map(&mb[i], B.p[i], N, Kd, B.ld16, sg::BN)) != B200Q_OK) return rc;
  }
  if (!in_stream_capture() && (rc = split_gemm_prepare()) != B200Q_OK) return rc;
  dim3 grid((unsigned)((N + sg::BN - 1) / sg::BN), (unsigned)((M + sg::BM - 1) / sg::BM));
  KernelScope scope("inv_gemm", 0, 2.0 * M * (double)N * Kd, st);   // dense count; tri flags skip part
  split_gemm_kernel<<<grid, sg::THREADS, sg::SMEM_BYTES, st>>>(
      ma[0], ma[1], ma[2], mb[0], mb[1], mb[2], A.planes == 3 ? 6 : 3, M, N, Kd, alpha, A.unscale,
      B.unscale, beta, C, ldc, flags);
  count_launch();
  return check_launch("split_gemm");
}

}  // namespace b200q
