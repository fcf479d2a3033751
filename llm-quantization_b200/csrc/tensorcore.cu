// Tensor-core stages (tcgen05 + TMEM accumulators, operands staged by TMA):
//   * GPTQ Hessian  H += sum_i a_i^2 X_i^T X_i,  a_i = 1/(||X_i||_F + 1e-5)   ref: gptq_quantizer.py:137-144
//   * AWQ scale search loss  tr(dW H dW^T) per candidate                        ref: awq_quantizer.py:116-119
//
// Hessian pipeline for one Linear (X is [T, K], T tokens of K input channels, samples are equal
// runs of rows):
//   1. sample_stats   per-sample sum of squares and |max|                       (HBM: read X once)
//   2. prescale       Xs = fp16(X * a_i * 2^s), 2^s chosen so the largest value sits near 2^14:
//                     every product of the GEMM is then exact in fp32 and inputs keep 11 bits
//                                                                                (read X, write Xs)
//   3. hessian_gemm   P[z] = Xs_z^T Xs_z on the tensor cores.  Both operands are tiles of the SAME
//                     row-major matrix with the channel (M/N) direction contiguous, i.e. MN-major
//                     UMMA operands: a TMA box of 64 tokens x 64 channels lands as 8x(8 rows x 128 B)
//                     swizzle atoms, exactly the canonical MN-major SWIZZLE_128B layout.  One CTA
//                     per 128x256 output tile and token split z; 4-stage mbarrier ring; warp 0 =
//                     TMA producer, warp 1 = MMA issuer (one thread), warps 4-7 = epilogue
//                     (tcgen05.ld -> fp32 stores).
//   4. reduce         H (+)= 2^-2s * sum_z P[z], splits added in a fixed order (deterministic).
// 16-bit activations skip the staging: the plain Gram matrix (AWQ search) runs stage 3 directly on
// the caller's tensor, and the normalised Hessian of samples of >= 512 rows runs stage 1 and a
// PER_SAMPLE variant of stage 3 that folds each sample in with its weight a_i^2 (see the kernel).
// Only tiles touching the upper triangle run (SYRK); a single split is written straight into H.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "sm100.cuh"

namespace b200q {

using namespace sm100;

// =================================================================================================
// 1. per-sample statistics
// =================================================================================================
// grid = (chunks, n_samples); each block reduces a slice of one sample; partial[(s*chunks + c)*2]
template <typename T>
__global__ void __launch_bounds__(256)
sample_stats_partial_kernel(const T* __restrict__ X, int64_t rows_per_sample, int64_t K,
                            float* __restrict__ partial) {
  constexpr int VEC = ST<T>::VEC;
  const int64_t s = blockIdx.y;
  const int64_t n = rows_per_sample * K;          // elements of this sample (contiguous)
  const T* p = X + s * n;
  const int64_t nvec = n / VEC;
  double acc = 0.0;
  float amax = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four 16-byte loads in flight per thread (HBM-bound pass: T*K*sz bytes read once)
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    float v[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_vec<T>(p + (i + u * stride) * VEC, v[u]);
    float sq = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) { sq = fmaf(v[u][j], v[u][j], sq); amax = fmaxf(amax, fabsf(v[u][j])); }
    acc += (double)sq;
  }
  for (; i < nvec; i += stride) {
    float v[VEC];
    load_vec<T>(p + i * VEC, v);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sq = fmaf(v[j], v[j], sq); amax = fmaxf(amax, fabsf(v[j])); }
    acc += (double)sq;
  }
  // block reduce
  __shared__ double s_acc[256];
  __shared__ float s_max[256];
  s_acc[threadIdx.x] = acc;
  s_max[threadIdx.x] = amax;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_acc[threadIdx.x] += s_acc[threadIdx.x + o];
      s_max[threadIdx.x] = fmaxf(s_max[threadIdx.x], s_max[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* out = partial + (s * gridDim.x + blockIdx.x) * 2;
    out[0] = (float)s_acc[0];
    out[1] = s_max[0];
  }
}

// one block: norms[i] = sqrt(sum), alpha[i] = 1/(norm + 1e-5); global power-of-two factor
// stats layout: [0, n) alpha_i * 2^s ; [n, 2n) norms ; [2n] = 2^-2s
__global__ void sample_stats_finish_kernel(const float* __restrict__ partial, int chunks,
                                           int n_samples, float* __restrict__ stats,
                                           int normalize) {
  __shared__ float s_big[256];
  float big = 0.f;
  for (int s = threadIdx.x; s < n_samples; s += blockDim.x) {
    double sum = 0.0;
    float amax = 0.f;
    for (int c = 0; c < chunks; ++c) {
      sum += (double)partial[((int64_t)s * chunks + c) * 2];
      amax = fmaxf(amax, partial[((int64_t)s * chunks + c) * 2 + 1]);
    }
    const float norm = (float)sqrt(sum);
    const float alpha = normalize ? 1.f / (norm + 1e-5f) : 1.f;   // gptq_quantizer.py:143
    stats[n_samples + s] = norm;
    stats[s] = alpha;
    big = fmaxf(big, alpha * amax);
  }
  s_big[threadIdx.x] = big;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_big[threadIdx.x] = fmaxf(s_big[threadIdx.x], s_big[threadIdx.x + o]);
    __syncthreads();
  }
  big = s_big[0];
  // 2^s: largest scaled magnitude lands in [2^13, 2^14) (fp16 max is 65504)
  int e = 0;
  if (big > 0.f && isfinite(big)) {
    frexpf(big, &e);                                      // big = m * 2^e, m in [0.5, 1)
    e = 14 - e;
  }
  e = max(-60, min(60, e));
  const float pow2 = ldexpf(1.f, e);
  __syncthreads();
  for (int s = threadIdx.x; s < n_samples; s += blockDim.x) stats[s] *= pow2;
  if (threadIdx.x == 0) stats[2 * n_samples] = ldexpf(1.f, -2 * e);
}

// =================================================================================================
// 2. prescale: Xs[t, k] = fp16( X[t, k] * alpha_{sample(t)} * 2^s )
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
prescale_kernel(const T* __restrict__ X, __half* __restrict__ Xs, int64_t rows_per_sample, int64_t K,
                int64_t total_vec, const float* __restrict__ stats) {
  constexpr int VEC = ST<T>::VEC;
  const int64_t vec_per_sample = rows_per_sample * K / VEC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float a = stats[i / vec_per_sample];
    float v[VEC];
    load_vec<T>(X + i * VEC, v);
    if constexpr (VEC == 8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[j] * a;
      store_vec<__half>(Xs + i * 8, o);
    } else {
      __half2 h0 = __floats2half2_rn(v[0] * a, v[1] * a);
      __half2 h1 = __floats2half2_rn(v[2] * a, v[3] * a);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(Xs + i * 4) = pk;
    }
  }
}

// =================================================================================================
// 3. Xs^T Xs on tcgen05
// =================================================================================================
namespace hg {
constexpr int BM = 128;          // output rows per CTA  (channels)
constexpr int BN = 256;          // output cols per CTA  (channels)
constexpr int BKT = 64;          // tokens per pipeline stage
constexpr int UMMA_K = 16;       // tokens per tcgen05.mma (16-bit inputs)
constexpr int STAGES = 4;
constexpr int BOX_CH = 64;       // channels per TMA box = one 128-byte swizzle row
constexpr int BOX_BYTES = BKT * BOX_CH * 2;                 // 8 KiB
constexpr int A_BYTES = (BM / BOX_CH) * BOX_BYTES;          // 16 KiB
constexpr int B_BYTES = (BN / BOX_CH) * BOX_BYTES;          // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;              // 48 KiB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 256;
constexpr int RASTER_M = 16;     // tile rows per rasterisation band
}  // namespace hg

// PER_SAMPLE = chunked accumulation.  The tensor core aligns every product to the accumulator's
// exponent and TRUNCATES it, so a sum of N same-sign terms (the diagonal of X^T X) comes out low by
// about N * 2^-24 relative: 0.2 % after 37 000 tokens (measured).  The MMAs therefore accumulate
// only `kb_per_sample` 64-token blocks at a time and the epilogue warps fold each finished chunk
// into a running total kept in TMEM columns [256, 512) on the FP32 pipe (round to nearest).
// With `norms` given (GPTQ Hessian of 16-bit activations, no staging pass) a chunk is exactly one
// calibration sample and is folded in as  total += chunk / (||x|| + 1e-5)^2 -- the per-sample
// normalisation of gptq_quantizer.py:143 applied to the sample's exact Gram matrix instead of to
// every activation.
template <bool BF16, bool PER_SAMPLE>
__global__ void __launch_bounds__(hg::THREADS, 1)
hessian_gemm_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ partial,
                    int64_t K, int64_t T, int64_t tokens_per_split, int tiles_n,
                    const float* __restrict__ norms, int kb_per_sample) {
  using namespace hg;
  constexpr uint32_t kTmemCols = PER_SAMPLE ? 512u : TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;        // MMA -> epilogue: (chunk) accumulator complete
  uint64_t* chunk_free_bar = tmem_full_bar + 1;        // epilogue -> MMA: chunk accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(chunk_free_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Tile order: bands of RASTER_M tile rows, column-major inside a band.  The ~148 CTAs resident
  // at any time then form a compact ~16 x 9 patch of tiles that touches only ~4.4K distinct
  // channels per token, start together and stream the tokens in lockstep, so every TMA box is
  // fetched from DRAM once per wave and hit in L2 by the other CTAs (ncu at K = 11008: the plain
  // row-major order ran at 38 % L2 hit rate and was DRAM-bound).
  const int tiles_m = (int)((K + BM - 1) / BM);
  const int band = blockIdx.x / (RASTER_M * tiles_n);
  const int within = blockIdx.x % (RASTER_M * tiles_n);
  const int band_rows = min(RASTER_M, tiles_m - band * RASTER_M);
  const int n_blk = within / band_rows;
  const int m_blk = band * RASTER_M + within % band_rows;
  const int z = blockIdx.y;
  // X^T X is symmetric: a tile whose columns all lie left of its first row holds nothing of the
  // upper triangle and is skipped (47 % of the tiles at K = 4096); the tiles that run write every
  // element with col >= row and its mirror image, so each output element is written exactly once.
  if ((int64_t)(n_blk + 1) * BN <= (int64_t)m_blk * BM) return;
  const int64_t t0 = (int64_t)z * tokens_per_split;
  const int64_t t1 = min(T, t0 + tokens_per_split);
  const int num_kb = (int)((t1 - t0 + BKT - 1) / BKT);
  // chunks of the k loop that are accumulated inside the tensor core: the whole loop, or one sample
  const int chunk_kb = PER_SAMPLE ? kb_per_sample : max(num_kb, 1);
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  const int sample0 = PER_SAMPLE ? (int)(t0 / ((int64_t)kb_per_sample * BKT)) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(chunk_free_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* a_dst = smem + stage * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        const int32_t tok = (int32_t)(t0 + (int64_t)kb * BKT);
#pragma unroll
        for (int j = 0; j < BM / BOX_CH; ++j)
          tma_load_2d(a_dst + j * BOX_BYTES, &tmap, &full_bar[stage],
                      m_blk * BM + j * BOX_CH, tok);
#pragma unroll
        for (int j = 0; j < BN / BOX_CH; ++j)
          tma_load_2d(b_dst + j * BOX_BYTES, &tmap, &full_bar[stage],
                      n_blk * BN + j * BOX_CH, tok);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, BF16, /*a_mn=*/true, /*b_mn=*/true);
      int stage = 0;
      uint32_t phase = 0;
      int kb = 0;
      for (int ch = 0; ch < num_chunks; ++ch) {
        if (PER_SAMPLE && ch > 0) {                  // the previous sample has been folded in
          mbar_wait(chunk_free_bar, (uint32_t)((ch - 1) & 1));
          tc_fence_after_sync();
        }
        const int kb_end = min(num_kb, kb + chunk_kb);
        for (bool first = true; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BKT / UMMA_K; ++k) {
            // 16 tokens = 16 rows of 128 bytes further down every box
            const uint32_t koff = k * UMMA_K * 128;
            const uint64_t da = make_smem_desc_sw128(a_addr + koff, BOX_BYTES, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + koff, BOX_BYTES, 1024);
            mma_f16_ss(tmem_base, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          mma_commit(&empty_bar[stage]);          // frees the smem stage once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(tmem_full_bar);                // (chunk) accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global (fp32) =====
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    float* dst_base = partial + (int64_t)z * K * K;
    const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    auto sample_weight = [&](int ch) {
      if (norms == nullptr) return 1.f;                         // plain chunked accumulation
      const float a = 1.f / (norms[sample0 + ch] + 1e-5f);     // gptq_quantizer.py:143
      return a * a;
    };
    if (PER_SAMPLE) {
      for (int ch = 0; ch + 1 < num_chunks; ++ch) {
        mbar_wait(tmem_full_bar, (uint32_t)(ch & 1));
        tc_fence_after_sync();
        const float wgt = sample_weight(ch);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32], t[32];
          tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
          if (ch > 0) {
            tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(wgt, __uint_as_float(v[j]), __uint_as_float(t[j])));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(wgt * __uint_as_float(v[j]));
          }
          tmem_st_32x32(lane_base + (uint32_t)(BN + c * 32), v);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(chunk_free_bar);
      }
    }
    if (num_chunks > 0) {
      mbar_wait(tmem_full_bar, (uint32_t)((num_chunks - 1) & 1));
      tc_fence_after_sync();
    }
    const float last_wgt = (PER_SAMPLE && num_chunks > 0) ? sample_weight(num_chunks - 1) : 1.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (num_chunks > 0) {
        tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
        if (PER_SAMPLE) {
          if (num_chunks > 1) {
            uint32_t t[32];
            tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(last_wgt, __uint_as_float(v[j]), __uint_as_float(t[j])));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(last_wgt * __uint_as_float(v[j]));
          }
        } else {
          tmem_ld_wait();
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int64_t col0 = (int64_t)n_blk * BN + c * 32;
      // chunks entirely left of this warp's first row carry no upper-triangle element
      const int64_t warp_row0 = (int64_t)m_blk * BM + q * 32;
      if (col0 + 32 <= warp_row0) continue;
      if (row < K) {
        float* dst = dst_base + row * K + col0;
        if (col0 >= row && col0 + 32 <= K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j)
            if (col0 + j >= row && col0 + j < K) dst[j] = __uint_as_float(v[j]);
        }
      }
      // mirror image: element (col, row) for col > row.  For a fixed column the 32 lanes of the
      // warp hold 32 consecutive rows, so each of these stores is one coalesced 128-byte line.
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t col = col0 + j;
        if (col > row && col < K && row < K) dst_base[col * K + row] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// =================================================================================================
// 3b. the same product on CTA PAIRS (cta_group::2)
// =================================================================================================
// Two CTAs of a cluster (one TPC) compute one 256 x 256 output tile: the leader issues
// tcgen05.mma.cta_group::2 with M = 256, each CTA stages ITS 128 rows of A and ITS 128 of the 256
// columns of B (32 KiB per stage instead of 48) and ends up with its 128 x 256 slice of the
// accumulator in its own TMEM.  Per flop that is half the shared-memory operand traffic and a
// third less L2 -> SM traffic than the one-CTA kernel -- what matters here, because the step runs
// at the 1 kW power limit and every byte not moved is clock.  Six stages fit.  Everything else
// (SYRK tile skipping, band rasterisation, chunked accumulation with the running total in TMEM
// columns [256, 512), per-sample weights, mirror-writing epilogue) is as in hessian_gemm_kernel.
namespace hg2 {
constexpr int BM = 256;          // output rows per CTA pair (128 per CTA)
constexpr int BN = 256;          // output cols per CTA pair (each CTA stages 128 of them)
constexpr int BKT = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 6;
constexpr int BOX_CH = 64;
constexpr int BOX_BYTES = BKT * BOX_CH * 2;                 // 8 KiB
constexpr int A_BYTES = 2 * BOX_BYTES;                      // this CTA's 128 channels of A
constexpr int B_BYTES = 2 * BOX_BYTES;                      // this CTA's 128 channels of B
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;              // 32 KiB per CTA
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 256;
constexpr int RASTER_M_DEFAULT = 8;   // 256-row tile rows per rasterisation band
}  // namespace hg2

template <bool BF16, bool PER_SAMPLE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(hg2::THREADS, 1)
hessian_gemm2_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ partial,
                     int64_t K, int64_t T, int64_t tokens_per_split, int tiles_n,
                     const float* __restrict__ norms, int kb_per_sample, int raster_m) {
  using namespace hg2;
  const int RASTER_M = raster_m;
  constexpr uint32_t kTmemCols = PER_SAMPLE ? 512u : 256u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // used in the leader
  uint64_t* empty_bar = full_bar + STAGES;             // one per CTA, fed by the multicast commit
  uint64_t* tmem_full_bar = empty_bar + STAGES;        // one per CTA, fed by the multicast commit
  uint64_t* chunk_free_bar = tmem_full_bar + 1;        // leader: 8 epilogue warps (4 per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(chunk_free_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tile = blockIdx.x >> 1;
  const int tiles_m = (int)((K + BM - 1) / BM);
  const int band = tile / (RASTER_M * tiles_n);
  const int within = tile % (RASTER_M * tiles_n);
  const int band_rows = min(RASTER_M, tiles_m - band * RASTER_M);
  const int n_blk = within / band_rows;
  const int m_blk = band * RASTER_M + within % band_rows;
  const int z = blockIdx.y;
  // tiles entirely below the diagonal do not run (both CTAs of the pair take the same decision)
  if (n_blk < m_blk) return;
  const int64_t t0 = (int64_t)z * tokens_per_split;
  const int64_t t1 = min(T, t0 + tokens_per_split);
  const int num_kb = (int)((t1 - t0 + BKT - 1) / BKT);
  const int chunk_kb = PER_SAMPLE ? kb_per_sample : max(num_kb, 1);
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  const int sample0 = PER_SAMPLE ? (int)(t0 / ((int64_t)kb_per_sample * BKT)) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(chunk_free_bar, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  cluster_sync_all();              // both CTAs' barriers and TMEM are ready before anyone signals
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs; completion is signalled on the LEADER's full barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
        const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
        uint8_t* a_dst = smem + stage * STAGE_BYTES;
        uint8_t* b_dst = a_dst + A_BYTES;
        const int32_t tok = (int32_t)(t0 + (int64_t)kb * BKT);
        const int32_t a_ch = m_blk * BM + (int)rank * 128;
        const int32_t b_ch = n_blk * BN + (int)rank * 128;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_2d_pair(a_dst + j * BOX_BYTES, &tmap, full_leader, a_ch + j * BOX_CH, tok);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          tma_load_2d_pair(b_dst + j * BOX_BYTES, &tmap, full_leader, b_ch + j * BOX_CH, tok);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader's one thread drives both SMs' tensor cores =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, BF16, /*a_mn=*/true, /*b_mn=*/true);
      int stage = 0;
      uint32_t phase = 0;
      int kb = 0;
      for (int ch = 0; ch < num_chunks; ++ch) {
        if (PER_SAMPLE && ch > 0) {                  // both CTAs have folded the previous chunk in
          mbar_wait(chunk_free_bar, (uint32_t)((ch - 1) & 1));
          tc_fence_after_sync();
        }
        const int kb_end = min(num_kb, kb + chunk_kb);
        for (bool first = true; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BKT / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 128;
            const uint64_t da = make_smem_desc_sw128(a_addr + koff, BOX_BYTES, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + koff, BOX_BYTES, 1024);
            mma_f16_ss_pair(tmem_base, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          mma_commit_pair(&empty_bar[stage]);     // frees the stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit_pair(tmem_full_bar);           // (chunk) accumulator complete, both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: this CTA's 128 rows x 256 columns =====
    const int q = warp & 3;
    float* dst_base = partial + (int64_t)z * K * K;
    const int64_t row0 = (int64_t)m_blk * BM + (int64_t)rank * 128;
    const int64_t row = row0 + q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t free_leader = mapa_u32(smem_u32(chunk_free_bar), 0);
    auto sample_weight = [&](int ch) {
      if (norms == nullptr) return 1.f;
      const float a = 1.f / (norms[sample0 + ch] + 1e-5f);     // gptq_quantizer.py:143
      return a * a;
    };
    if (PER_SAMPLE) {
      for (int ch = 0; ch + 1 < num_chunks; ++ch) {
        mbar_wait(tmem_full_bar, (uint32_t)(ch & 1));
        tc_fence_after_sync();
        const float wgt = sample_weight(ch);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32], t[32];
          tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
          if (ch > 0) {
            tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(wgt, __uint_as_float(v[j]), __uint_as_float(t[j])));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(wgt * __uint_as_float(v[j]));
          }
          tmem_st_32x32(lane_base + (uint32_t)(BN + c * 32), v);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(free_leader);
      }
    }
    if (num_chunks > 0) {
      mbar_wait(tmem_full_bar, (uint32_t)((num_chunks - 1) & 1));
      tc_fence_after_sync();
    }
    const float last_wgt = (PER_SAMPLE && num_chunks > 0) ? sample_weight(num_chunks - 1) : 1.f;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (num_chunks > 0) {
        tmem_ld_32x32(lane_base + (uint32_t)(c * 32), v);
        if (PER_SAMPLE) {
          if (num_chunks > 1) {
            uint32_t t[32];
            tmem_ld_32x32(lane_base + (uint32_t)(BN + c * 32), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = __float_as_uint(fmaf(last_wgt, __uint_as_float(v[j]), __uint_as_float(t[j])));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(last_wgt * __uint_as_float(v[j]));
          }
        } else {
          tmem_ld_wait();
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int64_t col0 = (int64_t)n_blk * BN + c * 32;
      const int64_t warp_row0 = row0 + q * 32;
      if (col0 + 32 <= warp_row0) continue;       // nothing of the upper triangle in this chunk
      if (row < K) {
        float* dst = dst_base + row * K + col0;
        if (col0 >= row && col0 + 32 <= K) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j)
            if (col0 + j >= row && col0 + j < K) dst[j] = __uint_as_float(v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t col = col0 + j;
        if (col > row && col < K && row < K) dst_base[col * K + row] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before_sync();
  cluster_sync_all();              // the peer may still be reading its TMEM / using our barriers
  if (warp == 2) tmem_dealloc_pair<kTmemCols>(tmem_base);
}

// =================================================================================================
// 4. split reduction and finalisation
// =================================================================================================
__global__ void hessian_reduce_kernel(const float* __restrict__ partial, int splits, int64_t KK,
                                      const float* __restrict__ unscale_ptr,
                                      float* __restrict__ H, int accumulate) {
  const float unscale = unscale_ptr ? *unscale_ptr : 1.f;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < KK;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 p = *reinterpret_cast<const float4*>(partial + (int64_t)z * KK + i);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    acc.x *= unscale; acc.y *= unscale; acc.z *= unscale; acc.w *= unscale;
    if (accumulate) {
      const float4 h = *reinterpret_cast<const float4*>(H + i);
      acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
    }
    *reinterpret_cast<float4*>(H + i) = acc;
  }
}

// H = H * scale + damp * I      (gptq_quantizer.py:150 and the 1e-6 ridge of :160)
__global__ void hessian_finalize_kernel(float* __restrict__ H, int64_t K, float scale, float damp) {
  const int64_t KK = K * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < KK;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = H[i] * scale;
    if (i / K == i % K) v += damp;
    H[i] = v;
  }
}

// =================================================================================================
// host side
// =================================================================================================
static EncodeTiledFn get_encode_fn() {
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

// tensor map over a row-major [rows, cols] 16-bit matrix, box = box_rows x 64 columns, 128B swizzle
static int make_tmap_2d_16bit(CUtensorMap* map, const void* base, int64_t rows, int64_t cols,
                              int box_rows, int box_cols, bool bf16 = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled failed");
  return B200Q_OK;
}

// How many token splits per output tile.  One CTA per (tile, split), one CTA per SM at a time, so
// the launch takes ceil(tiles * s / 148) rounds of (k-blocks per split) MMA stages plus an
// epilogue that writes the fp32 tile and its mirror image to HBM; s > 1 adds the pass that sums
// the partial matrices.  `straight` = the single-split result can go to H directly (no such pass).
// Constants are measured B200 figures (0.44 us per 128x256x64 stage under the sustained power
// limit, ~6.5 us for 148 concurrent tile epilogues, 6.2 TB/s for the reduction).
// CTA pairs (cta_group::2, hessian_gemm2_kernel) unless B200Q_HESSIAN_PAIR=0 (read once)
static bool hessian_pair() {
  static const bool on = []() {
    const char* e = std::getenv("B200Q_HESSIAN_PAIR");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

static int hessian_splits(int64_t K, int64_t T, bool straight) {
  static const int forced = []() {
    const char* e = std::getenv("B200Q_HESSIAN_SPLITS");     // experiments only
    const int v = e != nullptr ? std::atoi(e) : 0;
    return (v >= 1 && v <= 16) ? v : 0;
  }();
  if (forced > 0) return forced;
  const bool pair = hessian_pair();
  const int64_t bm = pair ? hg2::BM : hg::BM;
  const int64_t tm = (K + bm - 1) / bm, tn = (K + hg::BN - 1) / hg::BN;
  int64_t tiles = 0;                       // only tiles that touch the upper triangle run
  for (int64_t m = 0; m < tm; ++m)
    for (int64_t n = 0; n < tn; ++n)
      if ((n + 1) * hg::BN > m * bm) ++tiles;
  if (pair) tiles *= 2;                    // two SMs per tile
  const int64_t kblocks = (T + hg::BKT - 1) / hg::BKT;
  const int64_t max_s = std::max<int64_t>(1, std::min<int64_t>(16, kblocks / 8));
  int64_t best = 1;
  double best_us = 1e30;
  for (int64_t s = 1; s <= max_s; ++s) {
    const double rounds = (double)((tiles * s + kNumSMs - 1) / kNumSMs);
    const double kb = (double)((kblocks + s - 1) / s);
    double us = rounds * (kb * 0.44 + 6.5);
    if (s > 1 || !straight) us += (double)(s + 1) * (double)K * (double)K * 4.0 / 6.2e6;
    if (us < best_us * 0.995) { best_us = us; best = s; }
  }
  return (int)best;
}

struct HessianWork {
  float* stats;       // [2n + 1]
  float* partial_st;  // [n * chunks * 2]
  __half* Xs;         // [T, K]
  float* partial;     // [splits, K, K]
  int chunks, splits;
  int64_t bytes;
};

static HessianWork hessian_layout(void* work, int64_t T, int64_t K, int n_samples,
                                  int straight /* 0, 1, or -1 = size for either */) {
  HessianWork w;
  // CTAs of the sample-statistics pass: chunks x n_samples, ~8 per SM to keep HBM busy
  w.chunks = (int)std::max<int64_t>(1, std::min<int64_t>(64, (8 * kNumSMs + n_samples - 1) / std::max(1, n_samples)));
  w.splits = straight < 0 ? std::max(hessian_splits(K, T, false), hessian_splits(K, T, true))
                          : hessian_splits(K, T, straight != 0);
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t off = 0;
  uint8_t* base = static_cast<uint8_t*>(work);
  w.stats = reinterpret_cast<float*>(base + off);
  off += align(sizeof(float) * (2 * (int64_t)n_samples + 1));
  w.partial_st = reinterpret_cast<float*>(base + off);
  off += align(sizeof(float) * 2 * (int64_t)n_samples * w.chunks);
  w.Xs = reinterpret_cast<__half*>(base + off);
  off += align(2 * T * K);
  w.partial = reinterpret_cast<float*>(base + off);
  off += align(sizeof(float) * (int64_t)w.splits * K * K);
  w.bytes = off;
  return w;
}

// =================================================================================================
// AWQ scale search, stage 2: loss_c = sum_rows dW_c H dW_c^T = <dW_c H, dW_c>
// One tcgen05 GEMM over all candidates stacked along M:  P = D * Hb^T  (D [n_cand * rows_pad, K]
// bf16; Hb [K, K] bf16 is H folded onto its lower triangle, Hb[n, k] being the K-major B operand
// -- the quadratic form is unchanged and half of the k-blocks drop out), with the
// dot product <P, D> fused into the epilogue: the 128x256 fp32 tile leaves TMEM, is multiplied
// with the matching bf16 tile of D and reduced to ONE float per CTA.  Per-candidate sums are
// formed afterwards in a fixed order (deterministic).
// Both operands are K-major: a TMA box is 64 k-elements (128 bytes) x 128 / 256 rows, i.e. rows of
// one 128-byte swizzle row each; UMMA descriptors advance 32 bytes per K = 16 step inside the row.
// =================================================================================================
namespace ag {
constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;     // 16 KiB
constexpr int B_BYTES = BN * BK * 2;     // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;     // two 128 x 256 fp32 accumulators
constexpr int RASTER_M = 16;
}  // namespace ag

// PERSISTENT: one CTA per SM walks the tile list (stride gridDim.x) with TWO accumulators in TMEM,
// so the epilogue of tile i (TMEM -> registers, dot with dW from global memory) runs while the MMAs
// of tile i+1 fill the other accumulator; with the triangular k-ranges the average main loop is only
// ~34 k-blocks and an exposed epilogue cost 15-20 % of the kernel.
__global__ void __launch_bounds__(ag::THREADS, 1)
awq_loss_gemm_kernel(const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_h,
                     const __nv_bfloat16* __restrict__ D, float* __restrict__ tile_loss, int64_t Mtot,
                     int64_t K, int tiles_n) {
  using namespace ag;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full_bar = empty_bar + STAGES;      // [2] MMA -> epilogue
  uint64_t* acc_empty_bar = acc_full_bar + 2;       // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (int)((Mtot + BM - 1) / BM);
  const int total_tiles = tiles_m * tiles_n;
  // Hb is LOWER triangular (see h_to_lower_bf16_kernel): output columns [n0, n0 + BN) only receive
  // contributions from k < n0 + BN, so the k-loop of a tile stops there -- half the MMAs of the
  // dense product overall.  Tile order: bands of RASTER_M tile rows (their dW rows stay in L2),
  // heaviest (rightmost) tile column first inside a band.
  auto tile_coords = [&](int t, int& m_blk, int& n_blk, int& num_kb) {
    const int band = t / (RASTER_M * tiles_n);
    const int within = t % (RASTER_M * tiles_n);
    const int band_rows = min(RASTER_M, tiles_m - band * RASTER_M);
    n_blk = tiles_n - 1 - within / band_rows;
    m_blk = band * RASTER_M + within % band_rows;
    num_kb = (int)((min(K, (int64_t)(n_blk + 1) * BN) + BK - 1) / BK);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_d);
    tma_prefetch_desc(&tmap_h);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full_bar[s], 1); mbar_init(&acc_empty_bar[s], 4); }
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
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m_blk, n_blk, num_kb;
        tile_coords(t, m_blk, n_blk, num_kb);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(a_dst, &tmap_d, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(a_dst + A_BYTES, &tmap_h, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, /*bf16=*/true, /*a_mn=*/false, /*b_mn=*/false);
      int stage = 0; uint32_t phase = 0;
      int i = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
        int m_blk, n_blk, num_kb;
        tile_coords(t, m_blk, n_blk, num_kb);
        const int buf = i & 1;
        // the epilogue of the tile that used this accumulator two tiles ago must have drained it
        mbar_wait(&acc_empty_bar[buf], (uint32_t)(((i >> 1) & 1) ^ 1));
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * UMMA_K * 2, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * UMMA_K * 2, 16, 1024);
            mma_f16_ss(acc, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        mma_commit(&acc_full_bar[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      int m_blk, n_blk, num_kb;
      tile_coords(t, m_blk, n_blk, num_kb);
      const int buf = i & 1;
      const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
      mbar_wait(&acc_full_bar[buf], (uint32_t)((i >> 1) & 1));
      tc_fence_after_sync();
      const uint32_t acc_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      float acc = 0.f;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(acc_addr + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        const int64_t col0 = (int64_t)n_blk * BN + c * 32;
        if (row < Mtot && col0 < K) {
          const __nv_bfloat16* dp = D + row * K + col0;
          if (col0 + 32 <= K) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 raw = *reinterpret_cast<const uint4*>(dp + j);
              float d[8];
              unpack16<__nv_bfloat16>(raw, d);
#pragma unroll
              for (int u = 0; u < 8; ++u) acc = fmaf(__uint_as_float(v[j + u]), d[u], acc);
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < K; ++j)
              acc = fmaf(__uint_as_float(v[j]), __bfloat162float(dp[j]), acc);
          }
        }
      }
      // the accumulator has been read: hand it back before finishing the reduction
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[buf]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) tile_loss[((int64_t)m_blk * tiles_n + n_blk) * 4 + q] = acc;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// Hb[n][k] = bf16(H[n][k] + H[k][n]) for k < n, bf16(H[n][n]) on the diagonal, 0 above it.
// x H x^T = sum_n x_n (H_nn x_n + sum_{k<n} (H_nk + H_kn) x_k) for ANY square H, so the search GEMM
// can use this triangular operand and skip the k-blocks right of each output tile.
__global__ void __launch_bounds__(256)
h_to_lower_bf16_kernel(const float* __restrict__ H, __nv_bfloat16* __restrict__ Hb, int64_t K) {
  constexpr int TS = 64;                                        // 64 x 64 block per CTA
  __shared__ float tr[TS][TS + 1];
  const int64_t n0 = (int64_t)blockIdx.y * TS, k0 = (int64_t)blockIdx.x * TS;
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;  // 16 float4 per row, 16 rows / pass
  const bool vec = (K % 4 == 0);
  if (k0 >= n0 + TS) {                                          // entirely above the diagonal
    for (int r = r0; r < TS; r += 16) {
      const int64_t n = n0 + r;
      if (n >= K) break;
      for (int j = 0; j < 4; ++j)
        if (k0 + c4 + j < K) Hb[n * K + k0 + c4 + j] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  // block H[k0.., n0..] into shared memory (coalesced 16-byte loads), read back transposed
  for (int r = r0; r < TS; r += 16) {
    const int64_t kr = k0 + r, nc = n0 + c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kr < K) {
      if (vec && nc + 3 < K) v = *reinterpret_cast<const float4*>(H + kr * K + nc);
      else {
        if (nc < K) v.x = H[kr * K + nc];
        if (nc + 1 < K) v.y = H[kr * K + nc + 1];
        if (nc + 2 < K) v.z = H[kr * K + nc + 2];
        if (nc + 3 < K) v.w = H[kr * K + nc + 3];
      }
    }
    tr[r][c4] = v.x; tr[r][c4 + 1] = v.y; tr[r][c4 + 2] = v.z; tr[r][c4 + 3] = v.w;
  }
  __syncthreads();
  for (int r = r0; r < TS; r += 16) {
    const int64_t n = n0 + r;
    if (n >= K) break;
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t kc = k0 + c4;
    if (vec && kc + 3 < K) {
      const float4 v = *reinterpret_cast<const float4*>(H + n * K + kc);
      h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
    } else {
      for (int j = 0; j < 4; ++j) if (kc + j < K) h[j] = H[n * K + kc + j];
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = kc + j;
      o[j] = (k < n) ? h[j] + tr[c4 + j][r] : (k == n ? h[j] : 0.f);
    }
    if (vec && kc + 3 < K) {
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&p0);
      pk.y = *reinterpret_cast<const uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(Hb + n * K + kc) = pk;
    } else {
      for (int j = 0; j < 4; ++j) if (kc + j < K) Hb[n * K + kc + j] = __float2bfloat16_rn(o[j]);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Symmetric matrices between GPUs: the packed lower triangle.  X^T X partials are exactly symmetric
// (the SYRK epilogue mirrors every tile), so ranks exchange row n as its columns 0..n, stored at
// offset n(n+1)/2 -- half the bytes of the square.  For the AWQ search the folded operand
// Hb[n][k] = H[n][k] + H[k][n] is then simply 2 P[n,k] off the diagonal: the fold runs on each
// rank's reduce-scattered slice and the slices are all-gathered as bf16.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void packed_row_col(int64_t e, int64_t& n, int64_t& k) {
  n = (int64_t)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
  while (n * (n + 1) / 2 > e) --n;
  while ((n + 1) * (n + 2) / 2 <= e) ++n;
  k = e - n * (n + 1) / 2;
}

// grid (ceil(K/256), K): row n, columns k <= n
__global__ void __launch_bounds__(256)
sym_pack_lower_kernel(const float* __restrict__ H, int64_t K, float* __restrict__ P) {
  const int64_t n = blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k <= n) P[n * (n + 1) / 2 + k] = H[n * K + k];
}

// 32 x 32 tiles on and below the diagonal; off-diagonal tiles also write their mirror image
__global__ void __launch_bounds__(256)
sym_unpack_lower_kernel(const float* __restrict__ P, int64_t K, float* __restrict__ H) {
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t n0 = bi * 32, k0 = bj * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t n = n0 + i, k = k0 + tx;
    float v = 0.f;
    if (n < K && k < K) {
      const int64_t hi = max(n, k), lo = min(n, k);
      v = P[hi * (hi + 1) / 2 + lo];
      H[n * K + k] = v;
    }
    tile[i][tx] = v;
  }
  if (bj == bi) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t k = k0 + i, n = n0 + tx;          // H[k][n] = tile[n - n0][k - k0]
    if (n < K && k < K) H[k * K + n] = tile[tx][i];
  }
}

// packed elements [e0, e1) -> bf16(2 P) off the diagonal, bf16(P) on it
__global__ void __launch_bounds__(256)
sym_fold_packed_kernel(const float* __restrict__ P, int64_t e0, int64_t e1, int64_t len,
                       __nv_bfloat16* __restrict__ out) {
  for (int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e1;
       e += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (e < len) {
      int64_t n, k;
      packed_row_col(e, n, k);
      v = P[e - e0];
      if (k != n) v *= 2.f;
    }
    out[e - e0] = __float2bfloat16_rn(v);
  }
}

// packed folded bf16 triangle -> [K,K] bf16, zeros above the diagonal; grid (ceil(K/256), K)
__global__ void __launch_bounds__(256)
sym_unpack_folded_kernel(const __nv_bfloat16* __restrict__ Pb, int64_t K,
                         __nv_bfloat16* __restrict__ Hb) {
  const int64_t n = blockIdx.y;
  const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (k < K) Hb[n * K + k] = (k <= n) ? Pb[n * (n + 1) / 2 + k] : __float2bfloat16_rn(0.f);
}

// loss[c] (+)= sum over the tile rows of candidate c, fixed order, fp64 accumulate
__global__ void awq_loss_reduce_kernel(const float* __restrict__ tile_loss, int tiles_per_cand,
                                       int n_cand, float* __restrict__ loss) {
  const int c = blockIdx.x;
  if (c >= n_cand) return;
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < tiles_per_cand; i += blockDim.x)
    s += (double)tile_loss[(int64_t)c * tiles_per_cand + i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[c] += (float)sm[0];
}

static int make_tmap_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows,
                          int box_cols) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200Q_ECUDA, "cuTensorMapEncodeTiled (bf16) failed");
  return B200Q_OK;
}

struct AwqWork {
  __nv_bfloat16* D;
  __nv_bfloat16* Hb;
  float* tile_loss;
  int64_t rows_pad, bytes;
  int tiles_n, tiles_per_cand;
};

static AwqWork awq_layout(void* work, int64_t N, int64_t K, int n_cand) {
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  AwqWork w;
  w.rows_pad = (N + ag::BM - 1) / ag::BM * ag::BM;
  w.tiles_n = (int)((K + ag::BN - 1) / ag::BN);
  w.tiles_per_cand = (int)(w.rows_pad / ag::BM) * w.tiles_n;
  uint8_t* base = static_cast<uint8_t*>(work);
  int64_t off = 0;
  w.D = reinterpret_cast<__nv_bfloat16*>(base + off); off += align(2 * (int64_t)n_cand * w.rows_pad * K);
  w.Hb = reinterpret_cast<__nv_bfloat16*>(base + off); off += align(2 * K * K);
  w.tile_loss = reinterpret_cast<float*>(base + off); off += align(16 * (int64_t)n_cand * w.tiles_per_cand);
  w.bytes = off;
  return w;
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int64_t b200q_hessian_workspace(int64_t T, int64_t K, int n_samples) {
  if (T <= 0 || K <= 0 || n_samples <= 0) return 0;
  return hessian_layout(nullptr, T, K, n_samples, -1).bytes;
}

int b200q_hessian_accum(const void* X, int n_samples, int64_t rows_per_sample, int64_t K, int dtype,
                        int normalize, float* H, int accumulate, float* norms_out, void* work,
                        void* stream) {
  B200Q_REQUIRE(X && H && work, "hessian_accum: null pointer");
  B200Q_REQUIRE(n_samples > 0 && rows_per_sample > 0 && K > 0, "hessian_accum: bad shape");
  B200Q_REQUIRE(K % 8 == 0, "hessian_accum: in_features must be a multiple of 8");
  B200Q_REQUIRE(aligned16(X) && aligned16(H) && aligned16(work), "hessian_accum: unaligned pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t T = (int64_t)n_samples * rows_per_sample;
  B200Q_REQUIRE(T < (1ll << 31) && K < (1ll << 31), "hessian_accum: dimension too large");
  const double flops = 2.0 * (double)T * (double)K * (double)K;

  // 16-bit activations need no staging pass: fp16 / bf16 products are exact in the fp32
  // accumulator, so TMA reads the caller's tensor directly.
  //   direct      : plain Gram matrix (no normalisation, no norms wanted) -- no other pass at all
  //   per_sample  : GPTQ normalisation with samples of >= 512 rows in whole 64-token blocks -- one
  //                 read-only pass for the sample norms, then the kernel folds each sample in with
  //                 its weight (PER_SAMPLE above)
  // Everything else (fp32 input, short or ragged samples) goes through stats + prescale into fp16.
  const bool sixteen = dtype == B200Q_F16 || dtype == B200Q_BF16;
  const bool direct = sixteen && !normalize && norms_out == nullptr;
  const bool per_sample = sixteen && normalize && rows_per_sample % hg::BKT == 0 &&
                          rows_per_sample >= 8 * hg::BKT;
  const bool in_place = direct || per_sample;
  const bool bf16_ops = in_place && dtype == B200Q_BF16;
  HessianWork w = hessian_layout(work, T, K, n_samples, (direct && !accumulate) ? 1 : 0);
  if (!direct) {
    KernelScope scope("hessian_prescale", (per_sample ? 1.0 : 3.0) * T * K * elem_size(dtype), 0, st);
    B200Q_DISPATCH_DTYPE(dtype, Tt, {
      constexpr int VEC = ST<Tt>::VEC;
      B200Q_REQUIRE((rows_per_sample * K) % VEC == 0, "hessian_accum: sample not 16-byte sized");
      dim3 g1((unsigned)w.chunks, (unsigned)n_samples);
      sample_stats_partial_kernel<Tt><<<g1, 256, 0, st>>>(static_cast<const Tt*>(X),
                                                          rows_per_sample, K, w.partial_st);
      sample_stats_finish_kernel<<<1, 256, 0, st>>>(w.partial_st, w.chunks, n_samples, w.stats,
                                                    normalize);
      count_launch(2);
      if (!per_sample) {
        const int64_t total_vec = T * K / VEC;
        const int blocks = (int)std::min<int64_t>((total_vec + 255) / 256, (int64_t)kNumSMs * 16);
        prescale_kernel<Tt><<<blocks, 256, 0, st>>>(static_cast<const Tt*>(X), w.Xs, rows_per_sample,
                                                    K, total_vec, w.stats);
        count_launch();
      }
    });
    int rc = check_launch("hessian_accum/prescale");
    if (rc != B200Q_OK) return rc;
  }
  if (norms_out != nullptr)
    cudaMemcpyAsync(norms_out, w.stats + n_samples, sizeof(float) * n_samples,
                    cudaMemcpyDeviceToDevice, st);

  CUtensorMap tmap;
  int rc = make_tmap_2d_16bit(&tmap, in_place ? X : static_cast<const void*>(w.Xs), T, K, hg::BKT,
                              hg::BOX_CH, bf16_ops);
  if (rc != B200Q_OK) return rc;
  // (per-device attribute, set once per device and process)
  {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      bool ok = true;
#define B200Q_HG_ATTR(KERNEL, BYTES) \
      ok = ok && cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES) == cudaSuccess
      B200Q_HG_ATTR((hessian_gemm_kernel<false, false>), hg::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm_kernel<true, false>), hg::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm_kernel<false, true>), hg::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm_kernel<true, true>), hg::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm2_kernel<false, false>), hg2::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm2_kernel<true, false>), hg2::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm2_kernel<false, true>), hg2::SMEM_BYTES);
      B200Q_HG_ATTR((hessian_gemm2_kernel<true, true>), hg2::SMEM_BYTES);
#undef B200Q_HG_ATTR
      if (!ok) return fail(B200Q_ECUDA, "hessian_gemm: cannot raise shared memory");
      done[dev] = true;
    }
  }
  const int64_t kblocks = (T + hg::BKT - 1) / hg::BKT;
  int splits = w.splits;
  int64_t tokens_per_split;
  // chunk of the k loop accumulated inside the tensor core: one sample; 2048 tokens for the staged
  // Hessian; 4096 tokens for the plain Gram matrix (diagonal low by <= 2.4e-4, an eighth of the
  // bf16 rounding its consumer, the AWQ search, applies -- measured on one box: 2048-token chunks
  // cost 5.5 % of this kernel, no chunking leaves the diagonal 0.2 % low)
  const int kb_per_sample = per_sample ? (int)(rows_per_sample / hg::BKT) : (direct ? 64 : 32);
  if (per_sample) {
    // splits cover whole samples
    const int64_t samples_per_split = ((int64_t)n_samples + splits - 1) / splits;
    splits = (int)(((int64_t)n_samples + samples_per_split - 1) / samples_per_split);
    tokens_per_split = samples_per_split * rows_per_sample;
  } else {
    const int64_t kb_per_split = (kblocks + splits - 1) / splits;
    tokens_per_split = kb_per_split * hg::BKT;
  }
  // one split, nothing to add to and no scale to undo: the tiles go straight into H
  const bool straight = in_place && splits == 1 && !accumulate;
  float* gemm_out = straight ? H : w.partial;
  {
    KernelScope scope("hessian_gemm", 0, flops, st);
    const bool pair = hessian_pair();
    // band height of the pair kernel's tile order, in 256-row tile rows (B200Q_HESSIAN_RASTER
    // overrides, read per call).  Interleaved A/B at 262144 tokens: K = 11008 runs 27.9 ms with
    // bands of 4, 29.1 with 8, 32.2 with 16; K = 4096 (16 tile rows) is flat within 3 %.
    const int raster_m = [&]() {
      const char* e = std::getenv("B200Q_HESSIAN_RASTER");
      const int v = e != nullptr ? std::atoi(e) : 0;
      if (v >= 1 && v <= 64) return v;
      return (K > 6144) ? 4 : hg2::RASTER_M_DEFAULT;
    }();
    const int tiles_n = (int)((K + hg::BN - 1) / hg::BN);
    const int tiles_m = (int)((K + (pair ? hg2::BM : hg::BM) - 1) / (pair ? hg2::BM : hg::BM));
    // pairs: two CTAs (one cluster) per 256 x 256 tile
    dim3 grid((unsigned)(tiles_m * tiles_n * (pair ? 2 : 1)), (unsigned)splits);
    const float* norms = per_sample ? w.stats + n_samples : nullptr;
#define B200Q_HG_LAUNCH(BF, PS)                                                                   \
    do {                                                                                          \
      if (pair)                                                                                   \
        hessian_gemm2_kernel<BF, PS><<<grid, hg2::THREADS, hg2::SMEM_BYTES, st>>>(                 \
            tmap, gemm_out, K, T, tokens_per_split, tiles_n, norms, kb_per_sample, raster_m);     \
      else                                                                                        \
        hessian_gemm_kernel<BF, PS><<<grid, hg::THREADS, hg::SMEM_BYTES, st>>>(                    \
            tmap, gemm_out, K, T, tokens_per_split, tiles_n, norms, kb_per_sample);               \
    } while (0)
    // (B200Q_HESSIAN_CHUNKED=0 keeps one long accumulation for A/B timing; per-sample always chunks)
    static const bool chunk_env = []() {
      const char* e = std::getenv("B200Q_HESSIAN_CHUNKED");
      return !(e != nullptr && e[0] == '0');
    }();
    const bool chunked = per_sample || (chunk_env && kblocks / splits > kb_per_sample);   // long k loops only
    if (chunked) {
      if (bf16_ops) B200Q_HG_LAUNCH(true, true); else B200Q_HG_LAUNCH(false, true);
    } else {
      if (bf16_ops) B200Q_HG_LAUNCH(true, false); else B200Q_HG_LAUNCH(false, false);
    }
#undef B200Q_HG_LAUNCH
    count_launch();
    rc = check_launch("hessian_gemm");
    if (rc != B200Q_OK) return rc;
  }
  if (!straight) {
    KernelScope scope("hessian_reduce", sizeof(float) * (double)(splits + 1) * K * K, 0, st);
    const int64_t KK = K * K;
    const int blocks = (int)std::min<int64_t>((KK / 4 + 255) / 256, (int64_t)kNumSMs * 16);
    hessian_reduce_kernel<<<blocks, 256, 0, st>>>(w.partial, splits, KK,
                                                  in_place ? nullptr : w.stats + 2 * n_samples, H,
                                                  accumulate);
    count_launch();
    rc = check_launch("hessian_reduce");
  }
  return rc;
}

int b200q_hessian_finalize(float* H, int64_t K, float scale, float damp, void* stream) {
  B200Q_REQUIRE(H && K > 0, "hessian_finalize: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t KK = K * K;
  const int blocks = (int)std::min<int64_t>((KK + 255) / 256, (int64_t)kNumSMs * 16);
  hessian_finalize_kernel<<<blocks, 256, 0, st>>>(H, K, scale, damp);
  count_launch();
  return check_launch("hessian_finalize");
}


int64_t b200q_awq_search_workspace(int64_t N, int64_t K, int n_cand) {
  if (N <= 0 || K <= 0 || n_cand <= 0) return 0;
  return awq_layout(nullptr, N, K, n_cand).bytes;
}

// stage 1: dW_c = Q_c(W) - W for every candidate, bf16, into the workspace
static int awq_search_delta_stage(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                                  const uint8_t* salient, const float* sf_host, int n_cand, int dtype,
                                  void* work, void* stream) {
  B200Q_REQUIRE(W && salient && sf_host && work, "awq_search: null pointer");
  B200Q_REQUIRE(N > 0 && K > 0, "awq_search: bad shape");
  // the bf16 operands are read by TMA: 16-byte row pitch
  B200Q_REQUIRE(K % 8 == 0, "awq_search: in_features must be a multiple of 8");
  B200Q_REQUIRE(aligned16(work), "awq_search: unaligned workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AwqWork w = awq_layout(work, N, K, n_cand);
  const int64_t Mtot = (int64_t)n_cand * w.rows_pad;
  KernelScope scope("awq_search_delta", (double)N * K * (elem_size(dtype) + 2.0 * n_cand), 0, st);
  if (w.rows_pad != N)   // padding rows must be zero: they are read by the GEMM
    cudaMemsetAsync(w.D, 0, 2 * Mtot * K, st);
  return launch_awq_delta(W, w.D, salient, N, K, group, w.rows_pad * K, n_bit, sf_host, n_cand, dtype, st);
}

// stage 2: fold H (unless the caller supplies the folded operand), the loss GEMM, the reduction
static int awq_search_loss_stage(int64_t N, int64_t K, int n_cand, const float* H,
                                 const __nv_bfloat16* Hb_folded, void* work, float* loss,
                                 void* stream) {
  B200Q_REQUIRE((H || Hb_folded) && work && loss, "awq_search_loss: null pointer");
  B200Q_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && n_cand >= 1, "awq_search_loss: bad shape");
  B200Q_REQUIRE(aligned16(work), "awq_search_loss: unaligned workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AwqWork w = awq_layout(work, N, K, n_cand);
  const int64_t Mtot = (int64_t)n_cand * w.rows_pad;
  int rc;
  const __nv_bfloat16* Hb = Hb_folded;
  if (Hb == nullptr) {
    KernelScope scope("awq_search_fold", 7.0 * K * K, 0, st);   // lower half: 8 B in + 2 B out; upper: 2 B out
    const unsigned tb = (unsigned)((K + 63) / 64);
    h_to_lower_bf16_kernel<<<dim3(tb, tb), 256, 0, st>>>(H, w.Hb, K);
    count_launch();
    Hb = w.Hb;
  } else {
    B200Q_REQUIRE(aligned16(Hb), "awq_search_loss: unaligned folded operand");
  }
  CUtensorMap tmap_d, tmap_h;
  rc = make_tmap_bf16(&tmap_d, w.D, Mtot, K, ag::BM, ag::BK);
  if (rc != B200Q_OK) return rc;
  rc = make_tmap_bf16(&tmap_h, Hb, K, K, ag::BN, ag::BK);
  if (rc != B200Q_OK) return rc;
  cudaFuncSetAttribute(awq_loss_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       ag::SMEM_BYTES);
  {
    KernelScope scope("awq_search_gemm", 0, 2.0 * (double)n_cand * N * K * K, st);
    const int tiles_m = (int)(Mtot / ag::BM);
    const unsigned ctas = (unsigned)std::min<int64_t>((int64_t)tiles_m * w.tiles_n, kNumSMs);
    awq_loss_gemm_kernel<<<ctas, ag::THREADS, ag::SMEM_BYTES, st>>>(
        tmap_d, tmap_h, w.D, w.tile_loss, Mtot, K, w.tiles_n);
    count_launch();
    rc = check_launch("awq_loss_gemm");
    if (rc != B200Q_OK) return rc;
  }
  awq_loss_reduce_kernel<<<n_cand, 256, 0, st>>>(w.tile_loss, 4 * w.tiles_per_cand, n_cand, loss);
  count_launch();
  return check_launch("awq_loss_reduce");
}

int b200q_awq_search_loss(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                          const uint8_t* salient, const float* sf_host, int n_cand, const float* H,
                          int dtype, void* work, float* loss, void* stream) {
  B200Q_REQUIRE(H, "awq_search_loss: null pointer");
  int rc = awq_search_delta_stage(W, N, K, group, n_bit, salient, sf_host, n_cand, dtype, work, stream);
  if (rc != B200Q_OK) return rc;
  return awq_search_loss_stage(N, K, n_cand, H, nullptr, work, loss, stream);
}

int b200q_awq_search_loss_folded(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                                 const uint8_t* salient, const float* sf_host, int n_cand,
                                 const void* Hb_folded, int dtype, void* work, float* loss,
                                 void* stream) {
  B200Q_REQUIRE(Hb_folded, "awq_search_loss: null pointer");
  int rc = awq_search_delta_stage(W, N, K, group, n_bit, salient, sf_host, n_cand, dtype, work, stream);
  if (rc != B200Q_OK) return rc;
  return awq_search_loss_stage(N, K, n_cand, nullptr, static_cast<const __nv_bfloat16*>(Hb_folded), work,
                               loss, stream);
}

int b200q_awq_search_delta(const void* W, int64_t N, int64_t K, int64_t group, int n_bit,
                           const uint8_t* salient, const float* sf_host, int n_cand, int dtype,
                           void* work, void* stream) {
  return awq_search_delta_stage(W, N, K, group, n_bit, salient, sf_host, n_cand, dtype, work, stream);
}

int b200q_awq_search_loss_prepared(int64_t N, int64_t K, int n_cand, const float* H,
                                   const void* Hb_folded, void* work, float* loss, void* stream) {
  return awq_search_loss_stage(N, K, n_cand, H, static_cast<const __nv_bfloat16*>(Hb_folded), work, loss,
                               stream);
}

int64_t b200q_sym_packed_len(int64_t K) { return K > 0 ? K * (K + 1) / 2 : 0; }

int b200q_sym_pack_lower(const float* H, int64_t K, float* P, void* stream) {
  B200Q_REQUIRE(H && P && K > 0 && K < (1 << 30), "sym_pack_lower: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("sym_pack", 4.0 * K * (K + 1), 0, st);
  sym_pack_lower_kernel<<<dim3((unsigned)((K + 255) / 256), (unsigned)K), 256, 0, st>>>(H, K, P);
  count_launch();
  return check_launch("sym_pack_lower");
}

int b200q_sym_unpack_lower(const float* P, int64_t K, float* H, void* stream) {
  B200Q_REQUIRE(H && P && K > 0 && K < (1 << 30), "sym_unpack_lower: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("sym_unpack", 6.0 * K * K, 0, st);
  const unsigned tb = (unsigned)((K + 31) / 32);
  sym_unpack_lower_kernel<<<dim3(tb, tb), 256, 0, st>>>(P, K, H);
  count_launch();
  return check_launch("sym_unpack_lower");
}

int b200q_sym_fold_packed_bf16(const float* P_slice, int64_t K, int64_t e0, int64_t e1, void* out_bf16,
                               void* stream) {
  B200Q_REQUIRE(P_slice && out_bf16 && K > 0 && e0 >= 0 && e1 >= e0, "sym_fold_packed: bad argument");
  if (e1 == e0) return B200Q_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("sym_fold", 6.0 * (double)(e1 - e0), 0, st);
  const int blocks = (int)std::min<int64_t>((e1 - e0 + 255) / 256, (int64_t)kNumSMs * 16);
  sym_fold_packed_kernel<<<blocks, 256, 0, st>>>(P_slice, e0, e1, K * (K + 1) / 2,
                                                 static_cast<__nv_bfloat16*>(out_bf16));
  count_launch();
  return check_launch("sym_fold_packed");
}

int b200q_sym_unpack_folded_bf16(const void* Pb, int64_t K, void* Hb, void* stream) {
  B200Q_REQUIRE(Pb && Hb && K > 0 && K < (1 << 30), "sym_unpack_folded: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("sym_unpack_folded", 3.0 * K * K, 0, st);
  sym_unpack_folded_kernel<<<dim3((unsigned)((K + 255) / 256), (unsigned)K), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(Pb), K, static_cast<__nv_bfloat16*>(Hb));
  count_launch();
  return check_launch("sym_unpack_folded");
}

}  // extern "C"
