// Damped SPD inverse for GPTQ (ref: gptq_quantizer.py:160-165, torch.linalg.inv(H + 1e-6 I)).
//
// H is symmetric positive definite with cond(H) <= (1 + damp)/damp ~ 1e2 by construction (every
// normalised sample has trace 1), so a Cholesky route is safe and costs K^3
// flops against LU's 2 K^3:   H = L L^T  ->  L^-1  ->  H^-1 = L^-T L^-1.
// Blocked right-looking factorisation, NB = 128:
//   diag  : one CTA factors the 128x128 diagonal block in shared memory and inverts it
//   panel : L21 = A21 * L11^-T                      (GEMM with the inverted diagonal block)
//   trail : A22 -= L21 * L21^T  (lower tiles only)   (GEMM, the K^3/3 bulk)
// The triangular inverse and the final product are row-block sweeps of the same GEMM.
// All arithmetic is fp32 on the FP32 pipe (8x8 register tiles, 128x128x16 CTA tiles): the
// factorisation of a matrix that is later used to propagate quantisation errors needs fp32
// mantissas, and TF32 tensor-core inputs (10 bits) do not provide them.
//
// For the error-compensated GPTQ loop the quantity needed is U = chol(H^-1, upper).  With J the
// index reversal, J H J = Lr Lr^T gives H = R R^T with R = J Lr J upper triangular, hence
// H^-1 = R^-T R^-1 and U = R^-1 = J Lr^-1 J: the same factor-and-invert on the reversed matrix.
#include <algorithm>

#include "common.cuh"

namespace b200q {

namespace la {
constexpr int NB = 64;    // diagonal block: the one-CTA factor+inverse kernel is on the critical path K/NB times
constexpr int BM = 128, BN = 128, BK = 16;
}  // namespace la

// C[M,N] = alpha * op(A) * op(B) + beta * C, fp32.  op(A) is M x Kd: TA ? A stored [Kd, M] (lda)
// : A stored [M, Kd]; op(B) is Kd x N: TB ? B stored [N, Kd] (ldb) : B stored [Kd, N].
// tri: 0 = dense.  1 = C is needed on and below the diagonal only: skip CTA tiles strictly above
// it.  2 = op(A)^T.. product L^T L of a lower-triangular L (TA, !TB): A[k,m] = 0 for k < m and
// B[k,n] = 0 for k < n, so k starts at max(m0, n0).  3 = B lower triangular (B[k,n] = 0 for
// k < n): k starts at n0.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int Kd, float alpha, const float* __restrict__ A, int64_t lda,
             const float* __restrict__ B, int64_t ldb, float beta, float* __restrict__ C,
             int64_t ldc, int tri, int64_t batch_stride) {
  using namespace la;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (tri == 1 && n0 > m0 + BM - 1) return;
  // blockIdx.z = problem of a batch whose operands all advance by the same stride (the diagonal
  // block pairs of the divide-and-conquer triangular inverse)
  A += (int64_t)blockIdx.z * batch_stride;
  B += (int64_t)blockIdx.z * batch_stride;
  C += (int64_t)blockIdx.z * batch_stride;
  const int k_first = (tri == 2) ? (max(m0, n0) / BK) : (tri == 3 ? n0 / BK : 0);
  // tri == 4: A lower triangular (A[m,k] = 0 for k > m): k stops after this tile's last row
  if (tri == 4) Kd = min(Kd, m0 + BM);
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tiles = [&](int buf, int k0) {
    // A tile -> As[k][m]
    if constexpr (!TA) {
      // A[m, k]: 128 rows x 16 k; thread reads 8 consecutive k of one row
      const int r = tid >> 1, kc = (tid & 1) * 8;
      const int gm = m0 + r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gk = k0 + kc + j;
        As[buf][kc + j][r] = (gm < M && gk < Kd) ? A[(int64_t)gm * lda + gk] : 0.f;
      }
    } else {
      // A[k, m]: 16 k x 128 m; thread reads 8 consecutive m of one k
      const int kk = tid >> 4, mc = (tid & 15) * 8;
      const int gk = k0 + kk;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gm = m0 + mc + j;
        As[buf][kk][mc + j] = (gm < M && gk < Kd) ? A[(int64_t)gk * lda + gm] : 0.f;
      }
    }
    if constexpr (TB) {
      // B[n, k]
      const int r = tid >> 1, kc = (tid & 1) * 8;
      const int gn = n0 + r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gk = k0 + kc + j;
        Bs[buf][kc + j][r] = (gn < N && gk < Kd) ? B[(int64_t)gn * ldb + gk] : 0.f;
      }
    } else {
      // B[k, n]
      const int kk = tid >> 4, nc = (tid & 15) * 8;
      const int gk = k0 + kk;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int gn = n0 + nc + j;
        Bs[buf][kk][nc + j] = (gn < N && gk < Kd) ? B[(int64_t)gk * ldb + gn] : 0.f;
      }
    }
  };

  const int nk = (Kd + BK - 1) / BK;
  if (nk > k_first) load_tiles(k_first & 1, k_first * BK);
  __syncthreads();
  for (int t = k_first; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) load_tiles(buf ^ 1, (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty * 8 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + tx * 8 + j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      *c = (beta == 0.f) ? alpha * acc[i][j] : fmaf(alpha, acc[i][j], beta * *c);
    }
  }
}

template <bool TA, bool TB>
static void sgemm(cudaStream_t st, int M, int N, int Kd, float alpha, const float* A, int64_t lda,
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int tri = 0,
                  int batch = 1, int64_t batch_stride = 0) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  dim3 grid((N + la::BN - 1) / la::BN, (M + la::BM - 1) / la::BM, batch);
  sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, tri,
                                             batch_stride);
  count_launch();
}

// One CTA of 128 threads: Cholesky of the nb x nb (nb <= 128) diagonal block at A (lower, in place;
// the strict upper part of the block is zeroed) and its inverse into Linv (dense nb x nb copy,
// ld = NB) and Linv_big (ld = lda).  info = first non-positive pivot (1-based, offset by j0).
//
// Left-looking: in step k the threads of row r form A[r][k] - sum_{j<k} L[r][j] L[k][j] from
// their own row and the pivot row (a broadcast), and accumulate the pivot's own sum in the same
// loop, so every thread knows sqrt(pivot) without an exchange: two barriers per column and no
// rank-1 sweeps.  The inverse is a forward substitution per column.  This kernel sits on the
// critical path K/128 times, so its latency decides the factorisation time for K <= ~8K.
__global__ void __launch_bounds__(512)
potrf_inv_diag_kernel(float* __restrict__ A, int64_t lda, int nb, float* __restrict__ Linv,
                      float* __restrict__ Linv_big, int* __restrict__ info, int j0) {
  // FOUR threads per row (512 threads, 16 warps = 4 per scheduler): ncu on the one-thread-per-row
  // version showed 19 % issue utilisation, every instruction waiting ~5 cycles on the previous one
  // with a single warp per scheduler.  The 4 lanes of a row take j = q, q+4, ... and combine with
  // two shuffles.  LD = 132 (= 4 mod 32) makes (row, q) -> bank 4*row + q conflict free.
  extern __shared__ float sm[];
  constexpr int LD = la::NB + 4;
  float* L = sm;                 // [NB][LD]
  float* X = sm + la::NB * LD;   // [NB][LD]
  const int tid = threadIdx.x;
  for (int i = tid; i < la::NB * la::NB; i += blockDim.x) {
    const int r = i / la::NB, c = i % la::NB;
    L[r * LD + c] = (r < nb && c <= r) ? A[(int64_t)r * lda + c] : (r == c ? 1.f : 0.f);
    X[r * LD + c] = 0.f;
  }
  __syncthreads();
  const int r = tid >> 2, q = tid & 3;
  for (int k = 0; k < nb; ++k) {
    float t = 0.f, d = 0.f;
    const float* lr = L + r * LD;
    const float* lk = L + k * LD;
    // batches of 8 predicated loads first, FMAs after: the trip count is short (<= 32) and a
    // rolled load->FMA loop would pay the full shared-memory latency on every iteration
    float t1 = 0.f, d1 = 0.f;
    for (int base = q; base < k; base += 32) {
      float a[8], p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = base + 4 * u;
        const bool ok = j < k;
        a[u] = ok ? lr[j] : 0.f;
        p[u] = ok ? lk[j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        t = fmaf(-a[u], p[u], t);
        t1 = fmaf(-a[u + 1], p[u + 1], t1);
        d = fmaf(-p[u], p[u], d);
        d1 = fmaf(-p[u + 1], p[u + 1], d1);
      }
    }
    t += t1;
    d += d1;
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    t += lr[k];                            // A[r][k] - sum_j L[r][j] L[k][j]
    d += lk[k];                            // A[k][k] - sum_j L[k][j]^2
    const float piv = sqrtf(fmaxf(d, 1e-30f));
    if (tid == 0 && !(d > 0.f) && info != nullptr) atomicCAS(info, 0, j0 + k + 1);
    __syncthreads();                       // everyone has read column k of A / row k of L
    if (q == 0) {
      if (r == k) L[r * LD + k] = piv;
      else if (r > k) L[r * LD + k] = t / piv;
    }
    __syncthreads();
  }
  // inverse by forward substitution: four threads per column c solve L x = e_c
  const int c = tid >> 2;
  // the columns of one warp have different trip counts: synchronise the 4-lane team only
  const unsigned team = 0xFu << ((tid & 31) & ~3);
  if (c < nb) {
    for (int row = c; row < nb; ++row) {
      float s = 0.f;
      const float* lrow = L + row * LD;
      float s1 = 0.f;
      for (int base = c + q; base < row; base += 32) {
        float a[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = base + 4 * u;
          const bool ok = k < row;
          a[u] = ok ? lrow[k] : 0.f;
          x[u] = ok ? X[k * LD + c] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
          s = fmaf(-a[u], x[u], s);
          s1 = fmaf(-a[u + 1], x[u + 1], s1);
        }
      }
      s += s1;
      s += __shfl_xor_sync(team, s, 1);
      s += __shfl_xor_sync(team, s, 2);
      if (row == c) s += 1.f;
      if (q == 0) X[row * LD + c] = s / lrow[row];
      __syncwarp(team);
    }
  }
  __syncthreads();
  for (int i = tid; i < nb * nb; i += blockDim.x) {
    const int rr = i / nb, cc = i % nb;
    const float l = (cc <= rr) ? L[rr * LD + cc] : 0.f;
    const float x = (cc <= rr) ? X[rr * LD + cc] : 0.f;
    A[(int64_t)rr * lda + cc] = l;
    Linv[rr * la::NB + cc] = x;
    if (Linv_big != nullptr) Linv_big[(int64_t)rr * lda + cc] = x;
  }
}

// dst[i][j] = src[K-1-i][K-1-j]   (J * src * J)
__global__ void reverse_both_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                    int64_t K) {
  const int64_t n = K * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K, c = i % K;
    dst[i] = src[(K - 1 - r) * K + (K - 1 - c)];
  }
}

struct LinalgWork {
  float* A;      // [K,K] working copy -> L (lower)
  float* Linv;   // [K,K] -> L^-1 (lower)
  float* Dinv;   // [NB,NB] inverse of the current diagonal block
  float* T;      // [K,K] scratch for the block products of the triangular inverse
  int64_t bytes;
};

static LinalgWork linalg_layout(void* work, int64_t K) {
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  LinalgWork w;
  uint8_t* base = static_cast<uint8_t*>(work);
  int64_t off = 0;
  w.A = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.Linv = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.Dinv = reinterpret_cast<float*>(base + off); off += align(4 * la::NB * la::NB);
  w.T = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.bytes = off;
  return w;
}

// A (K x K, lower part valid) -> L in place (lower), Linv = L^-1 (lower, upper part zero).
static int cholesky_and_inverse(cudaStream_t st, const LinalgWork& w, int64_t K, int* info) {
  using namespace la;
  const int diag_smem = 2 * NB * (NB + 4) * (int)sizeof(float);
  cudaFuncSetAttribute(potrf_inv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem);
  cudaMemsetAsync(w.Linv, 0, sizeof(float) * K * K, st);
  {
    KernelScope scope("inv_potrf", 0, (double)K * K * K / 3.0, st);
    for (int64_t j = 0; j < K; j += NB) {
      const int nb = (int)std::min<int64_t>(NB, K - j);
      float* Ajj = w.A + j * K + j;
      // factor the diagonal block; its inverse goes to Dinv (dense copy for the panel GEMM) and
      // straight into the diagonal block of L^-1
      {
        KernelScope diag_scope("inv_diag", 0, 0, st);
        potrf_inv_diag_kernel<<<1, 4 * NB, diag_smem, st>>>(Ajj, K, nb, w.Dinv, w.Linv + j * K + j, info,
                                                         (int)j);
      }
      count_launch();
      const int rem = (int)(K - j - nb);
      if (rem > 0) {
        float* A21 = w.A + (j + nb) * K + j;
        // L21 = A21 * L11^-T, in place: the panel is one column tile wide (nb <= BN), so each CTA
        // reads exactly the rows it later overwrites
        sgemm<false, true>(st, rem, nb, nb, 1.f, A21, K, w.Dinv, NB, 0.f, A21, K);
        // A22 -= L21 L21^T (lower tiles)
        float* A22 = w.A + (j + nb) * K + (j + nb);
        sgemm<false, true>(st, rem, rem, nb, -1.f, A21, K, A21, K, 1.f, A22, K, /*tri=*/1);
      }
    }
  }
  // L^-1 by divide and conquer over the block diagonal: with M11, M22 the inverses of two adjacent
  // s x s diagonal blocks and L21 the block below the first,  M21 = -M22 * (L21 * M11).
  // Level s handles all K/(2s) pairs at once (batched launch, operands advance by 2s*(K+1)); the
  // last pair of a level may be ragged.  Every level is two GEMMs with full 2-D parallelism, where
  // a row- or column-sweep exposes only K/128 CTAs.
  {
    KernelScope scope("inv_trtri", 0, (double)K * K * K / 3.0, st);
    for (int64_t s = NB; s < K; s *= 2) {
      const int64_t full = K / (2 * s);                       // pairs with two complete blocks
      const int64_t stride = 2 * s * (K + 1);
      auto level = [&](int64_t a, int64_t s2, int batch) {
        const float* L21 = w.A + (a + s) * K + a;
        float* S21 = w.T + (a + s) * K + a;
        float* M21 = w.Linv + (a + s) * K + a;
        // S21 = L21 * M11          (M11 lower triangular: k starts at the column tile)
        sgemm<false, false>(st, (int)s2, (int)s, (int)s, 1.f, L21, K, w.Linv + a * K + a, K, 0.f, S21,
                            K, /*tri=*/3, batch, stride);
        // M21 = -M22 * S21         (M22 lower triangular: k stops at the row tile)
        sgemm<false, false>(st, (int)s2, (int)s, (int)s2, -1.f, w.Linv + (a + s) * K + (a + s), K, S21,
                            K, 0.f, M21, K, /*tri=*/4, batch, stride);
      };
      if (full > 0) level(0, s, (int)full);
      const int64_t a = full * 2 * s;                         // ragged tail: second block is short
      const int64_t s2 = K - a - s;
      if (s2 > 0) level(a, s2, 1);
    }
  }
  return check_launch("cholesky_and_inverse");
}

// =================================================================================================
// Error-compensated GPTQ column loop (opt-in; the reference sketches it in gptq_quantizer.py:173-197
// and then skips the compensation).  Frantar et al. 2022, Alg. 1, with the asymmetric per-group
// grid of pseudo_quantize_tensor:  for each column j:  q = quant(w_j);  e = (w_j - q) / U[j,j];
// w_{j+1..block end} -= e * U[j, j+1..];  after a block of 128 columns the accumulated errors are
// pushed into all later columns with one GEMM (the lazy rank-128 update).
// Rows are independent: ONE WARP owns one row of the 128-column block, 4 columns per lane in
// registers; the column being quantised is broadcast with a shuffle, every lane applies the rank-1
// update to its own columns from the U block held in shared memory.
// =================================================================================================
namespace gc {
constexpr int B = 128;   // block of columns = lazy-update rank
}

// scale / zero-point of pseudo_quantize_tensor from a (min, max) pair    quantization_utils.py:395-396
__device__ __forceinline__ void asym_params(float mx, float mn, float maxint, float& scale,
                                            float& zp) {
  scale = __fdiv_rn(fmaxf(mx - mn, 1e-5f), maxint);
  zp = clampf(-rintf(__fdiv_rn(mn, scale)), 0.f, maxint);
}

// per-row (min,max) over columns [c0, c0+G) -> scale/zero   (groups wider than one block)
__global__ void __launch_bounds__(256)
row_range_params_kernel(const float* __restrict__ W, int64_t N, int64_t K, int64_t c0, int64_t G,
                        float maxint, float* __restrict__ scales, float* __restrict__ zeros) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t c1 = min(K, c0 + G);
  for (int64_t r = warp; r < N; r += nwarps) {
    float mx = -INFINITY, mn = INFINITY;
    for (int64_t c = c0 + lane; c < c1; c += 32) {
      const float v = W[r * K + c];
      mx = fmaxf(mx, v); mn = fminf(mn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) asym_params(mx, mn, maxint, scales[r], zeros[r]);
  }
}

template <bool OWN_GROUP>
__global__ void __launch_bounds__(256)
gptq_block_kernel(float* __restrict__ W, float* __restrict__ Q, float* __restrict__ Err,
                  const float* __restrict__ U, int64_t N, int64_t K, int64_t c0, int nb,
                  float maxint, const float* __restrict__ scales, const float* __restrict__ zeros) {
  extern __shared__ float Ub[];            // [B][B+1] block of U; row j holds U[c0+j, c0+...]
  constexpr int LD = gc::B + 1;
  for (int i = threadIdx.x; i < gc::B * gc::B; i += blockDim.x) {
    const int r = i / gc::B, c = i % gc::B;
    Ub[r * LD + c] = (r < nb && c < nb) ? U[(c0 + r) * K + (c0 + c)] : (r == c ? 1.f : 0.f);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < N; r += nwarps) {
    float w[4], qv[4], ev[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      w[i] = (c < nb) ? W[r * K + c0 + c] : 0.f;
      qv[i] = 0.f; ev[i] = 0.f;
    }
    float scale, zp;
    if constexpr (OWN_GROUP) {
      float mx = -INFINITY, mn = INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (lane + 32 * i < nb) { mx = fmaxf(mx, w[i]); mn = fminf(mn, w[i]); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      }
      asym_params(mx, mn, maxint, scale, zp);
    } else {
      scale = scales[r]; zp = zeros[r];
    }
    const Divisor sd(scale);
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
      for (int o = 0; o < 32; ++o) {
        const int j = slot * 32 + o;
        if (j >= nb) break;
        const float wj = __shfl_sync(0xffffffffu, w[slot], o);
        const float code = clampf(rintf(sd.div(wj)) + zp, 0.f, maxint);
        const float q = (code - zp) * scale;
        const float e = __fdiv_rn(wj - q, Ub[j * LD + j]);
        if (lane == o) { qv[slot] = q; ev[slot] = e; }
        const float* urow = Ub + j * LD;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = lane + 32 * i;
          if (c > j) w[i] = fmaf(-e, urow[c], w[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i;
      if (c < nb) {
        Q[r * K + c0 + c] = qv[i];
        Err[r * gc::B + c] = ev[i];
      } else {
        Err[r * gc::B + c] = 0.f;
      }
    }
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int64_t b200q_spd_inverse_workspace(int64_t K) {
  if (K <= 0) return 0;
  // T must hold max(rem x NB, NB x K) floats: size it as K x NB
  return linalg_layout(nullptr, K).bytes;
}

// Hinv = inv(H); U (optional) = upper Cholesky factor of inv(H) (U^T U = inv(H)).
// Either output may be NULL.  info (device int, optional; zero it first): 0 = ok, j > 0 = the
// pivot of column j was not positive.
int b200q_spd_inverse(const float* H, float* Hinv, float* U, int64_t K, void* work, int* info,
                      void* stream) {
  B200Q_REQUIRE(H && work && K > 0 && (Hinv || U), "spd_inverse: bad argument");
  B200Q_REQUIRE(K < (1 << 30), "spd_inverse: K too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LinalgWork w = linalg_layout(work, K);
  const int blocks = (int)std::min<int64_t>((K * K + 255) / 256, (int64_t)kNumSMs * 16);
  KernelScope scope("spd_inverse", 0, (double)K * K * K * ((Hinv ? 1.0 : 0.0) + (U ? 2.0 / 3 : 0.0)),
                    st);
  int rc = B200Q_OK;
  if (Hinv != nullptr) {
    cudaMemcpyAsync(w.A, H, sizeof(float) * K * K, cudaMemcpyDeviceToDevice, st);
    rc = cholesky_and_inverse(st, w, K, info);
    if (rc != B200Q_OK) return rc;
    // H^-1 = L^-T L^-1
    sgemm<true, false>(st, (int)K, (int)K, (int)K, 1.f, w.Linv, K, w.Linv, K, 0.f, Hinv, K,
                       /*tri=*/2);
    rc = check_launch("spd_inverse/product");
    if (rc != B200Q_OK) return rc;
  }
  if (U != nullptr) {
    reverse_both_kernel<<<blocks, 256, 0, st>>>(H, w.A, K);
    count_launch();
    rc = cholesky_and_inverse(st, w, K, info);
    if (rc != B200Q_OK) return rc;
    reverse_both_kernel<<<blocks, 256, 0, st>>>(w.Linv, U, K);
    count_launch();
    rc = check_launch("spd_inverse/upper");
  }
  return rc;
}


int64_t b200q_gptq_compensated_workspace(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0) return 0;
  return (int64_t)sizeof(float) * (N * gc::B + 2 * N) + 512;
}

// W (fp32 [N,K], destroyed) -> Q (fp32 [N,K]) with U = upper Cholesky factor of H^-1.
// group: 128 (== block), a larger multiple of 128, or <= 0 (one group per row).
int b200q_gptq_compensated(float* W, float* Q, const float* U, int64_t N, int64_t K, int64_t group,
                           int n_bit, int blocksize, void* work, void* stream) {
  B200Q_REQUIRE(W && Q && U && work && N > 0 && K > 0, "gptq_compensated: bad argument");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 16, "gptq_compensated: n_bit must be in [1,16]");
  if (blocksize != gc::B)
    return fail(B200Q_EUNSUPPORTED, "gptq_compensated: blocksize must be 128");
  const int64_t G = group > 0 ? group : K;
  if (!(G == gc::B || G % gc::B == 0 || G >= K))
    return fail(B200Q_EUNSUPPORTED, "gptq_compensated: group must be 128, a multiple of 128 or per-row");
  B200Q_REQUIRE(K % G == 0 || G >= K, "gptq_compensated: in_features not divisible by group size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("gptq_compensated", 2.0 * N * K * 4, (double)N * K * K, st);
  float* Err = static_cast<float*>(work);
  float* scales = Err + N * gc::B;
  float* zeros = scales + N;
  const float maxint = (float)((1 << n_bit) - 1);
  const int smem = gc::B * (gc::B + 1) * (int)sizeof(float);
  cudaFuncSetAttribute(gptq_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(gptq_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int blocks = (int)std::min<int64_t>((N + 7) / 8, (int64_t)kNumSMs * 2);
  for (int64_t c0 = 0; c0 < K; c0 += gc::B) {
    const int nb = (int)std::min<int64_t>(gc::B, K - c0);
    if (G == gc::B) {
      gptq_block_kernel<true><<<blocks, 256, smem, st>>>(W, Q, Err, U, N, K, c0, nb, maxint, nullptr,
                                                         nullptr);
    } else {
      if (c0 % G == 0) {
        row_range_params_kernel<<<blocks, 256, 0, st>>>(W, N, K, c0, G, maxint, scales, zeros);
        count_launch();
      }
      gptq_block_kernel<false><<<blocks, 256, smem, st>>>(W, Q, Err, U, N, K, c0, nb, maxint, scales,
                                                          zeros);
    }
    count_launch();
    const int64_t rest = K - (c0 + nb);
    if (rest > 0) {
      // W[:, c1:] -= Err[N, nb] * U[c0:c1, c1:]          (lazy rank-128 update)
      sgemm<false, false>(st, (int)N, (int)rest, nb, -1.f, Err, gc::B, U + c0 * K + (c0 + nb), K, 1.f,
                          W + (c0 + nb), K);
    }
  }
  return check_launch("gptq_compensated");
}

}  // extern "C"
