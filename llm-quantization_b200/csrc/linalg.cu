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
constexpr int NB = 128;
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
             int64_t ldc, int tri) {
  using namespace la;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (tri == 1 && n0 > m0 + BM - 1) return;
  const int k_first = (tri == 2) ? (max(m0, n0) / BK) : (tri == 3 ? n0 / BK : 0);
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
                  const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int tri = 0) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + la::BN - 1) / la::BN, (M + la::BM - 1) / la::BM);
  sgemm_kernel<TA, TB><<<grid, 256, 0, st>>>(M, N, Kd, alpha, A, lda, B, ldb, beta, C, ldc, tri);
  count_launch();
}

// One CTA: Cholesky of the nb x nb diagonal block at A (lower, in place; the strict upper part of
// the block is zeroed) and its inverse into Linv (nb x nb, ld = NB, lower).  info = first
// non-positive pivot (1-based, offset by j0), left untouched on success.
__global__ void __launch_bounds__(256)
potrf_inv_diag_kernel(float* __restrict__ A, int64_t lda, int nb, float* __restrict__ Linv,
                      int* __restrict__ info, int j0) {
  extern __shared__ float sm[];
  float* L = sm;                          // [NB][NB+1]
  float* X = sm + la::NB * (la::NB + 1);  // [NB][NB+1]
  const int tid = threadIdx.x;
  constexpr int LD = la::NB + 1;
  for (int i = tid; i < nb * nb; i += blockDim.x) {
    const int r = i / nb, c = i % nb;
    L[r * LD + c] = (c <= r) ? A[(int64_t)r * lda + c] : 0.f;
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const float d = L[k * LD + k];
    if (!(d > 0.f)) {
      if (tid == 0 && info != nullptr) atomicCAS(info, 0, j0 + k + 1);
    }
    const float piv = sqrtf(fmaxf(d, 1e-30f));
    __syncthreads();
    // scale column k
    for (int r = k + tid; r < nb; r += blockDim.x)
      L[r * LD + k] = (r == k) ? piv : L[r * LD + k] / piv;
    __syncthreads();
    // rank-1 update of the trailing lower triangle
    const int rem = nb - k - 1;
    for (int i = tid; i < rem * rem; i += blockDim.x) {
      const int r = k + 1 + i / rem, c = k + 1 + i % rem;
      if (c <= r) L[r * LD + c] = fmaf(-L[r * LD + k], L[c * LD + k], L[r * LD + c]);
    }
    __syncthreads();
  }
  // inverse by forward substitution, one column per thread: L x = e_c
  for (int c = tid; c < nb; c += blockDim.x) {
    for (int r = 0; r < nb; ++r) {
      if (r < c) { X[r * LD + c] = 0.f; continue; }
      float s = (r == c) ? 1.f : 0.f;
      for (int k = c; k < r; ++k) s = fmaf(-L[r * LD + k], X[k * LD + c], s);
      X[r * LD + c] = s / L[r * LD + r];
    }
  }
  __syncthreads();
  for (int i = tid; i < nb * nb; i += blockDim.x) {
    const int r = i / nb, c = i % nb;
    A[(int64_t)r * lda + c] = L[r * LD + c];
    Linv[r * la::NB + c] = X[r * LD + c];
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

__global__ void copy_block_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                  int64_t ldd, int rows, int cols) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
    const int r = i / cols, c = i % cols;
    dst[(int64_t)r * ldd + c] = src[(int64_t)r * lds + c];
  }
}

struct LinalgWork {
  float* A;      // [K,K] working copy -> L (lower)
  float* Linv;   // [K,K] -> L^-1 (lower)
  float* Dinv;   // [NB,NB] inverse of the current diagonal block
  float* T;      // [NB,K] row-block temporary
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
  w.T = reinterpret_cast<float*>(base + off); off += align(4 * la::NB * K);
  w.bytes = off;
  return w;
}

// A (K x K, lower part valid) -> L in place (lower), Linv = L^-1 (lower, upper part zero).
static int cholesky_and_inverse(cudaStream_t st, const LinalgWork& w, int64_t K, int* info) {
  using namespace la;
  const int diag_smem = 2 * NB * (NB + 1) * (int)sizeof(float);
  cudaFuncSetAttribute(potrf_inv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem);
  cudaMemsetAsync(w.Linv, 0, sizeof(float) * K * K, st);
  for (int64_t j = 0; j < K; j += NB) {
    const int nb = (int)std::min<int64_t>(NB, K - j);
    float* Ajj = w.A + j * K + j;
    potrf_inv_diag_kernel<<<1, 256, diag_smem, st>>>(Ajj, K, nb, w.Dinv, info, (int)j);
    count_launch();
    // diagonal block of L^-1
    copy_block_kernel<<<16, 256, 0, st>>>(w.Dinv, NB, w.Linv + j * K + j, K, nb, nb);
    count_launch();
    const int rem = (int)(K - j - nb);
    if (rem > 0) {
      float* A21 = w.A + (j + nb) * K + j;
      // L21 = A21 * L11^-T   -> into T (rem x nb), then back
      sgemm<false, true>(st, rem, nb, nb, 1.f, A21, K, w.Dinv, NB, 0.f, w.T, NB);
      // reuse: copy T back over A21 (row-block temporaries are at most NB wide, so T is [rem, NB])
      copy_block_kernel<<<(unsigned)std::min<int64_t>(1024, ((int64_t)rem * nb + 255) / 256), 256, 0,
                          st>>>(w.T, NB, A21, K, rem, nb);
      count_launch();
      // A22 -= L21 L21^T (lower tiles)
      float* A22 = w.A + (j + nb) * K + (j + nb);
      sgemm<false, true>(st, rem, rem, nb, -1.f, A21, K, A21, K, 1.f, A22, K, /*tri=*/1);
    }
  }
  // L^-1 below the diagonal, block row i:  Linv[i, 0:i] = -Linv[i,i] * (L[i, 0:i] * Linv[0:i, 0:i])
  for (int64_t i = NB; i < K; i += NB) {
    const int nb = (int)std::min<int64_t>(NB, K - i);
    const float* Li = w.A + i * K;                 // L[i-block, 0:i]
    sgemm<false, false>(st, nb, (int)i, (int)i, 1.f, Li, K, w.Linv, K, 0.f, w.T, K, /*tri=*/3);
    sgemm<false, false>(st, nb, (int)i, nb, -1.f, w.Linv + i * K + i, K, w.T, K, 0.f,
                        w.Linv + i * K, K);
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
