// Damped SPD inverse for GPTQ (ref: gptq_quantizer.py:160-165, torch.linalg.inv(H + 1e-6 I)).
//
// H is symmetric positive definite with cond(H) <= (1 + damp)/damp ~ 1e2 by construction (every
// normalised sample has trace 1), so a Cholesky route is safe and costs K^3
// flops against LU's 2 K^3:   H = L L^T  ->  M = L^-1  ->  H^-1 = M^T M.
// The factorisation is RECURSIVE (factor_inv below): a block is halved, the first half factored
// and inverted, the panel and the trailing update formed as two large products, the second half
// factored, and the off-diagonal block of M as two more products.  Those products -- nearly all of
// the K^3 work -- and the final M^T M run on the tcgen05 tensor cores with fp32 operands split
// into fp16 planes and round-to-nearest accumulation across k-chunks (splitgemm.cu).  Blocks of
// <= 1024 columns are factored on the FP32 pipe: a one-CTA register-resident kernel per 64 x 64
// diagonal block, SIMT GEMMs (8x8 register tiles, 128x128x16 CTA tiles) for the rest.
//
// For the error-compensated GPTQ loop the quantity needed is U = chol(H^-1, upper).  With J the
// index reversal, J H J = Lr Lr^T gives H = R R^T with R = J Lr J upper triangular, hence
// H^-1 = R^-T R^-1 and U = R^-1 = J Lr^-1 J: the same factor-and-invert on the reversed matrix.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include "common.cuh"
#include "splitgemm.cuh"

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

// One CTA: Cholesky of the nb x nb (nb <= 64) diagonal block at A (lower, in place; the strict
// upper part of the block is zeroed) and its inverse into the same block of Linv (row stride lda).
// info = first non-positive pivot (1-based, offset by j0).
//
// This kernel sits on the critical path K/64 times, so its LATENCY decides the factorisation time
// (ncu on the earlier shared-memory versions: one SM, <20 % issue utilisation, ~1000 cycles per
// column).  Here the whole block lives in REGISTERS: 256 threads, thread (r, q) holds the 16
// entries [16q, 16q+16) of row r of A and, later, of L^-1.  Both phases are fully unrolled over the
// 64 columns so every register index is static:
//   factor : right-looking.  The owners of column j publish it (rows > j, zeros elsewhere) and the
//            pivot through shared memory; after ONE barrier every thread applies
//            a[r][c] -= (a[r][j] / d) * a[c][j] to its 16 columns.  The zeros make the update a
//            no-op for finished rows and columns, so no predicates are needed.
//   invert : column sweep.  Row k of L^-1 becomes final at step k (divide by L[k][k]), is
//            published, and every later row subtracts L[r][k] times it; L[r][k] comes from the
//            team-mate that owns column k by a shuffle.
// One barrier per column per phase, buffers double-buffered.  Blocks with nb < 64 are padded with
// the identity.
__global__ void __launch_bounds__(256, 1)
potrf_inv_diag_kernel(float* __restrict__ A, int64_t lda, int nb, float* __restrict__ Linv,
                      int* __restrict__ info, int j0) {
  constexpr int NB = la::NB;
  static_assert(NB == 64, "register layout assumes 64 x 64 blocks, 16 columns per thread");
  __shared__ __align__(16) float colbuf[2][NB];
  __shared__ __align__(16) float rowbuf[2][NB];
  __shared__ float pivbuf[2];
  // block <-> registers goes through this tile so that global memory sees whole 256-byte rows per
  // warp instruction (the per-thread layout alone gave 32 sectors per instruction: ncu showed the
  // load/store phases at 27 % of the kernel)
  __shared__ float stage[NB][NB + 1];
  const int tid = threadIdx.x, lane = tid & 31;
  const int r = tid >> 2, q = tid & 3, c0 = q * 16;
  const bool vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(Linv) & 15u) == 0) && nb == NB;
  for (int idx = tid; idx < NB * NB / 4; idx += 256) {
    const int rr = idx >> 4, cc = (idx & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec) {
      if (cc <= rr) v = *reinterpret_cast<const float4*>(A + (int64_t)rr * lda + cc);
    } else if (rr < nb) {
      if (cc < nb) v.x = A[(int64_t)rr * lda + cc];
      if (cc + 1 < nb) v.y = A[(int64_t)rr * lda + cc + 1];
      if (cc + 2 < nb) v.z = A[(int64_t)rr * lda + cc + 2];
      if (cc + 3 < nb) v.w = A[(int64_t)rr * lda + cc + 3];
    }
    stage[rr][cc] = v.x; stage[rr][cc + 1] = v.y; stage[rr][cc + 2] = v.z; stage[rr][cc + 3] = v.w;
  }
  __syncthreads();
  float a[16], x[16];
  float my_rinv = 1.f;                                   // 1 / L[r][r]
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + i;
    a[i] = (r < nb && c < nb) ? (c <= r ? stage[r][c] : 0.f) : (r == c ? 1.f : 0.f);
    x[i] = (c == r) ? 1.f : 0.f;
  }
  // (unrolled over the 16 columns of a chunk only: register indices stay static, and the code is a
  // quarter of the fully unrolled size -- this kernel runs once per launch, so its cost is as much
  // instruction fetch as arithmetic)
#pragma unroll 1
  for (int oq = 0; oq < 4; ++oq)
#pragma unroll
  for (int oi = 0; oi < 16; ++oi) {
    const int j = oq * 16 + oi;
    if (q == oq) {
      const float v = a[oi];
      if (r == j) pivbuf[j & 1] = v;
      colbuf[j & 1][r] = (r > j) ? v : 0.f;
    }
    __syncthreads();
    const float d = pivbuf[j & 1];
    if (tid == 0 && j < nb && !(d > 0.f) && info != nullptr) atomicCAS(info, 0, j0 + j + 1);
    // 1/sqrt(d): MUFU.RSQ plus one Newton step (full fp32 accuracy at a third of the latency of
    // sqrtf followed by a division -- this chain is on the critical path of every column)
    const float dd = fmaxf(d, 1e-30f);
    float rinv = rsqrtf(dd);
    rinv = rinv * fmaf(-0.5f * dd, rinv * rinv, 1.5f);
    const float piv = dd * rinv;
    if (r == j) my_rinv = rinv;
    const float lr = colbuf[j & 1][r] * rinv;            // L[r][j] (0 for r <= j)
    const float s = lr * rinv;
    const float4* cv = reinterpret_cast<const float4*>(&colbuf[j & 1][c0]);
#pragma unroll
    for (int v4 = 0; v4 < 4; ++v4) {
      const float4 cc = cv[v4];
      a[4 * v4 + 0] = fmaf(-s, cc.x, a[4 * v4 + 0]);
      a[4 * v4 + 1] = fmaf(-s, cc.y, a[4 * v4 + 1]);
      a[4 * v4 + 2] = fmaf(-s, cc.z, a[4 * v4 + 2]);
      a[4 * v4 + 3] = fmaf(-s, cc.w, a[4 * v4 + 3]);
    }
    if (q == oq) a[oi] = (r > j) ? lr : (r == j ? piv : 0.f);
  }
  // a[] now holds row r of L (entries right of the diagonal are scratch)
#pragma unroll 1
  for (int oq = 0; oq < 4; ++oq)
#pragma unroll
  for (int oi = 0; oi < 16; ++oi) {
    const int k = oq * 16 + oi;
    const float lrk = __shfl_sync(0xffffffffu, a[oi], (lane & ~3) | oq);     // L[r][k]
    if (r == k) {
      const float dinv = my_rinv;
      float4* rv = reinterpret_cast<float4*>(&rowbuf[k & 1][c0]);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] *= dinv;
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4)
        rv[v4] = make_float4(x[4 * v4], x[4 * v4 + 1], x[4 * v4 + 2], x[4 * v4 + 3]);
    }
    __syncthreads();
    if (r > k) {
      const float4* rv = reinterpret_cast<const float4*>(&rowbuf[k & 1][c0]);
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        const float4 xr = rv[v4];
        x[4 * v4 + 0] = fmaf(-lrk, xr.x, x[4 * v4 + 0]);
        x[4 * v4 + 1] = fmaf(-lrk, xr.y, x[4 * v4 + 1]);
        x[4 * v4 + 2] = fmaf(-lrk, xr.z, x[4 * v4 + 2]);
        x[4 * v4 + 3] = fmaf(-lrk, xr.w, x[4 * v4 + 3]);
      }
    }
  }
  // registers -> tile -> global, once for L and once for L^-1
  auto write_out = [&](const float (&reg)[16], float* __restrict__ dst) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) stage[r][c0 + i] = (c0 + i <= r) ? reg[i] : 0.f;
    __syncthreads();
    for (int idx = tid; idx < NB * NB / 4; idx += 256) {
      const int rr = idx >> 4, cc = (idx & 15) * 4;
      if (vec) {
        *reinterpret_cast<float4*>(dst + (int64_t)rr * lda + cc) =
            make_float4(stage[rr][cc], stage[rr][cc + 1], stage[rr][cc + 2], stage[rr][cc + 3]);
      } else if (rr < nb) {
        for (int u = 0; u < 4; ++u)
          if (cc + u < nb) dst[(int64_t)rr * lda + cc + u] = stage[rr][cc + u];
      }
    }
  };
  write_out(a, A);
  write_out(x, Linv);
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
  float* T;      // [K,K] scratch for the block products of the triangular inverse
  uint8_t* PA;   // fp16 planes of a GEMM operand (up to the full K x K matrix)
  uint8_t* PB;   // fp16 planes of the second operand (up to half the matrix each way)
  int* info;     // first non-positive pivot seen by the diagonal kernels (0 = none)
  int64_t bytes;
};

namespace la {
// blocks up to this size are factored by the SIMT path (B200Q_INVERSE_LEAF overrides, read once)
inline int64_t leaf() {
  static const int64_t v = []() {
    const char* e = std::getenv("B200Q_INVERSE_LEAF");
    const long x = e != nullptr ? std::atol(e) : 0;
    return (int64_t)((x >= 64 && x <= 4096) ? x : 1024);   // measured: 128..1024 within 10 %, 1024 best
  }();
  return v;
}
inline int64_t first_half(int64_t n) { return ((n / 2 + 127) / 128) * 128; }
// fp16 planes per operand of the tensor-core products.  Two (22 bits) are enough: measured against
// an fp64 inverse the result is as accurate as with three (33 bits) -- the error that remains is
// the factorisation's own -- at two thirds of the MMAs.  B200Q_INVERSE_PLANES=3 overrides (read once).
inline int planes() {
  static const int p = []() {
    const char* e = std::getenv("B200Q_INVERSE_PLANES");
    return (e != nullptr && e[0] == '3') ? 3 : 2;
  }();
  return p;
}
}  // namespace la

static LinalgWork linalg_layout(void* work, int64_t K) {
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  LinalgWork w;
  uint8_t* base = static_cast<uint8_t*>(work);
  int64_t off = 0;
  w.A = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.Linv = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.T = reinterpret_cast<float*>(base + off); off += align(4 * K * K);
  w.PA = base + off; off += align(split_operand_bytes((int)K, (int)K));
  const int half = (int)std::min<int64_t>(K, la::first_half(K) + 128);
  w.PB = base + off; off += align(split_operand_bytes(half, half));
  w.info = reinterpret_cast<int*>(base + off); off += 256;
  w.bytes = off;
  return w;
}

// Leaf: the n x n diagonal block at (a0, a0) of A (row stride ld) -> L in place and L^-1 into
// Linv, everything on the FP32 pipe.  Right-looking over 64-column steps; the triangular inverse
// of the block by divide and conquer with batched GEMMs.
static void leaf_factor_inv(cudaStream_t st, const LinalgWork& w, int64_t ld, int64_t a0, int64_t n,
                            int* info) {
  using namespace la;
  float* Ab = w.A + a0 * ld + a0;
  float* Mb = w.Linv + a0 * ld + a0;
  float* Tb = w.T + a0 * ld + a0;
  for (int64_t j = 0; j < n; j += NB) {
    const int nb = (int)std::min<int64_t>(NB, n - j);
    float* Ajj = Ab + j * ld + j;
    {
      KernelScope diag_scope("inv_diag", 0, 0, st);
      potrf_inv_diag_kernel<<<1, 4 * NB, 0, st>>>(Ajj, ld, nb, Mb + j * ld + j, info, (int)(a0 + j));
    }
    count_launch();
    const int rem = (int)(n - j - nb);
    if (rem > 0) {
      float* A21 = Ab + (j + nb) * ld + j;
      // L21 = A21 * L11^-T, in place: the panel is one column tile wide (nb <= BN), so each CTA
      // reads exactly the rows it later overwrites
      sgemm<false, true>(st, rem, nb, nb, 1.f, A21, ld, Mb + j * ld + j, ld, 0.f, A21, ld);
      // A22 -= L21 L21^T (lower tiles)
      float* A22 = Ab + (j + nb) * ld + (j + nb);
      sgemm<false, true>(st, rem, rem, nb, -1.f, A21, ld, A21, ld, 1.f, A22, ld, /*tri=*/1);
    }
  }
  // L^-1 by divide and conquer over the block diagonal: with M11, M22 the inverses of two adjacent
  // s x s diagonal blocks and L21 the block below the first,  M21 = -M22 * (L21 * M11).
  // Level s handles all n/(2s) pairs at once (batched launch, operands advance by 2s*(ld+1)); the
  // last pair of a level may be ragged.
  for (int64_t s = NB; s < n; s *= 2) {
    const int64_t full = n / (2 * s);                       // pairs with two complete blocks
    const int64_t stride = 2 * s * (ld + 1);
    auto level = [&](int64_t a, int64_t s2, int batch) {
      const float* L21 = Ab + (a + s) * ld + a;
      float* S21 = Tb + (a + s) * ld + a;
      float* M21 = Mb + (a + s) * ld + a;
      // S21 = L21 * M11          (M11 lower triangular: k starts at the column tile)
      sgemm<false, false>(st, (int)s2, (int)s, (int)s, 1.f, L21, ld, Mb + a * ld + a, ld, 0.f, S21,
                          ld, /*tri=*/3, batch, stride);
      // M21 = -M22 * S21         (M22 lower triangular: k stops at the row tile)
      sgemm<false, false>(st, (int)s2, (int)s, (int)s2, -1.f, Mb + (a + s) * ld + (a + s), ld, S21,
                          ld, 0.f, M21, ld, /*tri=*/4, batch, stride);
    };
    if (full > 0) level(0, s, (int)full);
    const int64_t a = full * 2 * s;                         // ragged tail: second block is short
    const int64_t s2 = n - a - s;
    if (s2 > 0) level(a, s2, 1);
  }
}

// Recursive blocked factor-and-invert of the n x n diagonal block at (a0, a0):
//     A = [A11 . ; A21 A22]   ->   L11, M11 = L11^-1        (recursion)
//                                  L21 = A21 M11^T           (GEMM; M11 triangular: half the k range)
//                                  A22 -= L21 L21^T          (GEMM, lower tiles)
//                                  L22, M22                  (recursion)
//                                  M21 = -M22 (L21 M11)      (two GEMMs, triangular k ranges)
// so that nearly all of the K^3 work sits in a few large products, which run on the tensor cores
// (splitgemm.cu); only blocks of <= la::leaf() columns use the FP32 pipe.
static int factor_inv(cudaStream_t st, const LinalgWork& w, int64_t K, int64_t a0, int64_t n,
                      int* info) {
  if (n <= la::leaf()) {
    leaf_factor_inv(st, w, K, a0, n, info);
    return B200Q_OK;
  }
  const int64_t n1 = la::first_half(n), n2 = n - n1;
  int rc = factor_inv(st, w, K, a0, n1, info);
  if (rc != B200Q_OK) return rc;
  float* A21 = w.A + (a0 + n1) * K + a0;
  float* A22 = w.A + (a0 + n1) * K + (a0 + n1);
  const float* M11 = w.Linv + a0 * K + a0;
  const float* M22 = w.Linv + (a0 + n1) * K + (a0 + n1);
  float* M21 = w.Linv + (a0 + n1) * K + a0;
  float* S21 = w.T + (a0 + n1) * K + a0;
  SplitOperand a, b;
#define LA_TRY(x) do { rc = (x); if (rc != B200Q_OK) return rc; } while (0)
  LA_TRY(split_operand(st, A21, K, (int)n2, (int)n1, false, la::planes(), w.PA, &a));
  LA_TRY(split_operand(st, M11, K, (int)n1, (int)n1, false, la::planes(), w.PB, &b));
  LA_TRY(split_gemm(st, a, b, 1.f, 0.f, A21, K, SG_KE_N));              // L21 = A21 M11^T
  LA_TRY(split_operand(st, A21, K, (int)n2, (int)n1, false, la::planes(), w.PA, &a));
  LA_TRY(split_gemm(st, a, a, -1.f, 1.f, A22, K, SG_LOWER));            // A22 -= L21 L21^T
  LA_TRY(factor_inv(st, w, K, a0 + n1, n2, info));
  LA_TRY(split_operand(st, A21, K, (int)n2, (int)n1, false, la::planes(), w.PA, &a));  // the recursion reused PA
  LA_TRY(split_operand(st, M11, K, (int)n1, (int)n1, true, la::planes(), w.PB, &b));   // B[n][k] = M11[k][n]
  LA_TRY(split_gemm(st, a, b, 1.f, 0.f, S21, K, SG_KB_N));              // S = L21 M11
  LA_TRY(split_operand(st, M22, K, (int)n2, (int)n2, false, la::planes(), w.PA, &a));
  LA_TRY(split_operand(st, S21, K, (int)n2, (int)n1, true, la::planes(), w.PB, &b));   // B[n][k] = S[k][n]
  LA_TRY(split_gemm(st, a, b, -1.f, 0.f, M21, K, SG_KE_M));             // M21 = -M22 S
  return B200Q_OK;
}

__global__ void merge_info_kernel(const int* __restrict__ src, int* __restrict__ dst) {
  if (*src != 0) atomicCAS(dst, 0, *src);
}

// The factorisation of one matrix is ~500 small dependent launches (64-column diagonal kernels,
// leaf GEMMs, operand splits) whose host-side cost -- launch calls and tensor-map encodes -- would
// leave the GPU waiting.  The sequence only depends on K and on the workspace address, so it is
// recorded ONCE into a CUDA graph (captured on a private stream: the caller's may be the legacy
// default stream, which cannot be captured) and replayed with a single launch afterwards.
// B200Q_INVERSE_GRAPH=0 keeps the direct launches (per-stage profiling scopes need them).
struct FactorGraph {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
};
static std::mutex g_graph_mu;
static std::map<std::tuple<int, const void*, int64_t>, FactorGraph> g_graphs;

static bool graphs_enabled() {
  static const bool on = []() {
    const char* e = std::getenv("B200Q_INVERSE_GRAPH");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

static int factor_direct(cudaStream_t st, const LinalgWork& w, int64_t K) {
  cudaMemsetAsync(w.Linv, 0, sizeof(float) * K * K, st);
  cudaMemsetAsync(w.info, 0, sizeof(int), st);
  int rc = factor_inv(st, w, K, 0, K, w.info);
  if (rc != B200Q_OK) return rc;
  return check_launch("cholesky_and_inverse");
}

// A (K x K, lower part valid) -> L in place (lower), Linv = L^-1 (lower, upper part zero).
static int cholesky_and_inverse(cudaStream_t st, const LinalgWork& w, int64_t K, int* info) {
  KernelScope scope("inv_factor", 0, 2.0 * (double)K * K * K / 3.0, st);
  int rc = B200Q_OK;
  if (!graphs_enabled()) {
    rc = factor_direct(st, w, K);
  } else {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_graph_mu);
    const auto key = std::make_tuple(dev, static_cast<const void*>(w.A), K);
    auto it = g_graphs.find(key);
    if (it == g_graphs.end()) {
      // One graph per (workspace, size): the model walker keeps up to 16 streams' workspaces busy with
      // two or three matrix sizes each, so well over 32 graphs are live at a time -- re-capturing
      // (tens of ms of host time per graph) must only happen when workspaces really come and go.
      if (g_graphs.size() >= 256) {
        for (auto& kv : g_graphs) cudaGraphExecDestroy(kv.second.exec);
        g_graphs.clear();
      }
      rc = split_gemm_prepare();                 // function attributes: not during capture
      if (rc != B200Q_OK) return rc;
      cudaStream_t cap = nullptr;
      if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess)
        return fail(B200Q_ECUDA, "spd_inverse: cannot create the capture stream");
      FactorGraph fg;
      cudaGraph_t graph = nullptr;
      const int64_t before = launches_so_far();
      set_stream_capture(true);
      cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        rc = factor_direct(cap, w, K);
        e = cudaStreamEndCapture(cap, &graph);
      }
      set_stream_capture(false);
      fg.launches = launches_so_far() - before;
      count_launch((int)-fg.launches);           // nothing has run yet; replays are counted below
      if (e == cudaSuccess && rc == B200Q_OK) e = cudaGraphInstantiate(&fg.exec, graph, 0);
      if (graph != nullptr) cudaGraphDestroy(graph);
      cudaStreamDestroy(cap);
      if (rc != B200Q_OK) return rc;
      if (e != cudaSuccess)
        return fail(B200Q_ECUDA, std::string("spd_inverse: graph capture: ") + cudaGetErrorString(e));
      it = g_graphs.emplace(key, fg).first;
    }
    if (cudaGraphLaunch(it->second.exec, st) != cudaSuccess)
      return fail(B200Q_ECUDA, "spd_inverse: graph launch failed");
    count_launch((int)it->second.launches);
  }
  if (rc != B200Q_OK) return rc;
  if (info != nullptr) {
    merge_info_kernel<<<1, 1, 0, st>>>(w.info, info);
    count_launch();
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

// GROUP_MODE 0: the quantisation group IS the 128-column block (q_group_size = 128): scale and
// zero point come from the block's own registers.
// GROUP_MODE 1: any group size G (K % G == 0, or one group per row).  A group starts wherever the
// global column index is a multiple of G; its (min, max) is taken over the group's columns AS THEY
// ARE at that moment (Alg. 1 / the oracle: `if j % G == 0: blk = W[:, j:j+G]`): columns inside the
// current block come from the registers (they carry the in-block updates), columns beyond the
// block from global memory (they carry the lazy updates of all earlier blocks and nothing of this
// one, exactly as in the lazily-updated algorithm).  A group that began in an earlier block hands
// its parameters over through the per-row scales / zeros arrays.
template <int GROUP_MODE>
__global__ void __launch_bounds__(256)
gptq_block_kernel(float* __restrict__ W, float* __restrict__ Q, float* __restrict__ Err,
                  const float* __restrict__ U, int64_t N, int64_t K, int64_t c0, int nb,
                  float maxint, int64_t G, float* __restrict__ scales, float* __restrict__ zeros) {
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
    float scale = 1.f, zp = 0.f;
    if constexpr (GROUP_MODE == 0) {
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
      if (c0 % G != 0) { scale = scales[r]; zp = zeros[r]; }   // the group began in an earlier block
    }
    Divisor sd(scale);
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
      for (int o = 0; o < 32; ++o) {
        const int j = slot * 32 + o;
        if (j >= nb) break;
        if constexpr (GROUP_MODE == 1) {
          if ((c0 + j) % G == 0) {
            const int64_t g_end = min(K, c0 + j + G);          // global column range [c0 + j, g_end)
            float mx = -INFINITY, mn = INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = lane + 32 * i;
              if (c >= j && c < nb && c0 + c < g_end) { mx = fmaxf(mx, w[i]); mn = fminf(mn, w[i]); }
            }
            for (int64_t c = c0 + nb + lane; c < g_end; c += 32) {
              const float v = W[r * K + c];
              mx = fmaxf(mx, v); mn = fminf(mn, v);
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
              mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
              mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
            }
            asym_params(mx, mn, maxint, scale, zp);
            sd = Divisor(scale);
          }
        }
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
    if constexpr (GROUP_MODE == 1) {
      if (lane == 0) { scales[r] = scale; zeros[r] = zp; }
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
    {
      KernelScope product_scope("inv_product", 0, (double)K * K * K / 3.0, st);
      SplitOperand mt;                                    // A[m][k] = B[m][k] = Linv[k][m]
      rc = split_operand(st, w.Linv, K, (int)K, (int)K, true, la::planes(), w.PA, &mt);
      if (rc != B200Q_OK) return rc;
      rc = split_gemm(st, mt, mt, 1.f, 0.f, Hinv, K, SG_KB_M | SG_KB_N | SG_SYMM);
      if (rc != B200Q_OK) return rc;
    }
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
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  return align((int64_t)sizeof(float) * (N * gc::B + 2 * N)) + align(split_operand_bytes((int)N, gc::B)) +
         align(split_operand_bytes((int)K, (int)K)) + 512;
}

// W (fp32 [N,K], destroyed) -> Q (fp32 [N,K]) with U = upper Cholesky factor of H^-1.
// group: any size dividing K, or <= 0 (one group per row); 128 takes the register-only fast path.
// blocksize: the lazy-update batch of Alg. 1.  Values up to 128 are honoured exactly (the kernel
// takes that many columns per launch).  Larger values batch by 128: the lazy update only regroups
// the SAME rank-1 updates, so the result differs from a true larger batch only where a quantisation
// group straddles a batch boundary and its (min, max) is taken over columns that have / have not
// yet received the pending updates (group sizes dividing 128, or multiples of the blocksize, are
// unaffected).
int b200q_gptq_compensated(float* W, float* Q, const float* U, int64_t N, int64_t K, int64_t group,
                           int n_bit, int blocksize, void* work, void* stream) {
  B200Q_REQUIRE(W && Q && U && work && N > 0 && K > 0, "gptq_compensated: bad argument");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 16, "gptq_compensated: n_bit must be in [1,16]");
  B200Q_REQUIRE(blocksize >= 1, "gptq_compensated: blocksize must be positive");
  const int64_t G = (group > 0 && group < K) ? group : K;
  B200Q_REQUIRE(K % G == 0, "gptq_compensated: in_features not divisible by group size");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope scope("gptq_compensated", 2.0 * N * K * 4, (double)N * K * K, st);
  B200Q_REQUIRE(N < (1 << 30) && K < (1 << 30), "gptq_compensated: dimension too large");
  auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
  B200Q_REQUIRE((reinterpret_cast<uintptr_t>(work) & 255u) == 0, "gptq_compensated: workspace not 256-byte aligned");
  float* Err = static_cast<float*>(work);
  float* scales = Err + N * gc::B;
  float* zeros = scales + N;
  uint8_t* planes_err = static_cast<uint8_t*>(work) + align((int64_t)sizeof(float) * (N * gc::B + 2 * N));
  uint8_t* planes_u = planes_err + align(split_operand_bytes((int)N, gc::B));
  const float maxint = (float)((1 << n_bit) - 1);
  const int smem = gc::B * (gc::B + 1) * (int)sizeof(float);
  cudaFuncSetAttribute(gptq_block_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(gptq_block_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int blocks = (int)std::min<int64_t>((N + 7) / 8, (int64_t)kNumSMs * 2);
  // U^T as tensor-core planes, once: the lazy update of block c0 multiplies by U[c0:c1, c1:], whose
  // transpose is the sub-block [c1:, c0:c1] of these planes
  SplitOperand ut;
  int rc = split_operand(st, U, K, (int)K, (int)K, true, 2, planes_u, &ut);
  if (rc != B200Q_OK) return rc;
  const int step = std::min(blocksize, gc::B);
  for (int64_t c0 = 0; c0 < K; c0 += step) {
    const int nb = (int)std::min<int64_t>(step, K - c0);
    if (G == gc::B && step == gc::B) {
      gptq_block_kernel<0><<<blocks, 256, smem, st>>>(W, Q, Err, U, N, K, c0, nb, maxint, G, nullptr,
                                                      nullptr);
    } else {
      gptq_block_kernel<1><<<blocks, 256, smem, st>>>(W, Q, Err, U, N, K, c0, nb, maxint, G, scales,
                                                      zeros);
    }
    count_launch();
    const int64_t rest = K - (c0 + nb);
    if (rest > 0) {
      // W[:, c1:] -= Err[N, nb] * U[c0:c1, c1:]    (lazy rank-128 update, on the tensor cores:
      // fp32 operands split into fp16 planes, splitgemm.cu)
      SplitOperand e;
      rc = split_operand(st, Err, gc::B, (int)N, nb, false, 2, planes_err, &e);
      if (rc != B200Q_OK) return rc;
      const SplitOperand u = split_view(ut, c0 + nb, c0, (int)rest, nb);
      rc = split_gemm(st, e, u, -1.f, 1.f, W + (c0 + nb), K, 0);
      if (rc != B200Q_OK) return rc;
    }
  }
  return check_launch("gptq_compensated");
}

}  // extern "C"
