// Interface of the split-fp16 tensor-core GEMM (splitgemm.cu), used by the Cholesky inverse.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace b200q {

enum SplitGemmFlags {
  SG_KB_M = 1,    // A[m, k] = 0 for k < m      : a tile's k-loop starts at its first row
  SG_KB_N = 2,    // B[n, k] = 0 for k < n      : ... at its first column
  SG_KE_M = 4,    // A[m, k] = 0 for k > m      : the k-loop stops after the tile's last row
  SG_KE_N = 8,    // B[n, k] = 0 for k > n      : ... after its last column
  SG_LOWER = 16,  // only the tiles touching the lower triangle of C are computed
  SG_SYMM = 32,   // C symmetric: entries with col <= row are computed and mirrored
};

// An fp32 matrix prepared for the tensor cores: logical [rows, cols = k] K-major fp16 planes.
struct SplitOperand {
  const __half* p[3];       // planes, most significant first
  int planes;               // 2 or 3
  int64_t ld16;
  int rows, cols;
  const float* unscale;     // device: 2^-s
};

constexpr int kSplitMaxPlanes = 3;

// bytes of scratch split_operand needs for an operand with `rows` x `cols` entries (sized for
// kSplitMaxPlanes planes)
int64_t split_operand_bytes(int rows, int cols);

// Split the [src_rows, src_cols] region at src (row stride ld) -- transposed if asked -- into `buf`
// (256-byte aligned, split_operand_bytes large).
int split_operand(cudaStream_t st, const float* src, int64_t ld, int src_rows, int src_cols,
                  bool transpose, int planes, void* buf, SplitOperand* out);

// The [rows, cols] sub-block at (row0, col0) of a prepared operand (col0 a multiple of 8).
inline SplitOperand split_view(const SplitOperand& full, int64_t row0, int64_t col0, int rows, int cols) {
  SplitOperand v = full;
  for (int i = 0; i < 3; ++i) v.p[i] = full.p[i] + row0 * full.ld16 + col0;
  v.rows = rows;
  v.cols = cols;
  return v;
}

// One-time function attributes of the GEMM kernel (call before recording launches into a graph).
int split_gemm_prepare();

// C[M, N] = alpha * A B^T + beta * C   (M = A.rows, N = B.rows, inner = A.cols = B.cols)
int split_gemm(cudaStream_t st, const SplitOperand& A, const SplitOperand& B, float alpha, float beta,
               float* C, int64_t ldc, int flags);

}  // namespace b200q
