// Packed integer export: the reference only fake-quantizes and never
// stores codes (benchmark_runner.py:732-743 saves metrics); these kernels turn the uint8 code planes
// the quantizer kernels already emit into the on-disk form and back.
//
// Layout: every row is an independent little-endian bit stream; code k of a row occupies bits
// [k*b, (k+1)*b) of the stream, which is stored as ceil(K*b/32) uint32 words (the last word zero
// padded).  For b = 4 that is the familiar eight codes per int32, lowest nibble first; b = 3 packs
// 32 codes into three words.  HBM-bound byte work: one thread per output word, 128-bit loads of the
// codes where the word's codes are 16-byte aligned (b = 2, 4, 8), byte loads otherwise.
#include "common.cuh"

namespace b200q {

__global__ void __launch_bounds__(256)
pack_codes_kernel(const uint8_t* __restrict__ codes, int64_t N, int64_t K, int n_bit,
                  int64_t words_per_row, uint32_t* __restrict__ packed) {
  const int64_t total = N * words_per_row;
  const uint32_t mask = (n_bit >= 32) ? 0xffffffffu : ((1u << n_bit) - 1u);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / words_per_row, w = i % words_per_row;
    const uint8_t* row = codes + r * K;
    const int64_t bit0 = w * 32;
    int64_t k = bit0 / n_bit;                         // first code with bits in this word
    int shift = (int)(k * n_bit - bit0);              // <= 0: bits of code k below the word start
    uint32_t word = 0;
    if (n_bit == 4 && shift == 0 && k + 8 <= K && ((reinterpret_cast<uintptr_t>(row + k) & 7u) == 0)) {
      const uint2 raw = *reinterpret_cast<const uint2*>(row + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        word |= ((raw.x >> (8 * j)) & 15u) << (4 * j);
        word |= ((raw.y >> (8 * j)) & 15u) << (16 + 4 * j);
      }
    } else {
      for (; k < K && shift < 32; ++k, shift += n_bit) {
        const uint32_t c = (uint32_t)row[k] & mask;
        word |= (shift >= 0) ? (c << shift) : (c >> (-shift));
      }
    }
    packed[i] = word;
  }
}

__global__ void __launch_bounds__(256)
unpack_codes_kernel(const uint32_t* __restrict__ packed, int64_t N, int64_t K, int n_bit,
                    int64_t words_per_row, uint8_t* __restrict__ codes) {
  const int64_t total = N * K;
  const uint32_t mask = (1u << n_bit) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K, k = i % K;
    const int64_t bit = k * n_bit;
    const uint32_t* row = packed + r * words_per_row;
    const int64_t w = bit >> 5;
    const int off = (int)(bit & 31);
    uint32_t v = row[w] >> off;
    if (off + n_bit > 32) v |= row[w + 1] << (32 - off);
    codes[i] = (uint8_t)(v & mask);
  }
}

}  // namespace b200q

using namespace b200q;

extern "C" {

int64_t b200q_packed_words_per_row(int64_t K, int n_bit) {
  if (K <= 0 || n_bit < 1 || n_bit > 8) return 0;
  return (K * n_bit + 31) / 32;
}

int b200q_pack_codes(const uint8_t* codes, int64_t N, int64_t K, int n_bit, uint32_t* packed,
                     void* stream) {
  B200Q_REQUIRE(codes && packed && N > 0 && K > 0, "pack_codes: bad argument");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 8, "pack_codes: n_bit must be in [1,8]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t wpr = b200q_packed_words_per_row(K, n_bit);
  KernelScope scope("pack_codes", (double)N * K + 4.0 * N * wpr, 0, st);
  const int blocks = (int)std::min<int64_t>((N * wpr + 255) / 256, (int64_t)kNumSMs * 16);
  pack_codes_kernel<<<blocks, 256, 0, st>>>(codes, N, K, n_bit, wpr, packed);
  count_launch();
  return check_launch("pack_codes");
}

int b200q_unpack_codes(const uint32_t* packed, int64_t N, int64_t K, int n_bit, uint8_t* codes,
                       void* stream) {
  B200Q_REQUIRE(codes && packed && N > 0 && K > 0, "unpack_codes: bad argument");
  B200Q_REQUIRE(n_bit >= 1 && n_bit <= 8, "unpack_codes: n_bit must be in [1,8]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t wpr = b200q_packed_words_per_row(K, n_bit);
  KernelScope scope("unpack_codes", (double)N * K + 4.0 * N * wpr, 0, st);
  const int blocks = (int)std::min<int64_t>((N * K + 255) / 256, (int64_t)kNumSMs * 16);
  unpack_codes_kernel<<<blocks, 256, 0, st>>>(packed, N, K, n_bit, wpr, codes);
  count_launch();
  return check_launch("unpack_codes");
}

}  // extern "C"
