// misc_ops.cu — small ops adjacent to the decode path, sm_100a.
//   permute_cols            kernels/permute_cols.cu:1-87          (act-order column gather of A before the Marlin GEMM)
//   awq_dequantize          kernels/quantization/awq/gemm_kernels.cu:720-780 + dequantize.cuh (AWQ int4 -> fp16)
//   advance_step_flashattn  kernels/prepare_inputs/advance_step.cu:13-52   (on-GPU input advance for multi-step decode)
// All three are index / element-wise kernels: bit-exact integer work, 16-byte accesses where the layout allows.
#include "common.cuh"

namespace b200 {

// out[m, k] = a[m, perm[k]]   (2-byte elements)
__global__ void __launch_bounds__(256)
permute_cols_kernel(const uint16_t* __restrict__ a, const int32_t* __restrict__ perm, uint16_t* __restrict__ out,
                    int64_t M, int K) {
  const int64_t total = M * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K;
    const int k = (int)(i % K);
    out[i] = a[m * K + perm[k]];
  }
}

// AWQ: qweight int32 [K, N/8] (nibble p of a word = column [0,2,4,6,1,3,5,7][p] of its group of 8), zeros int32
// [K/G, N/8] packed the same way, scales fp16 [K/G, N]; out fp16 [K, N] = (q - z) * s with fp16 arithmetic
// ((q - z) exact, one rounding in the multiply — sub.f16x2 + fma(.., 0) in the reference).
__global__ void __launch_bounds__(256)
awq_dequantize_kernel(const uint32_t* __restrict__ qweight, const __half* __restrict__ scales,
                      const uint32_t* __restrict__ zeros, __half* __restrict__ out, int64_t K, int N8, int G) {
  const int64_t total = K * N8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / N8;
    const int c = (int)(i % N8);
    const uint32_t q = qweight[i];
    const uint32_t z = zeros[(k / G) * N8 + c];
    const uint4 sv = *reinterpret_cast<const uint4*>(scales + ((k / G) * N8 + c) * 8);
    const __half* s = reinterpret_cast<const __half*>(&sv);
    union { uint4 raw; __half h[8]; } o;
#pragma unroll
    for (int col = 0; col < 8; ++col) {
      const int p = ((col & 1) << 2) | (col >> 1);              // nibble holding logical column `col`
      const int qi = (q >> (4 * p)) & 0xF, zi = (z >> (4 * p)) & 0xF;
      o.h[col] = __hmul(__int2half_rn(qi - zi), s[col]);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = o.raw;
  }
}

__global__ void __launch_bounds__(256)
advance_step_flashattn_kernel(int num_seqs, int num_queries, int block_size, int64_t* input_tokens,
                              const int64_t* sampled_token_ids, int64_t* input_positions, int32_t* seq_lens,
                              int64_t* slot_mapping, const int32_t* block_tables, int64_t block_tables_stride) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= num_queries) return;
  input_tokens[q] = sampled_token_ids[q];
  const int next_len = seq_lens[q] + 1;
  const int next_pos = next_len - 1;
  seq_lens[q] = next_len;
  input_positions[q] = next_pos;
  const int32_t* bt = block_tables + block_tables_stride * q;
  slot_mapping[q] = (int64_t)(bt[next_pos / block_size] * block_size + next_pos % block_size);
}

}  // namespace b200

using namespace b200;

static inline int flat_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int b200_permute_cols(const void* a, const int32_t* perm, void* out, int64_t size_m, int size_k,
                                 void* stream) {
  if (size_m == 0 || size_k == 0) return 0;
  permute_cols_kernel<<<flat_blocks(size_m * size_k), 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)a, perm, (uint16_t*)out, size_m, size_k);
  return check_launch("permute_cols_kernel");
}

extern "C" int b200_awq_dequantize(const void* qweight, const void* scales, const void* zeros, void* out,
                                   int64_t in_c, int qout_c, int group_size, void* stream) {
  B200_CHECK(group_size > 0 && in_c % group_size == 0, "awq_dequantize: in_c must be a multiple of the group size");
  B200_CHECK((reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
             "awq_dequantize: scales / out must be 16-byte aligned");
  if (in_c == 0 || qout_c == 0) return 0;
  awq_dequantize_kernel<<<flat_blocks(in_c * qout_c), 256, 0, (cudaStream_t)stream>>>(
      (const uint32_t*)qweight, (const __half*)scales, (const uint32_t*)zeros, (__half*)out, in_c, qout_c, group_size);
  return check_launch("awq_dequantize_kernel");
}

extern "C" int b200_advance_step_flashattn(int num_seqs, int num_queries, int block_size, int64_t* input_tokens,
                                           const int64_t* sampled_token_ids, int64_t* input_positions,
                                           int32_t* seq_lens, int64_t* slot_mapping, const int32_t* block_tables,
                                           int64_t block_tables_stride, void* stream) {
  if (num_queries == 0) return 0;
  advance_step_flashattn_kernel<<<(num_queries + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      num_seqs, num_queries, block_size, input_tokens, sampled_token_ids, input_positions, seq_lens, slot_mapping,
      block_tables, block_tables_stride);
  return check_launch("advance_step_flashattn_kernel");
}
