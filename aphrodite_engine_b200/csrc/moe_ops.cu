// moe_ops.cu — MoE routing helpers, sm_100a.
//
// Replaces moe_align_block_size (kernels/moe/align_block_size_kernel.cu:22-133, schema
// kernels/torch_bindings.cpp:394-399) and topk_softmax (kernels/moe/softmax.cu:107-161, :496-518, schema
// kernels/moe/torch_bindings.cpp:11-14). Both are tiny latency-bound integer / reduction kernels.
//
// moe_align_block_size is a STABLE counting sort of the flat (token, k) slots by expert id, each expert's
// segment padded to a multiple of block_size — bit-exact with the reference: inside an expert the slots keep
// increasing index order, padding slots are left untouched (the caller pre-fills them with numel), expert_ids
// is written for every block of every segment and total_tokens_post_pad is the padded total.
// The reference launches one CTA of `num_experts` threads with an (E+1) x E shared counter matrix; here the
// thread count (256) is independent of E and the counter matrix is (T+1) x E.
#include "common.cuh"

#include <algorithm>

namespace b200 {

template <typename I>
__global__ void __launch_bounds__(256)
moe_align_kernel(const I* __restrict__ topk_ids, int32_t* __restrict__ sorted_ids,
                 int32_t* __restrict__ expert_ids, int32_t* __restrict__ total_post_pad,
                 int num_experts, int block_size, int64_t numel, int T) {
  extern __shared__ int32_t sm[];
  int32_t* cnt = sm;                              // [(T+1)][E]
  int32_t* cumsum = sm + (size_t)(T + 1) * num_experts;  // [E+1]
  const int t = threadIdx.x;
  const int64_t per = (numel + T - 1) / T;
  const int64_t lo = (int64_t)t * per, hi = min(numel, lo + per);
  if (t < T)
    for (int e = 0; e < num_experts; ++e) cnt[(size_t)(t + 1) * num_experts + e] = 0;
  if (t == 0)
    for (int e = 0; e < num_experts; ++e) cnt[e] = 0;
  __syncthreads();
  if (t < T)
    for (int64_t i = lo; i < hi; ++i) ++cnt[(size_t)(t + 1) * num_experts + (int)topk_ids[i]];
  __syncthreads();
  // exclusive prefix over threads, per expert: cnt[t][e] = #slots of expert e in shards < t
  for (int e = t; e < num_experts; e += blockDim.x) {
    int32_t run = 0;
    for (int s = 1; s <= T; ++s) {
      const int32_t c = cnt[(size_t)s * num_experts + e];
      cnt[(size_t)s * num_experts + e] = run;       // row s = count of shard s-1 -> becomes its exclusive prefix
      run += c;
    }
    cnt[e] = run;                                   // row 0: total per expert
  }
  __syncthreads();
  if (t == 0) {
    cumsum[0] = 0;
    for (int e = 0; e < num_experts; ++e)
      cumsum[e + 1] = cumsum[e] + (cnt[e] + block_size - 1) / block_size * block_size;
    *total_post_pad = cumsum[num_experts];
  }
  __syncthreads();
  for (int e = t; e < num_experts; e += blockDim.x)
    for (int i = cumsum[e]; i < cumsum[e + 1]; i += block_size) expert_ids[i / block_size] = e;
  if (t < T) {
    int32_t* mine = cnt + (size_t)(t + 1) * num_experts;  // running rank of this shard inside each expert
    for (int64_t i = lo; i < hi; ++i) {
      const int e = (int)topk_ids[i];
      sorted_ids[cumsum[e] + mine[e]++] = (int32_t)i;
    }
  }
}

// one warp per token: softmax over experts (fp32), then k rounds of arg-max (ties -> lowest expert id)
__global__ void __launch_bounds__(128)
topk_softmax_kernel(const float* __restrict__ gating, float* __restrict__ weights,
                    int32_t* __restrict__ indices, int32_t* __restrict__ source_rows, int num_tokens,
                    int num_experts, int topk) {
  extern __shared__ float probs[];  // [warps][num_experts]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= num_tokens) return;
  float* p = probs + (size_t)warp * num_experts;
  const float* g = gating + (size_t)row * num_experts;
  float mx = -INFINITY;
  for (int e = lane; e < num_experts; e += 32) mx = fmaxf(mx, g[e]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int e = lane; e < num_experts; e += 32) {
    const float v = expf(g[e] - mx);
    p[e] = v;
    sum += v;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int e = lane; e < num_experts; e += 32) p[e] *= inv;
  __syncwarp();
  for (int k = 0; k < topk; ++k) {
    float best = -1.f;  // probabilities are >= 0
    int bi = 0;
    for (int e = lane; e < num_experts; e += 32) {
      const float v = p[e];
      if (v > best) { best = v; bi = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      weights[(size_t)row * topk + k] = best;
      indices[(size_t)row * topk + k] = bi;
      source_rows[(size_t)row * topk + k] = k * num_tokens + row;
      p[bi] = -1.f;  // exclude from the next rounds
    }
    __syncwarp();
  }
}

// final[t, :] (+)= T(cur[t, :] * w_e[t]),  w_e[t] = sum_k topk_weights[t, k] * (topk_ids[t, k] == expert)
// The expert loop body of the reference's dense-per-expert MoE (aphrodite/modeling/models/mixtral_quant.py:141-152:
// expert_mask -> expert_weights -> current.mul_(expert_weights) -> final.add_(current)) as one pass: the product is
// formed in fp32 and rounded to T (in-place mul_ of a 16-bit tensor by an fp32 one), the accumulation is a T add.
template <typename T>
__global__ void __launch_bounds__(256)
moe_expert_scale_add_kernel(T* __restrict__ final_out, const T* __restrict__ cur, const float* __restrict__ topk_weights,
                            const int32_t* __restrict__ topk_ids, int num_tokens, int hidden, int topk, int expert,
                            int first) {
  constexpr int N = 16 / sizeof(T);
  const int per_tok = hidden / N;
  const int64_t total = (int64_t)num_tokens * per_tok;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = idx / per_tok;
    float w = 0.f;
    for (int k = 0; k < topk; ++k)
      w += (topk_ids[tok * topk + k] == expert) ? topk_weights[tok * topk + k] : 0.f;
    union { uint4 raw; T e[N]; } c, f;
    c.raw = __ldg(reinterpret_cast<const uint4*>(cur) + idx);
    if (!first) f.raw = reinterpret_cast<const uint4*>(final_out)[idx];
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const T scaled = from_f32<T>(__fmul_rn(to_f32<T>(c.e[e]), w));
      f.e[e] = first ? scaled : from_f32<T>(__fadd_rn(to_f32<T>(f.e[e]), to_f32<T>(scaled)));
    }
    reinterpret_cast<uint4*>(final_out)[idx] = f.raw;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_moe_align_block_size(const void* topk_ids, int ids_are_int64, int64_t numel,
                                         int num_experts, int block_size, int32_t* sorted_token_ids,
                                         int32_t* expert_ids, int32_t* num_tokens_post_pad, void* stream) {
  B200_CHECK(num_experts > 0 && block_size > 0, "moe_align_block_size: bad num_experts / block_size");
  int T = 256;
  while (T > 1 && ((size_t)(T + 1) * num_experts + num_experts + 1) * 4 > 200 * 1024) T >>= 1;
  const size_t smem = ((size_t)(T + 1) * num_experts + num_experts + 1) * 4;
  B200_CHECK(smem <= 220 * 1024, "moe_align_block_size: too many experts for shared memory");
  cudaStream_t st = (cudaStream_t)stream;
  if (ids_are_int64) {
    auto k = moe_align_kernel<int64_t>;
    B200_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<1, 256, smem, st>>>((const int64_t*)topk_ids, sorted_token_ids, expert_ids, num_tokens_post_pad,
                            num_experts, block_size, numel, T);
  } else {
    auto k = moe_align_kernel<int32_t>;
    B200_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<1, 256, smem, st>>>((const int32_t*)topk_ids, sorted_token_ids, expert_ids, num_tokens_post_pad,
                            num_experts, block_size, numel, T);
  }
  return check_launch("moe_align_kernel");
}

extern "C" int b200_topk_softmax(float* topk_weights, int32_t* topk_indices, int32_t* token_expert_indices,
                                 const float* gating_output, int num_tokens, int num_experts, int topk,
                                 void* stream) {
  B200_CHECK(topk >= 1 && topk <= num_experts, "topk_softmax: topk must be in [1, num_experts]");
  if (num_tokens == 0) return 0;
  const int warps = 4;
  const size_t smem = (size_t)warps * num_experts * sizeof(float);
  B200_CHECK(smem <= 48 * 1024, "topk_softmax: too many experts");
  topk_softmax_kernel<<<(num_tokens + warps - 1) / warps, warps * 32, smem, (cudaStream_t)stream>>>(
      gating_output, topk_weights, topk_indices, token_expert_indices, num_tokens, num_experts, topk);
  return check_launch("topk_softmax_kernel");
}

extern "C" int b200_moe_expert_scale_add(void* final_out, const void* cur, const float* topk_weights,
                                         const int32_t* topk_ids, int num_tokens, int hidden, int topk, int expert,
                                         int first, int dtype, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "moe_expert_scale_add: float16 / bfloat16 only");
  B200_CHECK(hidden % 8 == 0 && ((reinterpret_cast<uintptr_t>(final_out) | reinterpret_cast<uintptr_t>(cur)) & 15) == 0,
             "moe_expert_scale_add: rows must be 16-byte aligned multiples of 8 elements");
  if (num_tokens == 0) return 0;
  const int64_t work = (int64_t)num_tokens * (hidden / 8);
  const int grid = (int)std::min<int64_t>((work + 255) / 256, (int64_t)num_sms() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_BF16)
    moe_expert_scale_add_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)final_out, (const __nv_bfloat16*)cur,
                                                                    topk_weights, topk_ids, num_tokens, hidden, topk,
                                                                    expert, first);
  else
    moe_expert_scale_add_kernel<__half><<<grid, 256, 0, st>>>((__half*)final_out, (const __half*)cur, topk_weights,
                                                             topk_ids, num_tokens, hidden, topk, expert, first);
  return check_launch("moe_expert_scale_add_kernel");
}
