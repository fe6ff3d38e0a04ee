// sampling.cu — the sampler's probability-space kernels, sm_100a.
//
// Replaces kernels/sampling/sampling.cu + sampling.cuh of the reference (FlashInfer-derived):
//   sampling_from_probs              sampling.cu:43-73    kernel SamplingFromProbKernel           sampling.cuh:262-298
//   top_k_sampling_from_probs        sampling.cu:125-172  kernel TopKSamplingFromProbKernel       sampling.cuh:300-385
//   top_p_sampling_from_probs        sampling.cu:75-123   kernel TopPSamplingFromProbKernel       sampling.cuh:387-473
//   min_p_sampling_from_probs        sampling.cu:174-220  kernel MinPSamplingFromProbKernel       sampling.cuh:475-594
//   top_k_top_p_sampling_from_probs  sampling.cu:222-280  kernel TopKTopPSamplingFromProbKernel   sampling.cuh:596-690
//   top_p_renorm_prob / top_k_renorm_prob / top_k_mask_logits   sampling.cu:282-390, sampling.cuh:909-1310
//
// Contract kept: fp32 rows [batch, vocab]; the rejection samplers consume uniform_samples[round, row] in the
// reference's order (round r draws u = uniform[r, row] * q with q = mass above the running pivot, samples by inverse
// CDF over the entries above the pivot, raises the pivot to the sampled probability and stops as soon as the entries
// above the pivot satisfy the filter), so the same uniforms give the same token; `success` = the filter was met within
// the given rounds. The renorm / mask kernels return the reference's fixed point: keep x >= t*, t* = the largest value
// with count(x >= t*) >= k resp. sum(x >= t*) >= p, found by bisection on the value range with the reference's
// termination rule (no element strictly between the bounds).
//
// Not a port: no CUB; one 1024-thread CTA per row, 16-byte row loads, block scans / reductions on warp shuffles with a
// FIXED combination order (every result is deterministic, the `deterministic` flag of the reference is accepted and
// has nothing left to choose), the (sum, count) reductions fused in one pass. The rows are re-read once per pass: a
// 128 256-entry row is 513 KB, the ~148 rows in flight stay L2-resident (76 MB of 126 MB) after the first, HBM-bound,
// pass. Algorithmic bytes = batch * vocab * 4 (+ the same again for the ops that write a row).
#include "common.cuh"

#include <math.h>

namespace b200 {

static constexpr int SP_THREADS = 1024;
static constexpr int SP_WARPS = SP_THREADS / 32;
static constexpr int SP_CHUNK = SP_THREADS * 4;

struct SpShared {
  float wf[SP_WARPS];
  float wf2[SP_WARPS];
  int wi[SP_WARPS];
  float bf, bf2;
  int bi;
  int sampled_id;
};

// four consecutive row entries starting at `i0` (zeros / `fill` beyond the row)
__device__ __forceinline__ void load4(const float* __restrict__ row, int i0, int V, bool vec, float fill, float (&x)[4]) {
  if (vec && i0 + 3 < V) {
    const float4 v = *reinterpret_cast<const float4*>(row + i0);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = (i0 + j < V) ? row[i0 + j] : fill;
  }
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block reductions, result broadcast to every thread; fixed order (lane tree, then warp 0 over the warp results)
__device__ __forceinline__ float block_sum(float v, SpShared& sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh.wf[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(sh.wf[threadIdx.x]);
    if (threadIdx.x == 0) sh.bf = t;
  }
  __syncthreads();
  const float r = sh.bf;
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_max(float v, SpShared& sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh.wf[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_max(sh.wf[threadIdx.x]);
    if (threadIdx.x == 0) sh.bf = t;
  }
  __syncthreads();
  const float r = sh.bf;
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_min(float v, SpShared& sh) {
  v = warp_min(v);
  if ((threadIdx.x & 31) == 0) sh.wf[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_min(sh.wf[threadIdx.x]);
    if (threadIdx.x == 0) sh.bf = t;
  }
  __syncthreads();
  const float r = sh.bf;
  __syncthreads();
  return r;
}
// (sum, count) in one round trip
__device__ __forceinline__ void block_sum_count(float& s, int& c, SpShared& sh) {
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) { sh.wf[threadIdx.x >> 5] = s; sh.wi[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float ts = warp_sum(sh.wf[threadIdx.x]);
    int tc = sh.wi[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tc += __shfl_xor_sync(0xffffffffu, tc, o);
    if (threadIdx.x == 0) { sh.bf = ts; sh.bi = tc; }
  }
  __syncthreads();
  s = sh.bf;
  c = sh.bi;
  __syncthreads();
}

// Inverse-CDF sample over the entries of the row that are > pivot: the first index i (entry > pivot) whose inclusive
// running sum exceeds u; vocab - 1 if the sum never does (sampling.cuh:186-260 DeviceSamplingFromProb + the callers'
// loops). Every thread returns the id.
__device__ int sample_above(const float* __restrict__ row, int V, bool vec, float pivot, float u, SpShared& sh) {
  // 16 consecutive entries per thread and round: one block scan (3 barriers) per 16 384 entries, and four 16-byte loads
  // in flight per thread (the first version scanned 4 096 entries per round: 0.55 of the HBM peak on the first pass)
  constexpr int EPT = 16;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) sh.sampled_id = V - 1;
  __syncthreads();
  float agg = 0.f;
  for (int base = 0; base < V; base += SP_THREADS * EPT) {
    const int i0 = base + tid * EPT;
    float x[EPT];
#pragma unroll
    for (int k = 0; k < EPT / 4; ++k) {
      float t[4];
      load4(row, i0 + 4 * k, V, vec, 0.f, t);
      x[4 * k] = t[0]; x[4 * k + 1] = t[1]; x[4 * k + 2] = t[2]; x[4 * k + 3] = t[3];
    }
    float q4[EPT / 4];
#pragma unroll
    for (int k = 0; k < EPT / 4; ++k) {
      const float y0 = x[4 * k] > pivot ? x[4 * k] : 0.f, y1 = x[4 * k + 1] > pivot ? x[4 * k + 1] : 0.f;
      const float y2 = x[4 * k + 2] > pivot ? x[4 * k + 2] : 0.f, y3 = x[4 * k + 3] > pivot ? x[4 * k + 3] : 0.f;
      q4[k] = (y0 + y1) + (y2 + y3);
    }
    const float tsum = (q4[0] + q4[1]) + (q4[2] + q4[3]);
    // block exclusive scan of the per-thread sums
    float incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) sh.wf[w] = incl;
    __syncthreads();
    if (w == 0) {
      const float wt = sh.wf[lane];
      float wi = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      sh.wf2[lane] = wi - wt;                 // exclusive prefix of warp `lane`
      if (lane == 31) sh.bf = wi;             // chunk total
    }
    __syncthreads();
    const float total = sh.bf;
    const float excl = sh.wf2[w] + (incl - tsum);
    const bool hit = agg + total > u;         // uniform over the CTA
    if (hit) {
      float c = agg + excl;
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const bool in = x[j] > pivot;
        c += in ? x[j] : 0.f;
        if (c > u && in && i0 + j < V) {
          atomicMin(&sh.sampled_id, i0 + j);
          break;
        }
      }
    }
    __syncthreads();                          // sampled_id final / wf, wf2, bf reusable
    if (hit) break;
    agg += total;
  }
  const int id = sh.sampled_id;
  __syncthreads();
  return id;
}

// (sum, count) of the entries > pivot
__device__ void mass_above(const float* __restrict__ row, int V, bool vec, float pivot, float& sum, int& cnt, SpShared& sh) {
  float s = 0.f;
  int c = 0;
  for (int base = 0; base < V; base += SP_CHUNK) {
    const int i0 = base + threadIdx.x * 4;
    float x[4];
    load4(row, i0, V, vec, 0.f, x);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x[j] > pivot && i0 + j < V) { s += x[j]; ++c; }
  }
  block_sum_count(s, c, sh);
  sum = s;
  cnt = c;
}

__device__ __forceinline__ bool row_vec_ok(const float* row, int V) {
  return (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
}

__global__ void __launch_bounds__(SP_THREADS) sampling_from_probs_kernel(const float* __restrict__ probs,
                                                                         const float* __restrict__ uniform,
                                                                         int* __restrict__ out, int V) {
  __shared__ SpShared sh;
  const float* row = probs + (size_t)blockIdx.x * V;
  const int id = sample_above(row, V, row_vec_ok(row, V), 0.f, uniform[blockIdx.x], sh);
  if (threadIdx.x == 0) out[blockIdx.x] = id;
}

enum { SP_TOPK = 0, SP_TOPP = 1, SP_MINP = 2, SP_TOPK_TOPP = 3 };

template <int MODE>
__global__ void __launch_bounds__(SP_THREADS) rejection_sampling_kernel(
    const float* __restrict__ probs, const float* __restrict__ uniform, int* __restrict__ out, bool* __restrict__ success,
    const int* __restrict__ k_arr, int k_val, const float* __restrict__ p_arr, float p_val, int V, int rounds) {
  __shared__ SpShared sh;
  const int B = gridDim.x, b = blockIdx.x;
  const float* row = probs + (size_t)b * V;
  const bool vec = row_vec_ok(row, V);
  const unsigned k = (unsigned)(k_arr ? k_arr[b] : k_val);        // compared as unsigned like the reference (uint32_t k)
  const float p = p_arr ? p_arr[b] : p_val;
  float q = 1.f, pivot = 0.f, scaled_p = 0.f;
  int cnt = 0, id = V - 1;
  bool ok = false;
  if (MODE == SP_MINP) {
    float m = 0.f;
    for (int base = 0; base < V; base += SP_CHUNK) {
      float x[4];
      load4(row, base + threadIdx.x * 4, V, vec, 0.f, x);
      m = fmaxf(fmaxf(m, fmaxf(x[0], x[1])), fmaxf(x[2], x[3]));
    }
    scaled_p = block_max(m, sh) * p;
  }
  for (int r = 0; r < rounds; ++r) {
    const float u = uniform[(size_t)r * B + b] * q;
    id = sample_above(row, V, vec, pivot, u, sh);
    pivot = fmaxf(pivot, row[id]);
    if (MODE == SP_MINP) {
      if (pivot >= scaled_p) { ok = true; break; }
    }
    mass_above(row, V, vec, pivot, q, cnt, sh);
    if (MODE == SP_TOPK) { if ((unsigned)cnt < k) { ok = true; break; } }
    if (MODE == SP_TOPP) { if (q < p) { ok = true; break; } }
    if (MODE == SP_TOPK_TOPP) { if ((unsigned)cnt < k && q < p) { ok = true; break; } }
  }
  if (threadIdx.x == 0) {
    out[b] = id;
    if (success != nullptr) success[b] = ok;
  }
}

// ---- renormalisation / masking --------------------------------------------------------------------------------------
// Bisection of the value range for the kept set {x > low}: invariant f(low) >= target > f(high) with f(t) = count or mass
// of the entries > t; stops when no entry lies strictly between the bounds (min{x > low} == max{x <= high}), the
// reference's loop (sampling.cuh:955-1010, 1085-1140, 1215-1275). MODE 0: count >= k, 1: mass >= p.
template <int MODE>
__device__ void bisect_pivot(const float* __restrict__ row, int V, bool vec, float lo0, float hi0, unsigned k, float p,
                             float& low_out, float& sum_low_out, SpShared& sh) {
  float low = lo0, high = hi0, sum_low = 1.f;
  float min_gt_low, max_le_high;
  do {
    const float mid = (low + high) / 2;
    float s = 0.f;
    int c = 0;
    float mn = high, mx = low;
    for (int base = 0; base < V; base += SP_CHUNK) {
      const int i0 = base + threadIdx.x * 4;
      float x[4];
      load4(row, i0, V, vec, 0.f, x);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (i0 + j < V) {
          if (x[j] > mid) { s += x[j]; ++c; }
          if (x[j] > low) mn = fminf(mn, x[j]);
          if (x[j] <= high) mx = fmaxf(mx, x[j]);
        }
      }
    }
    block_sum_count(s, c, sh);
    min_gt_low = block_min(mn, sh);
    max_le_high = block_max(mx, sh);
    const bool ge = MODE == 0 ? ((unsigned)c >= k) : (s >= p);
    if (ge) {
      low = mid;
      sum_low = s;
    } else {
      high = fminf(mid, max_le_high);
    }
  } while (min_gt_low != max_le_high);
  low_out = low;
  sum_low_out = sum_low;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// out = x > pivot ? x * normalizer : 0   (MASK: x > pivot ? x : -inf)
template <bool MASK>
__device__ void write_row(const float* __restrict__ row, float* __restrict__ orow, int V, bool vec, float pivot, float norm) {
  const float dead = MASK ? -INFINITY : 0.f;
  const bool ovec = vec && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
  for (int base = 0; base < V; base += SP_CHUNK) {
    const int i0 = base + threadIdx.x * 4;
    if (i0 >= V) break;
    float x[4];
    load4(row, i0, V, vec, 0.f, x);
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = x[j] > pivot ? (MASK ? x[j] : x[j] * norm) : dead;
    if (ovec && i0 + 3 < V) {
      *reinterpret_cast<float4*>(orow + i0) = make_float4(x[0], x[1], x[2], x[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j < V) orow[i0 + j] = x[j];
    }
  }
}

__device__ float row_max(const float* __restrict__ row, int V, bool vec, float init, SpShared& sh) {
  float m = init;
  for (int base = 0; base < V; base += SP_CHUNK) {
    float x[4];
    load4(row, base + threadIdx.x * 4, V, vec, init, x);
    m = fmaxf(fmaxf(m, fmaxf(x[0], x[1])), fmaxf(x[2], x[3]));
  }
  return block_max(m, sh);
}

__global__ void __launch_bounds__(SP_THREADS) top_p_renorm_kernel(const float* __restrict__ probs, float* __restrict__ out,
                                                                  const float* __restrict__ p_arr, float p_val, int V) {
  __shared__ SpShared sh;
  const float* row = probs + (size_t)blockIdx.x * V;
  const bool vec = row_vec_ok(row, V);
  const float p = p_arr ? p_arr[blockIdx.x] : p_val;
  const float mx = row_max(row, V, vec, 0.f, sh);
  float low, sum_low;
  bisect_pivot<1>(row, V, vec, 0.f, mx, 0u, p, low, sum_low, sh);
  write_row<false>(row, out + (size_t)blockIdx.x * V, V, vec, low, rcp_approx(fmaxf(sum_low, 1e-8f)));
}

__global__ void __launch_bounds__(SP_THREADS) top_k_renorm_kernel(const float* __restrict__ probs, float* __restrict__ out,
                                                                  const int* __restrict__ k_arr, int k_val, int V) {
  __shared__ SpShared sh;
  const float* row = probs + (size_t)blockIdx.x * V;
  const bool vec = row_vec_ok(row, V);
  const unsigned k = (unsigned)(k_arr ? k_arr[blockIdx.x] : k_val);
  float pivot = -INFINITY, norm = 1.f;
  if (k < (unsigned)V) {
    const float mx = row_max(row, V, vec, 0.f, sh);
    float sum_low;
    bisect_pivot<0>(row, V, vec, 0.f, mx, k, 0.f, pivot, sum_low, sh);
    norm = rcp_approx(fmaxf(sum_low, 1e-8f));
  }
  write_row<false>(row, out + (size_t)blockIdx.x * V, V, vec, pivot, norm);
}

__global__ void __launch_bounds__(SP_THREADS) top_k_mask_logits_kernel(const float* __restrict__ logits, float* __restrict__ out,
                                                                       const int* __restrict__ k_arr, int k_val, int V) {
  __shared__ SpShared sh;
  const float* row = logits + (size_t)blockIdx.x * V;
  const bool vec = row_vec_ok(row, V);
  const unsigned k = (unsigned)(k_arr ? k_arr[blockIdx.x] : k_val);
  float pivot = -INFINITY;
  if (k < (unsigned)V) {
    const float mx = row_max(row, V, vec, -INFINITY, sh);
    float mn = INFINITY;
    for (int base = 0; base < V; base += SP_CHUNK) {
      float x[4];
      load4(row, base + threadIdx.x * 4, V, vec, INFINITY, x);
      mn = fminf(fminf(mn, fminf(x[0], x[1])), fminf(x[2], x[3]));
    }
    mn = block_min(mn, sh);
    float sum_low;
    bisect_pivot<0>(row, V, vec, mn - 1.f, mx, k, 0.f, pivot, sum_low, sh);
  }
  write_row<true>(row, out + (size_t)blockIdx.x * V, V, vec, pivot, 1.f);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sampling_from_probs(const float* probs, const float* uniform_samples, int32_t* samples,
                                        int batch_size, int vocab_size, int deterministic, void* stream) {
  (void)deterministic;
  B200_CHECK(probs && uniform_samples && samples, "sampling_from_probs: null pointer");
  B200_CHECK(vocab_size > 0, "sampling_from_probs: empty vocabulary");
  if (batch_size == 0) return 0;
  sampling_from_probs_kernel<<<batch_size, SP_THREADS, 0, (cudaStream_t)stream>>>(probs, uniform_samples, samples, vocab_size);
  return check_launch("sampling_from_probs_kernel");
}

extern "C" int b200_rejection_sampling_from_probs(int mode, const float* probs, const float* uniform_samples,
                                                  int32_t* samples, uint8_t* success, const int32_t* top_k_arr,
                                                  int top_k_val, const float* top_p_arr, float top_p_val, int batch_size,
                                                  int vocab_size, int max_rounds, int deterministic, void* stream) {
  (void)deterministic;
  B200_CHECK(probs && uniform_samples && samples, "sampling: null pointer");
  B200_CHECK(vocab_size > 0 && max_rounds > 0, "sampling: empty vocabulary or no rounds");
  if (batch_size == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  bool* ok = reinterpret_cast<bool*>(success);
#define B200_SP(M) rejection_sampling_kernel<M><<<batch_size, SP_THREADS, 0, st>>>(probs, uniform_samples, samples, ok, \
                                               top_k_arr, top_k_val, top_p_arr, top_p_val, vocab_size, max_rounds)
  switch (mode) {
    case B200_SAMPLE_TOP_K: B200_SP(SP_TOPK); break;
    case B200_SAMPLE_TOP_P: B200_SP(SP_TOPP); break;
    case B200_SAMPLE_MIN_P: B200_SP(SP_MINP); break;
    case B200_SAMPLE_TOP_K_TOP_P: B200_SP(SP_TOPK_TOPP); break;
    default: return fail("sampling: unknown mode");
  }
#undef B200_SP
  return check_launch("rejection_sampling_kernel");
}

extern "C" int b200_top_p_renorm_prob(const float* probs, float* renorm_probs, const float* top_p_arr, float top_p_val,
                                      int batch_size, int vocab_size, void* stream) {
  B200_CHECK(probs && renorm_probs && vocab_size > 0, "top_p_renorm_prob: bad arguments");
  if (batch_size == 0) return 0;
  top_p_renorm_kernel<<<batch_size, SP_THREADS, 0, (cudaStream_t)stream>>>(probs, renorm_probs, top_p_arr, top_p_val, vocab_size);
  return check_launch("top_p_renorm_kernel");
}

extern "C" int b200_top_k_renorm_prob(const float* probs, float* renorm_probs, const int32_t* top_k_arr, int top_k_val,
                                      int batch_size, int vocab_size, void* stream) {
  B200_CHECK(probs && renorm_probs && vocab_size > 0, "top_k_renorm_prob: bad arguments");
  if (batch_size == 0) return 0;
  top_k_renorm_kernel<<<batch_size, SP_THREADS, 0, (cudaStream_t)stream>>>(probs, renorm_probs, top_k_arr, top_k_val, vocab_size);
  return check_launch("top_k_renorm_kernel");
}

extern "C" int b200_top_k_mask_logits(const float* logits, float* masked_logits, const int32_t* top_k_arr, int top_k_val,
                                      int batch_size, int vocab_size, void* stream) {
  B200_CHECK(logits && masked_logits && vocab_size > 0, "top_k_mask_logits: bad arguments");
  if (batch_size == 0) return 0;
  top_k_mask_logits_kernel<<<batch_size, SP_THREADS, 0, (cudaStream_t)stream>>>(logits, masked_logits, top_k_arr, top_k_val, vocab_size);
  return check_launch("top_k_mask_logits_kernel");
}
