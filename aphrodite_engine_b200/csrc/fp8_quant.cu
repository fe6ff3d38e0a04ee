// fp8_quant.cu — activation quantisation to fp8-e4m3 ahead of the W8A8 GEMM (scaled_mm.cu), sm_100a.
//
// Replaces kernels/quantization/fp8/common.cu of the reference:
//   static_scaled_fp8_quant             :260-277  (kernel scaled_fp8_quant_kernel :182-194)
//   dynamic_scaled_fp8_quant            :279-299  (segmented_max_reduction :71-105 + the same quant kernel)
//   dynamic_per_token_scaled_fp8_quant  :301-321  (kernel :196-256)
// Numerics kept (results are bit-identical to the reference kernels, tests/test_gpu_fp8_quant.py):
//   * per-tensor: the element is MULTIPLIED by 1.0f / scale (one IEEE division per thread), clamped to +-448, rounded
//     to e4m3 (round-to-nearest-even); NaN clamps to +448 through fminf/fmaxf like the reference (:55)
//   * per-token: scale = max(min(absmax, ub) / 448, 1 / (448 * 512)), the element is DIVIDED by the scale (:246-255,
//     "so we can match FBGemm")
//   * dynamic per-tensor: the scale tensor must be <= 0 on entry (the caller zeroes it, _custom_ops.py:scaled_fp8_quant);
//     absmax / 448 is folded in with an atomic max on the non-negative float's bit pattern (:33-43)
// HBM-bound byte movers: algorithmic bytes = numel * (sizeof(T) [+ sizeof(T) for a dynamic scale's first pass] + 1).
// 16-byte loads / 8-byte stores when the row length and the pointers allow, scalar otherwise.
#include "common.cuh"

#include <algorithm>

namespace b200 {

static constexpr float FP8_E4M3_MAX = 448.0f;

__device__ __forceinline__ uint8_t to_e4m3_mul(float v, float inv_scale) {
  const float x = v * inv_scale;
  const float r = fmaxf(-FP8_E4M3_MAX, fminf(x, FP8_E4M3_MAX));
  return (uint8_t)__nv_cvt_float_to_fp8(r, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint8_t to_e4m3_div(float v, float scale) {
  const float x = v / scale;
  const float r = fmaxf(-FP8_E4M3_MAX, fminf(x, FP8_E4M3_MAX));
  return (uint8_t)__nv_cvt_float_to_fp8(r, __NV_SATFINITE, __NV_E4M3);
}

template <typename T> struct Vec8 {            // 8 elements: 16 bytes of a 16-bit type, 32 bytes of float
  T v[8];
};
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]) {
  if constexpr (sizeof(T) == 2) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = to_f32<T>(e[i]);
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}

template <typename T> __device__ __forceinline__ bool vec_ok(const T* in, const uint8_t* out, int64_t n) {
  return (n % 8 == 0) && ((reinterpret_cast<uintptr_t>(in) & (sizeof(T) * 8 - 1)) == 0) &&
         ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
}

// out[i] = e4m3(in[i] * (1 / *scale)), flat over numel
template <typename T>
__global__ void __launch_bounds__(512) fp8_quant_tensor_kernel(uint8_t* __restrict__ out, const T* __restrict__ in,
                                                               const float* __restrict__ scale, int64_t numel) {
  const float inv = 1.0f / (*scale);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(in, out, numel)) {
    // two 16-byte loads in flight per thread and iteration
    for (int64_t i = tid * 8; i < numel; i += step * 16) {
      const int64_t i2 = i + step * 8;
      const bool two = i2 < numel;
      float f[8], g[8];
      load8(in + i, f);
      if (two) load8(in + i2, g);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo |= (uint32_t)to_e4m3_mul(f[j], inv) << (8 * j);
        hi |= (uint32_t)to_e4m3_mul(f[4 + j], inv) << (8 * j);
      }
      *reinterpret_cast<uint2*>(out + i) = make_uint2(lo, hi);
      if (two) {
        lo = 0; hi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          lo |= (uint32_t)to_e4m3_mul(g[j], inv) << (8 * j);
          hi |= (uint32_t)to_e4m3_mul(g[4 + j], inv) << (8 * j);
        }
        *reinterpret_cast<uint2*>(out + i2) = make_uint2(lo, hi);
      }
    }
  } else {
    for (int64_t i = tid; i < numel; i += step) out[i] = to_e4m3_mul(to_f32<T>(in[i]), inv);
  }
}

// *scale = max(*scale, absmax(in) / 448); *scale must be <= 0 on entry
template <typename T>
__global__ void __launch_bounds__(512) fp8_absmax_kernel(float* __restrict__ scale, const T* __restrict__ in,
                                                         int64_t numel) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  float m = 0.f;
  if ((numel % 8 == 0) && ((reinterpret_cast<uintptr_t>(in) & (sizeof(T) * 8 - 1)) == 0)) {
    for (int64_t i = tid * 8; i < numel; i += step * 8) {
      float f[8];
      load8(in + i, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(f[j]));
    }
  } else {
    for (int64_t i = tid; i < numel; i += step) m = fmaxf(m, fabsf(to_f32<T>(in[i])));
  }
  m = warp_max(m);
  __shared__ float wmax[16];
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? wmax[threadIdx.x] : 0.f;
    m = warp_max(m);
    // non-negative floats order like their bit patterns
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(scale), __float_as_int(m / FP8_E4M3_MAX));
  }
}

// one CTA per token: scales[token] = max(min(absmax, *ub) / 448, 1 / (448 * 512)); out = e4m3(in / scale)
template <typename T>
__global__ void __launch_bounds__(1024) fp8_quant_token_kernel(uint8_t* __restrict__ out, float* __restrict__ scales,
                                                               const T* __restrict__ in, const float* __restrict__ scale_ub,
                                                               int hidden) {
  const T* row = in + (int64_t)blockIdx.x * hidden;
  uint8_t* orow = out + (int64_t)blockIdx.x * hidden;
  const bool vec = vec_ok(row, orow, hidden);
  // rows of up to 8 elements per thread (hidden <= 8192 at 1024 threads) stay in registers between the absmax pass and
  // the conversion: the row crosses HBM / L2 once
  const bool keep = vec && hidden <= (int)blockDim.x * 8;
  float kept[8];
  const bool mine = keep && (int)threadIdx.x * 8 < hidden;
  float m = 0.f;
  if (keep) {
    if (mine) {
      load8(row + threadIdx.x * 8, kept);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(kept[j]));
    }
  } else if (vec) {
    for (int i = threadIdx.x * 8; i < hidden; i += blockDim.x * 8) {
      float f[8];
      load8(row + i, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(f[j]));
    }
  } else {
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) m = fmaxf(m, fabsf(to_f32<T>(row[i])));
  }
  m = warp_max(m);
  __shared__ float wmax[32];
  __shared__ float s_scale;
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < ((blockDim.x + 31) >> 5) ? wmax[threadIdx.x] : 0.f;
    m = warp_max(m);
    if (threadIdx.x == 0) {
      float s = scale_ub ? fminf(m, *scale_ub) : m;
      s = fmaxf(s / FP8_E4M3_MAX, 1.0f / (FP8_E4M3_MAX * 512.f));
      scales[blockIdx.x] = s;
      s_scale = s;
    }
  }
  __syncthreads();
  const float s = s_scale;
  if (keep) {
    if (mine) {
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo |= (uint32_t)to_e4m3_div(kept[j], s) << (8 * j);
        hi |= (uint32_t)to_e4m3_div(kept[4 + j], s) << (8 * j);
      }
      *reinterpret_cast<uint2*>(orow + threadIdx.x * 8) = make_uint2(lo, hi);
    }
  } else if (vec) {
    for (int i = threadIdx.x * 8; i < hidden; i += blockDim.x * 8) {     // second read of the row: L1 / L2 resident
      float f[8];
      load8(row + i, f);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lo |= (uint32_t)to_e4m3_div(f[j], s) << (8 * j);
        hi |= (uint32_t)to_e4m3_div(f[4 + j], s) << (8 * j);
      }
      *reinterpret_cast<uint2*>(orow + i) = make_uint2(lo, hi);
    }
  } else {
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) orow[i] = to_e4m3_div(to_f32<T>(row[i]), s);
  }
}

template <typename T>
static int run_tensor_quant(void* out, const void* in, float* scale, int64_t numel, bool dynamic, cudaStream_t st) {
  if (numel == 0) return 0;
  const int threads = 512;
  const int64_t want = (numel / 8 + threads - 1) / threads;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms() * 8));
  if (dynamic) {
    fp8_absmax_kernel<T><<<grid, threads, 0, st>>>(scale, (const T*)in, numel);
    if (int rc = check_launch("fp8_absmax_kernel")) return rc;
  }
  fp8_quant_tensor_kernel<T><<<grid, threads, 0, st>>>((uint8_t*)out, (const T*)in, scale, numel);
  return check_launch("fp8_quant_tensor_kernel");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_static_scaled_fp8_quant(void* out, const void* input, const float* scale, int64_t numel, int dtype,
                                            void* stream) {
  B200_CHECK(out != nullptr && input != nullptr && scale != nullptr, "static_scaled_fp8_quant: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B200_F32: return run_tensor_quant<float>(out, input, const_cast<float*>(scale), numel, false, st);
    case B200_F16: return run_tensor_quant<__half>(out, input, const_cast<float*>(scale), numel, false, st);
    case B200_BF16: return run_tensor_quant<__nv_bfloat16>(out, input, const_cast<float*>(scale), numel, false, st);
  }
  return fail("static_scaled_fp8_quant: unsupported input dtype");
}

extern "C" int b200_dynamic_scaled_fp8_quant(void* out, const void* input, float* scale, int64_t numel, int dtype,
                                             void* stream) {
  B200_CHECK(out != nullptr && input != nullptr && scale != nullptr, "dynamic_scaled_fp8_quant: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case B200_F32: return run_tensor_quant<float>(out, input, scale, numel, true, st);
    case B200_F16: return run_tensor_quant<__half>(out, input, scale, numel, true, st);
    case B200_BF16: return run_tensor_quant<__nv_bfloat16>(out, input, scale, numel, true, st);
  }
  return fail("dynamic_scaled_fp8_quant: unsupported input dtype");
}

extern "C" int b200_dynamic_per_token_scaled_fp8_quant(void* out, const void* input, float* scales,
                                                       const float* scale_ub, int num_tokens, int hidden_size,
                                                       int dtype, void* stream) {
  B200_CHECK(out != nullptr && input != nullptr && scales != nullptr, "dynamic_per_token_scaled_fp8_quant: null pointer");
  if (num_tokens == 0 || hidden_size == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int threads = std::min(1024, std::max(32, (hidden_size / 8 + 31) / 32 * 32));
  if (hidden_size % 8 != 0) threads = std::min(1024, (hidden_size + 31) / 32 * 32);
  switch (dtype) {
    case B200_F32:
      fp8_quant_token_kernel<float><<<num_tokens, threads, 0, st>>>((uint8_t*)out, scales, (const float*)input, scale_ub, hidden_size);
      break;
    case B200_F16:
      fp8_quant_token_kernel<__half><<<num_tokens, threads, 0, st>>>((uint8_t*)out, scales, (const __half*)input, scale_ub, hidden_size);
      break;
    case B200_BF16:
      fp8_quant_token_kernel<__nv_bfloat16><<<num_tokens, threads, 0, st>>>((uint8_t*)out, scales, (const __nv_bfloat16*)input, scale_ub, hidden_size);
      break;
    default:
      return fail("dynamic_per_token_scaled_fp8_quant: unsupported input dtype");
  }
  return check_launch("fp8_quant_token_kernel");
}
