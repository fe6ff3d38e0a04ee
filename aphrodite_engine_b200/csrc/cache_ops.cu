// cache_ops.cu — paged KV-cache writers / movers, sm_100a.
//
// Replaces kernels/cache_kernels.cu of the reference: reshape_and_cache (:152-204, :263-289),
// reshape_and_cache_flash (:206-245, :304-330), copy_blocks (:67-148), swap_blocks (:24-63),
// convert_fp8 (:334-410). All integer/index work is bit-exact with the reference by construction:
// the destination index formulas are the layout definition
//   key_cache  [num_blocks, num_heads, head_size/x, block_size, x]
//   value_cache[num_blocks, num_heads, head_size, block_size]
// These ops are HBM/latency-bound byte movers (1 MB per layer at 256 tokens): the B200 work is
// 16-byte vector accesses along the contiguous `x` run of K and wide grids, not tensor cores.
#include "common.cuh"

namespace b200 {

// one CTA per token; K is moved in x-element (16-byte) units, V element-wise (token-minor layout)
template <typename T, int KV>
__global__ void __launch_bounds__(256)
reshape_and_cache_kernel(const T* __restrict__ key, const T* __restrict__ value,
                         void* __restrict__ key_cache, void* __restrict__ value_cache,
                         const int64_t* __restrict__ slot_mapping, int64_t key_stride,
                         int64_t value_stride, int num_heads, int head_size, int block_size, int x,
                         float k_scale, float v_scale, bool vec_ok) {
  const int64_t token = blockIdx.x;
  const int64_t slot = slot_mapping[token];
  if (slot < 0) return;  // padding token
  const int64_t blk = slot / block_size;
  const int64_t boff = slot % block_size;
  const int n = num_heads * head_size;
  const T* ksrc = key + token * key_stride;
  const T* vsrc = value + token * value_stride;

  if constexpr (KV == B200_KV_AUTO) {
    T* kc = reinterpret_cast<T*>(key_cache);
    T* vc = reinterpret_cast<T*>(value_cache);
    if (vec_ok) {
      // x elements == 16 bytes: one uint4 per (head, x-chunk)
      const int units = n / x;
      const int cpx = head_size / x;
      for (int u = threadIdx.x; u < units; u += blockDim.x) {
        const int h = u / cpx, c = u % cpx;
        const uint4 v = *reinterpret_cast<const uint4*>(ksrc + (int64_t)u * x);
        const int64_t dst = (((blk * num_heads + h) * cpx + c) * block_size + boff) * x;
        *reinterpret_cast<uint4*>(kc + dst) = v;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int h = i / head_size, ho = i % head_size;
        const int64_t dst =
            (((blk * num_heads + h) * (head_size / x) + ho / x) * block_size + boff) * x + ho % x;
        kc[dst] = ksrc[i];
      }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int h = i / head_size, ho = i % head_size;
      const int64_t dst = ((blk * num_heads + h) * head_size + ho) * block_size + boff;
      vc[dst] = vsrc[i];
    }
  } else {
    uint8_t* kc = reinterpret_cast<uint8_t*>(key_cache);
    uint8_t* vc = reinterpret_cast<uint8_t*>(value_cache);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int h = i / head_size, ho = i % head_size;
      const int64_t kd =
          (((blk * num_heads + h) * (head_size / x) + ho / x) * block_size + boff) * x + ho % x;
      const int64_t vd = ((blk * num_heads + h) * head_size + ho) * block_size + boff;
      kc[kd] = fp8_quant<T, KV>(ksrc[i], k_scale);
      vc[vd] = fp8_quant<T, KV>(vsrc[i], v_scale);
    }
  }
}

// flash layout [num_blocks, block_size, num_heads, head_size]: a token's K (V) row is contiguous
template <typename T, int KV>
__global__ void __launch_bounds__(256)
reshape_and_cache_flash_kernel(const T* __restrict__ key, const T* __restrict__ value,
                               void* __restrict__ key_cache, void* __restrict__ value_cache,
                               const int64_t* __restrict__ slot_mapping, int64_t block_stride,
                               int64_t key_stride, int64_t value_stride, int num_heads,
                               int head_size, int block_size, float k_scale, float v_scale) {
  const int64_t token = blockIdx.x;
  const int64_t slot = slot_mapping[token];
  if (slot < 0) return;
  const int64_t blk = slot / block_size;
  const int64_t boff = slot % block_size;
  const int n = num_heads * head_size;
  const int64_t base = blk * block_stride + boff * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T k = key[token * key_stride + i];
    const T v = value[token * value_stride + i];
    if constexpr (KV == B200_KV_AUTO) {
      reinterpret_cast<T*>(key_cache)[base + i] = k;
      reinterpret_cast<T*>(value_cache)[base + i] = v;
    } else {
      reinterpret_cast<uint8_t*>(key_cache)[base + i] = fp8_quant<T, KV>(k, k_scale);
      reinterpret_cast<uint8_t*>(value_cache)[base + i] = fp8_quant<T, KV>(v, v_scale);
    }
  }
}

// grid (layers, pairs): byte-exact block copy inside each layer's K and V cache
__global__ void __launch_bounds__(256)
copy_blocks_kernel(const int64_t* __restrict__ key_ptrs, const int64_t* __restrict__ value_ptrs,
                   const int64_t* __restrict__ block_mapping, int64_t block_bytes) {
  const int layer = blockIdx.x, pair = blockIdx.y;
  const int64_t src = block_mapping[2 * pair], dst = block_mapping[2 * pair + 1];
  uint8_t* caches[2] = {reinterpret_cast<uint8_t*>(key_ptrs[layer]),
                        reinterpret_cast<uint8_t*>(value_ptrs[layer])};
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint8_t* s = caches[c] + src * block_bytes;
    uint8_t* d = caches[c] + dst * block_bytes;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)block_bytes) & 15) == 0) {
      const uint4* s4 = reinterpret_cast<const uint4*>(s);
      uint4* d4 = reinterpret_cast<uint4*>(d);
      for (int64_t i = threadIdx.x; i < block_bytes / 16; i += blockDim.x) d4[i] = s4[i];
    } else {
      for (int64_t i = threadIdx.x; i < block_bytes; i += blockDim.x) d[i] = s[i];
    }
  }
}

// fp8 <-> {f32,f16,bf16}; `TO_FP8` selects the direction
template <typename T, int KV, bool TO_FP8>
__global__ void __launch_bounds__(256)
convert_fp8_kernel(void* __restrict__ dst, const void* __restrict__ src, int64_t numel, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    if constexpr (TO_FP8) {
      reinterpret_cast<uint8_t*>(dst)[i] = fp8_quant<T, KV>(reinterpret_cast<const T*>(src)[i], scale);
    } else {
      reinterpret_cast<T*>(dst)[i] = fp8_dequant<T, KV>(reinterpret_cast<const uint8_t*>(src)[i], scale);
    }
  }
}

}  // namespace b200

using namespace b200;

#define B200_DISPATCH_T_KV(dtype, kv_dtype, FN)                                         \
  do {                                                                                  \
    if (dtype == B200_BF16) {                                                           \
      if (kv_dtype == B200_KV_AUTO) { FN(__nv_bfloat16, B200_KV_AUTO); }                \
      else if (kv_dtype == B200_KV_FP8_E4M3) { FN(__nv_bfloat16, B200_KV_FP8_E4M3); }   \
      else { FN(__nv_bfloat16, B200_KV_FP8_E5M2); }                                     \
    } else if (dtype == B200_F16) {                                                     \
      if (kv_dtype == B200_KV_AUTO) { FN(__half, B200_KV_AUTO); }                       \
      else if (kv_dtype == B200_KV_FP8_E4M3) { FN(__half, B200_KV_FP8_E4M3); }          \
      else { FN(__half, B200_KV_FP8_E5M2); }                                            \
    } else {                                                                            \
      if (kv_dtype == B200_KV_AUTO) { FN(float, B200_KV_AUTO); }                        \
      else if (kv_dtype == B200_KV_FP8_E4M3) { FN(float, B200_KV_FP8_E4M3); }           \
      else { FN(float, B200_KV_FP8_E5M2); }                                             \
    }                                                                                   \
  } while (0)

extern "C" int b200_reshape_and_cache(const void* key, const void* value, void* key_cache,
                                      void* value_cache, const int64_t* slot_mapping,
                                      int num_tokens, int num_heads, int head_size, int block_size,
                                      int x, int64_t key_stride, int64_t value_stride, int dtype,
                                      int kv_dtype, float k_scale, float v_scale, void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "Unsupported input type of kv cache");
  B200_CHECK(kv_dtype >= B200_KV_AUTO && kv_dtype <= B200_KV_FP8_E5M2,
             "Unsupported data type of kv cache");
  B200_CHECK(x > 0 && head_size % x == 0, "head_size must be a multiple of x");
  if (num_tokens == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = dtype == B200_F32 ? 4 : 2;
  const bool vec_ok = kv_dtype == B200_KV_AUTO && x * esz == 16 &&
                      (reinterpret_cast<uintptr_t>(key) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(key_cache) & 15) == 0 &&
                      (key_stride * esz) % 16 == 0;
#define FN(T, KVD)                                                                               \
  reshape_and_cache_kernel<T, KVD><<<num_tokens, 256, 0, st>>>(                                  \
      (const T*)key, (const T*)value, key_cache, value_cache, slot_mapping, key_stride,          \
      value_stride, num_heads, head_size, block_size, x, k_scale, v_scale, vec_ok)
  B200_DISPATCH_T_KV(dtype, kv_dtype, FN);
#undef FN
  return check_launch("reshape_and_cache_kernel");
}

extern "C" int b200_reshape_and_cache_flash(const void* key, const void* value, void* key_cache,
                                            void* value_cache, const int64_t* slot_mapping,
                                            int num_tokens, int num_heads, int head_size,
                                            int block_size, int64_t block_stride,
                                            int64_t key_stride, int64_t value_stride, int dtype,
                                            int kv_dtype, float k_scale, float v_scale,
                                            void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "Unsupported input type of kv cache");
  B200_CHECK(kv_dtype >= B200_KV_AUTO && kv_dtype <= B200_KV_FP8_E5M2,
             "Unsupported data type of kv cache");
  if (num_tokens == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define FN(T, KVD)                                                                         \
  reshape_and_cache_flash_kernel<T, KVD><<<num_tokens, 256, 0, st>>>(                      \
      (const T*)key, (const T*)value, key_cache, value_cache, slot_mapping, block_stride,  \
      key_stride, value_stride, num_heads, head_size, block_size, k_scale, v_scale)
  B200_DISPATCH_T_KV(dtype, kv_dtype, FN);
#undef FN
  return check_launch("reshape_and_cache_flash_kernel");
}

extern "C" int b200_copy_blocks(const int64_t* key_cache_ptrs, const int64_t* value_cache_ptrs,
                                const int64_t* block_mapping, int num_layers, int num_pairs,
                                int64_t block_bytes, void* stream) {
  if (num_layers == 0 || num_pairs == 0) return 0;
  dim3 grid(num_layers, num_pairs);
  copy_blocks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(key_cache_ptrs, value_cache_ptrs,
                                                            block_mapping, block_bytes);
  return check_launch("copy_blocks_kernel");
}

extern "C" int b200_swap_blocks(const void* src, void* dst, const int64_t* block_mapping_host,
                                int num_pairs, int64_t block_bytes, int kind, void* stream) {
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyDeviceToDevice
                               : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice);
  B200_CHECK(kind >= 0 && kind <= 2, "Invalid device combination");
  for (int i = 0; i < num_pairs; ++i) {
    const int64_t s = block_mapping_host[2 * i], d = block_mapping_host[2 * i + 1];
    B200_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(dst) + d * block_bytes,
                                 static_cast<const char*>(src) + s * block_bytes, block_bytes, k,
                                 (cudaStream_t)stream));
  }
  return 0;
}

extern "C" int b200_convert_fp8(void* dst, const void* src, int64_t numel, int src_dtype,
                                int dst_dtype, int kv_dtype, float scale, void* stream) {
  B200_CHECK((src_dtype < 0) != (dst_dtype < 0), "exactly one side of convert_fp8 must be fp8");
  B200_CHECK(kv_dtype == B200_KV_AUTO || kv_dtype == B200_KV_FP8_E4M3,
             "Unsupported data type: convert_fp8 handles auto/fp8/fp8_e4m3");
  if (numel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)((numel + 256 * 8 - 1) / (256 * 8) < 148 * 16
                               ? (numel + 256 * 8 - 1) / (256 * 8)
                               : 148 * 16);
  const bool to_fp8 = dst_dtype < 0;
  const int dt = to_fp8 ? src_dtype : dst_dtype;
  B200_CHECK(dt >= B200_F32 && dt <= B200_BF16, "Unsupported data type");
#define CV(T)                                                                                       \
  if (to_fp8)                                                                                       \
    convert_fp8_kernel<T, B200_KV_FP8_E4M3, true><<<blocks, 256, 0, st>>>(dst, src, numel, scale);  \
  else                                                                                              \
    convert_fp8_kernel<T, B200_KV_FP8_E4M3, false><<<blocks, 256, 0, st>>>(dst, src, numel, scale)
  if (dt == B200_BF16) { CV(__nv_bfloat16); }
  else if (dt == B200_F16) { CV(__half); }
  else { CV(float); }
#undef CV
  return check_launch("convert_fp8_kernel");
}
