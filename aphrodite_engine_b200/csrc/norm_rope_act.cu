// norm_rope_act.cu — rms_norm / fused_add_rms_norm / rotary_embedding / *_and_mul activations, sm_100a.
//
// Replaces kernels/layernorm_kernels.cu (:24-50, :204-286), kernels/pos_encoding_kernels.cu (:10-93)
// and kernels/activation_kernels.cu (:13-160) of the reference. These are HBM-bound element-wise /
// row-reduction ops: the B200 design is 16-byte vector accesses, the activation row kept in registers
// between the reduction and the scaling pass (one global read), and flat grids that fill 148 SMs
// regardless of the token count. The ROUNDING POINTS follow the reference exactly, because they define
// the result in 16-bit types:
//   rms_norm           out = T( T(x * rsqrt(mean(x^2)+eps)) * w )                (layernorm_kernels.cu:47-48)
//   fused_add_rms_norm z = T(x + r); r = z; x = T( T(z * s) * w )                (:128-142, :231-250)
//   rotary             x' = T( T(x*cos) - T(y*sin) ), y' = T( T(y*cos) + T(x*sin) ) (pos_encoding_kernels.cu:31-34)
//   silu_and_mul       out = T( T(x / (1 + expf(-x))) * y )                      (activation_kernels.cu:21-31)
#include "common.cuh"

#include <math.h>

namespace b200 {

template <typename T> __device__ __forceinline__ T tmul(T a, T b) {
  return from_f32<T>(__fmul_rn(to_f32<T>(a), to_f32<T>(b)));
}
template <typename T> __device__ __forceinline__ T tadd(T a, T b) {
  return from_f32<T>(__fadd_rn(to_f32<T>(a), to_f32<T>(b)));
}
template <typename T> __device__ __forceinline__ T tsub(T a, T b) {
  return from_f32<T>(__fsub_rn(to_f32<T>(a), to_f32<T>(b)));
}

template <typename T> struct Vec16 {  // 16 bytes of T
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T e[N];
  };
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm (optionally fused with the residual add). One CTA per token, row cached in registers.
// ------------------------------------------------------------------------------------------------
template <typename T, bool FUSED_ADD, int VPT>
__global__ void __launch_bounds__(256)
rms_norm_vec_kernel(T* __restrict__ out, T* __restrict__ input, T* __restrict__ residual,
                    const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[8];
  constexpr int N = Vec16<T>::N;
  const int nvec = hidden / N;
  const int64_t row = (int64_t)blockIdx.x * hidden;
  Vec16<T> z[VPT];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * 256;
    if (v < nvec) {
      z[k].raw = *reinterpret_cast<const uint4*>(input + row + (int64_t)v * N);
      if (FUSED_ADD) {
        Vec16<T> r;
        r.raw = *reinterpret_cast<const uint4*>(residual + row + (int64_t)v * N);
#pragma unroll
        for (int e = 0; e < N; ++e) z[k].e[e] = tadd<T>(z[k].e[e], r.e[e]);
        *reinterpret_cast<uint4*>(residual + row + (int64_t)v * N) = z[k].raw;
      }
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float f = to_f32<T>(z[k].e[e]);
        ss = fmaf(f, f, ss);
      }
    }
  }
  const float tot = block_sum_256(ss, red);
  const float s = rsqrtf(tot / (float)hidden + eps);
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = threadIdx.x + k * 256;
    if (v < nvec) {
      Vec16<T> w;
      w.raw = *reinterpret_cast<const uint4*>(weight + (int64_t)v * N);
#pragma unroll
      for (int e = 0; e < N; ++e)
        z[k].e[e] = tmul<T>(from_f32<T>(__fmul_rn(to_f32<T>(z[k].e[e]), s)), w.e[e]);
      *reinterpret_cast<uint4*>(out + row + (int64_t)v * N) = z[k].raw;
    }
  }
}

// any hidden size / alignment: two passes over global memory
template <typename T, bool FUSED_ADD>
__global__ void __launch_bounds__(256)
rms_norm_scalar_kernel(T* __restrict__ out, T* __restrict__ input, T* __restrict__ residual,
                       const T* __restrict__ weight, float eps, int hidden) {
  __shared__ float red[8];
  const int64_t row = (int64_t)blockIdx.x * hidden;
  float ss = 0.f;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    T z = input[row + i];
    if (FUSED_ADD) {
      z = tadd<T>(z, residual[row + i]);
      residual[row + i] = z;
    }
    const float f = to_f32<T>(z);
    ss = fmaf(f, f, ss);
  }
  const float tot = block_sum_256(ss, red);
  const float s = rsqrtf(tot / (float)hidden + eps);
  const T* src = FUSED_ADD ? residual : input;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    const float f = to_f32<T>(src[row + i]);
    out[row + i] = tmul<T>(from_f32<T>(__fmul_rn(f, s)), weight[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Rotary embedding (GPT-NeoX halves / GPT-J interleaved), in place on q and k.
// Flat grid over (token, head, pair-chunk): VEC pairs per thread.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void rot_pair(T& x, T& y, T c, T s) {
  const T nx = tsub<T>(tmul<T>(x, c), tmul<T>(y, s));
  const T ny = tadd<T>(tmul<T>(y, c), tmul<T>(x, s));
  x = nx;
  y = ny;
}

template <typename T, bool NEOX, bool VEC>
__global__ void __launch_bounds__(256)
rotary_kernel(const int64_t* __restrict__ positions, T* __restrict__ query, T* __restrict__ key,
              const T* __restrict__ cache, const int64_t* __restrict__ offsets, int num_tokens,
              int num_heads, int num_kv_heads, int head_size, int rot_dim, int64_t q_stride,
              int64_t k_stride) {
  constexpr int N = VEC ? Vec16<T>::N : 1;           // elements per 16-byte vector
  constexpr int PAIRS = VEC ? (NEOX ? N : N / 2) : 1;  // pairs handled per thread
  const int embed = rot_dim / 2;
  const int chunks = embed / PAIRS;                    // per head
  const int heads = num_heads + num_kv_heads;
  const int64_t total = (int64_t)num_tokens * heads * chunks;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % chunks);
    const int h = (int)((idx / chunks) % heads);
    const int64_t tok = idx / ((int64_t)chunks * heads);
    int64_t pos = positions[tok];
    if (offsets != nullptr) pos += offsets[tok];
    const T* cosp = cache + pos * rot_dim;
    const T* sinp = cosp + embed;
    T* arr = (h < num_heads) ? query + tok * q_stride + (int64_t)h * head_size
                             : key + tok * k_stride + (int64_t)(h - num_heads) * head_size;
    const int p0 = c * PAIRS;
    if constexpr (!VEC) {
      const int xi = NEOX ? p0 : 2 * p0, yi = NEOX ? embed + p0 : 2 * p0 + 1;
      T x = arr[xi], y = arr[yi];
      rot_pair<T>(x, y, cosp[p0], sinp[p0]);
      arr[xi] = x;
      arr[yi] = y;
    } else if constexpr (NEOX) {
      Vec16<T> x, y, cs, sn;
      x.raw = *reinterpret_cast<const uint4*>(arr + p0);
      y.raw = *reinterpret_cast<const uint4*>(arr + embed + p0);
      cs.raw = *reinterpret_cast<const uint4*>(cosp + p0);
      sn.raw = *reinterpret_cast<const uint4*>(sinp + p0);
#pragma unroll
      for (int e = 0; e < N; ++e) rot_pair<T>(x.e[e], y.e[e], cs.e[e], sn.e[e]);
      *reinterpret_cast<uint4*>(arr + p0) = x.raw;
      *reinterpret_cast<uint4*>(arr + embed + p0) = y.raw;
    } else {
      Vec16<T> xy;
      xy.raw = *reinterpret_cast<const uint4*>(arr + 2 * p0);
#pragma unroll
      for (int e = 0; e < N / 2; ++e) rot_pair<T>(xy.e[2 * e], xy.e[2 * e + 1], cosp[p0 + e], sinp[p0 + e]);
      *reinterpret_cast<uint4*>(arr + 2 * p0) = xy.raw;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// rotary_embedding + reshape_and_cache in ONE launch (decode: both are latency-bound singletons).
// Same outputs as the two reference kernels run back to back (pos_encoding_kernels.cu:71-93 then
// cache_kernels.cu:152-204): q and k rotated in place, the ROTATED k and v scattered to slot_mapping[t].
// Flat grid over (token, {q heads | k heads | v heads}, 16-byte chunk). NeoX style, rot_dim == head_size.
// ------------------------------------------------------------------------------------------------
template <typename T, int KV>
__global__ void __launch_bounds__(256)
rope_and_cache_kernel(const int64_t* __restrict__ positions, T* __restrict__ query, T* __restrict__ key,
                      const T* __restrict__ value, const T* __restrict__ cache, void* __restrict__ key_cache,
                      void* __restrict__ value_cache, const int64_t* __restrict__ slot_mapping, int num_tokens,
                      int num_heads, int num_kv_heads, int head_size, int64_t q_stride, int64_t k_stride,
                      int64_t v_stride, int block_size, int x, float k_scale, float v_scale) {
  constexpr int N = Vec16<T>::N;                      // 8 elements per 16-byte vector (16-bit types only)
  const int embed = head_size / 2;
  const int rc = embed / N;                           // rotary chunks per head (each = N pairs)
  const int vcn = head_size / N;                      // value chunks per head
  const int q_items = num_heads * rc, k_items = num_kv_heads * rc, v_items = num_kv_heads * vcn;
  const int per_tok = q_items + k_items + v_items;
  const int64_t total = (int64_t)num_tokens * per_tok;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = idx / per_tok;
    int r = (int)(idx % per_tok);
    if (r < q_items + k_items) {
      const bool is_k = r >= q_items;
      if (is_k) r -= q_items;
      const int h = r / rc, c = r % rc;
      const int64_t pos = positions[tok];
      const T* cosp = cache + pos * head_size;
      const T* sinp = cosp + embed;
      T* arr = is_k ? key + tok * k_stride + (int64_t)h * head_size : query + tok * q_stride + (int64_t)h * head_size;
      const int p0 = c * N;
      Vec16<T> xv, yv, cs, sn;
      xv.raw = *reinterpret_cast<const uint4*>(arr + p0);
      yv.raw = *reinterpret_cast<const uint4*>(arr + embed + p0);
      cs.raw = __ldg(reinterpret_cast<const uint4*>(cosp + p0));
      sn.raw = __ldg(reinterpret_cast<const uint4*>(sinp + p0));
#pragma unroll
      for (int e = 0; e < N; ++e) rot_pair<T>(xv.e[e], yv.e[e], cs.e[e], sn.e[e]);
      *reinterpret_cast<uint4*>(arr + p0) = xv.raw;
      *reinterpret_cast<uint4*>(arr + embed + p0) = yv.raw;
      if (is_k) {
        const int64_t slot = slot_mapping[tok];
        if (slot >= 0) {
          const int64_t blk = slot / block_size, boff = slot % block_size;
          const int64_t hb = (blk * num_kv_heads + h) * (head_size / x);
          if constexpr (KV == B200_KV_AUTO) {
            T* kc = reinterpret_cast<T*>(key_cache);      // x == N: one 16-byte run per chunk
            *reinterpret_cast<uint4*>(kc + ((hb + p0 / x) * block_size + boff) * x) = xv.raw;
            *reinterpret_cast<uint4*>(kc + ((hb + (embed + p0) / x) * block_size + boff) * x) = yv.raw;
          } else {
            uint8_t* kc = reinterpret_cast<uint8_t*>(key_cache);
            union { uint2 raw; uint8_t b[8]; } qx, qy;
#pragma unroll
            for (int e = 0; e < N; ++e) {
              qx.b[e] = fp8_quant<T, KV>(xv.e[e], k_scale);
              qy.b[e] = fp8_quant<T, KV>(yv.e[e], k_scale);
            }
            // x == 16 cache bytes per run; an 8-element chunk is half a run (8-byte aligned)
            *reinterpret_cast<uint2*>(kc + ((hb + p0 / x) * block_size + boff) * x + p0 % x) = qx.raw;
            *reinterpret_cast<uint2*>(kc + ((hb + (embed + p0) / x) * block_size + boff) * x + (embed + p0) % x) = qy.raw;
          }
        }
      }
    } else {
      r -= q_items + k_items;
      const int64_t slot = slot_mapping[tok];
      if (slot < 0) continue;
      const int h = r / vcn, c = r % vcn;
      const int64_t blk = slot / block_size, boff = slot % block_size;
      Vec16<T> vv;
      vv.raw = __ldg(reinterpret_cast<const uint4*>(value + tok * v_stride + (int64_t)h * head_size + c * N));
      const int64_t dst = ((blk * num_kv_heads + h) * head_size + c * N) * block_size + boff;
      if constexpr (KV == B200_KV_AUTO) {
        T* vc = reinterpret_cast<T*>(value_cache);
#pragma unroll
        for (int e = 0; e < N; ++e) vc[dst + (int64_t)e * block_size] = vv.e[e];
      } else {
        uint8_t* vc = reinterpret_cast<uint8_t*>(value_cache);
#pragma unroll
        for (int e = 0; e < N; ++e) vc[dst + (int64_t)e * block_size] = fp8_quant<T, KV>(vv.e[e], v_scale);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Activations
// ------------------------------------------------------------------------------------------------
enum { ACT_SILU = 0, ACT_GELU = 1, ACT_GELU_TANH = 2 };
enum { ACT_GELU_NEW = 0, ACT_GELU_FAST = 1, ACT_GELU_QUICK = 2 };

template <typename T, int ACT> __device__ __forceinline__ T gate_act(T xv) {
  const float f = to_f32<T>(xv);
  if constexpr (ACT == ACT_SILU) {
    return from_f32<T>(f / (1.0f + expf(-f)));
  } else if constexpr (ACT == ACT_GELU) {
    return from_f32<T>(f * 0.5f * (1.0f + erff(f * 0.70710678118654752440f)));
  } else {
    const float beta = 1.41421356237309504880f * 1.12837916709551257390f * 0.5f;
    const float inner = beta * (f + 0.044715f * f * f * f);
    return from_f32<T>(0.5f * f * (1.0f + tanhf(inner)));
  }
}

template <typename T, int ACT> __device__ __forceinline__ T unary_act(T x) {
  const float f = to_f32<T>(x);
  if constexpr (ACT == ACT_GELU_NEW) {
    const float x3 = to_f32<T>(tmul<T>(tmul<T>(x, x), x));
    const T inner = tadd<T>(x, from_f32<T>(0.044715f * x3));
    const T t = from_f32<T>(tanhf(to_f32<T>(from_f32<T>(0.79788456f * to_f32<T>(inner)))));
    return tmul<T>(tmul<T>(from_f32<T>(0.5f), x), tadd<T>(from_f32<T>(1.0f), t));
  } else if constexpr (ACT == ACT_GELU_FAST) {
    const T a = from_f32<T>(f * 0.79788456f);
    const T b = tadd<T>(from_f32<T>(1.0f), tmul<T>(from_f32<T>(0.044715f * f), x));
    const T t = from_f32<T>(tanhf(to_f32<T>(tmul<T>(a, b))));
    return tmul<T>(tmul<T>(from_f32<T>(0.5f), x), tadd<T>(from_f32<T>(1.0f), t));
  } else {
    return from_f32<T>(f / (1.0f + expf(-1.702f * f)));
  }
}

// out[t, i] = act(in[t, i]) * in[t, d + i]; flat grid over 16-byte vectors
template <typename T, int ACT, bool VEC>
__global__ void __launch_bounds__(256)
act_and_mul_kernel(T* __restrict__ out, const T* __restrict__ in, int64_t num_tokens, int d) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  const int per_tok = d / N;
  const int64_t total = num_tokens * per_tok;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = idx / per_tok;
    const int c = (int)(idx % per_tok);
    const T* xr = in + tok * 2 * d + (int64_t)c * N;
    if constexpr (VEC) {
      Vec16<T> x, y;
      x.raw = __ldg(reinterpret_cast<const uint4*>(xr));
      y.raw = __ldg(reinterpret_cast<const uint4*>(xr + d));
#pragma unroll
      for (int e = 0; e < N; ++e) x.e[e] = tmul<T>(gate_act<T, ACT>(x.e[e]), y.e[e]);
      *reinterpret_cast<uint4*>(out + tok * d + (int64_t)c * N) = x.raw;
    } else {
      out[tok * d + c] = tmul<T>(gate_act<T, ACT>(xr[0]), xr[d]);
    }
  }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
activation_kernel(T* __restrict__ out, const T* __restrict__ in, int64_t numel) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < numel;
       idx += (int64_t)gridDim.x * blockDim.x)
    out[idx] = unary_act<T, ACT>(in[idx]);
}

static inline int flat_grid(int64_t work_items) {
  int64_t b = (work_items + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T, bool FUSED>
static int launch_rms(T* out, T* input, T* residual, const T* weight, float eps, int num_tokens,
                      int hidden, cudaStream_t st) {
  constexpr int N = Vec16<T>::N;
  const bool aligned = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(input) |
                         reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0;
  const int nvec = hidden / N;
  if (aligned && hidden % N == 0 && nvec <= 256 * 4) {
    if (nvec <= 256)
      rms_norm_vec_kernel<T, FUSED, 1><<<num_tokens, 256, 0, st>>>(out, input, residual, weight, eps, hidden);
    else if (nvec <= 512)
      rms_norm_vec_kernel<T, FUSED, 2><<<num_tokens, 256, 0, st>>>(out, input, residual, weight, eps, hidden);
    else
      rms_norm_vec_kernel<T, FUSED, 4><<<num_tokens, 256, 0, st>>>(out, input, residual, weight, eps, hidden);
  } else {
    rms_norm_scalar_kernel<T, FUSED><<<num_tokens, 256, 0, st>>>(out, input, residual, weight, eps, hidden);
  }
  return check_launch("rms_norm_kernel");
}

}  // namespace b200

using namespace b200;

#define B200_DISPATCH_T(dtype, FN)                      \
  do {                                                  \
    if (dtype == B200_BF16) { FN(__nv_bfloat16); }      \
    else if (dtype == B200_F16) { FN(__half); }         \
    else { FN(float); }                                 \
  } while (0)

extern "C" int b200_rms_norm(void* out, const void* input, const void* weight, float epsilon,
                             int num_tokens, int hidden_size, int dtype, void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "rms_norm: unsupported dtype");
  if (num_tokens == 0 || hidden_size == 0) return 0;
  int rc = 0;
#define FN(T) rc = launch_rms<T, false>((T*)out, (T*)input, (T*)nullptr, (const T*)weight, epsilon, \
                                        num_tokens, hidden_size, (cudaStream_t)stream)
  B200_DISPATCH_T(dtype, FN);
#undef FN
  return rc;
}

extern "C" int b200_fused_add_rms_norm(void* input, void* residual, const void* weight,
                                       float epsilon, int num_tokens, int hidden_size, int dtype,
                                       void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "fused_add_rms_norm: unsupported dtype");
  if (num_tokens == 0 || hidden_size == 0) return 0;
  int rc = 0;
#define FN(T) rc = launch_rms<T, true>((T*)input, (T*)input, (T*)residual, (const T*)weight, epsilon, \
                                       num_tokens, hidden_size, (cudaStream_t)stream)
  B200_DISPATCH_T(dtype, FN);
#undef FN
  return rc;
}

extern "C" int b200_rotary_embedding(const int64_t* positions, void* query, void* key,
                                     const void* cos_sin_cache,
                                     const int64_t* cos_sin_cache_offsets, int num_tokens,
                                     int num_heads, int num_kv_heads, int head_size, int rot_dim,
                                     int64_t query_stride, int64_t key_stride, int is_neox,
                                     int dtype, void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "rotary_embedding: unsupported dtype");
  B200_CHECK(rot_dim % 2 == 0 && rot_dim <= head_size, "rotary_embedding: bad rot_dim");
  if (num_tokens == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = dtype == B200_F32 ? 4 : 2;
  const int n = 16 / esz;
  const int embed = rot_dim / 2;
  const bool aligned =
      ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(key) |
        reinterpret_cast<uintptr_t>(cos_sin_cache)) & 15) == 0 &&
      (query_stride * esz) % 16 == 0 && (key_stride * esz) % 16 == 0 && (head_size * esz) % 16 == 0 &&
      (rot_dim * esz) % 16 == 0;
  const bool vec = aligned && (is_neox ? (embed % n == 0) : (rot_dim % n == 0 && n >= 2));
  const int pairs = vec ? (is_neox ? n : n / 2) : 1;
  const int64_t work = (int64_t)num_tokens * (num_heads + num_kv_heads) * (embed / pairs);
  const int grid = flat_grid(work);
#define FN(T)                                                                                      \
  do {                                                                                             \
    if (is_neox) {                                                                                 \
      if (vec) rotary_kernel<T, true, true><<<grid, 256, 0, st>>>(positions, (T*)query, (T*)key,   \
          (const T*)cos_sin_cache, cos_sin_cache_offsets, num_tokens, num_heads, num_kv_heads,     \
          head_size, rot_dim, query_stride, key_stride);                                           \
      else rotary_kernel<T, true, false><<<grid, 256, 0, st>>>(positions, (T*)query, (T*)key,      \
          (const T*)cos_sin_cache, cos_sin_cache_offsets, num_tokens, num_heads, num_kv_heads,     \
          head_size, rot_dim, query_stride, key_stride);                                           \
    } else {                                                                                       \
      if (vec) rotary_kernel<T, false, true><<<grid, 256, 0, st>>>(positions, (T*)query, (T*)key,  \
          (const T*)cos_sin_cache, cos_sin_cache_offsets, num_tokens, num_heads, num_kv_heads,     \
          head_size, rot_dim, query_stride, key_stride);                                           \
      else rotary_kernel<T, false, false><<<grid, 256, 0, st>>>(positions, (T*)query, (T*)key,     \
          (const T*)cos_sin_cache, cos_sin_cache_offsets, num_tokens, num_heads, num_kv_heads,     \
          head_size, rot_dim, query_stride, key_stride);                                           \
    }                                                                                              \
  } while (0)
  B200_DISPATCH_T(dtype, FN);
#undef FN
  return check_launch("rotary_kernel");
}

extern "C" int b200_rotary_embedding_and_cache(const int64_t* positions, void* query, void* key, const void* value,
                                               const void* cos_sin_cache, void* key_cache, void* value_cache,
                                               const int64_t* slot_mapping, int num_tokens, int num_heads,
                                               int num_kv_heads, int head_size, int rot_dim, int64_t query_stride,
                                               int64_t key_stride, int64_t value_stride, int is_neox, int block_size,
                                               int x, int dtype, int kv_dtype, float k_scale, float v_scale,
                                               void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "rotary_embedding_and_cache: unsupported dtype");
  B200_CHECK(kv_dtype >= B200_KV_AUTO && kv_dtype <= B200_KV_FP8_E5M2, "rotary_embedding_and_cache: bad kv_cache_dtype");
  if (num_tokens == 0) return 0;
  const bool aligned =
      ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(value) |
        reinterpret_cast<uintptr_t>(cos_sin_cache) | reinterpret_cast<uintptr_t>(key_cache)) & 15) == 0 &&
      (query_stride % 8) == 0 && (key_stride % 8) == 0 && (value_stride % 8) == 0;
  const bool fused = dtype != B200_F32 && is_neox && rot_dim == head_size && head_size % 16 == 0 && aligned &&
                     x == (kv_dtype == B200_KV_AUTO ? 8 : 16);
  if (!fused) {   // any other configuration: the two kernels back to back (identical results by definition)
    int rc = b200_rotary_embedding(positions, query, key, cos_sin_cache, nullptr, num_tokens, num_heads, num_kv_heads,
                                   head_size, rot_dim, query_stride, key_stride, is_neox, dtype, stream);
    if (rc != 0) return rc;
    return b200_reshape_and_cache(key, value, key_cache, value_cache, slot_mapping, num_tokens, num_kv_heads, head_size,
                                  block_size, x, key_stride, value_stride, dtype, kv_dtype, k_scale, v_scale, stream);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int per_tok = (num_heads + num_kv_heads) * (head_size / 16) + num_kv_heads * (head_size / 8);
  const int grid = flat_grid((int64_t)num_tokens * per_tok);
#define B200_RC(T, KVD)                                                                                          \
  rope_and_cache_kernel<T, KVD><<<grid, 256, 0, st>>>(positions, (T*)query, (T*)key, (const T*)value,             \
      (const T*)cos_sin_cache, key_cache, value_cache, slot_mapping, num_tokens, num_heads, num_kv_heads, head_size, \
      query_stride, key_stride, value_stride, block_size, x, k_scale, v_scale)
#define B200_RC_T(T)                                              \
  do {                                                            \
    if (kv_dtype == B200_KV_AUTO) B200_RC(T, B200_KV_AUTO);       \
    else if (kv_dtype == B200_KV_FP8_E4M3) B200_RC(T, B200_KV_FP8_E4M3); \
    else B200_RC(T, B200_KV_FP8_E5M2);                            \
  } while (0)
  if (dtype == B200_BF16) B200_RC_T(__nv_bfloat16); else B200_RC_T(__half);
#undef B200_RC_T
#undef B200_RC
  return check_launch("rope_and_cache_kernel");
}

extern "C" int b200_act_and_mul(void* out, const void* input, int num_tokens, int d, int act,
                                int dtype, void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "act_and_mul: unsupported dtype");
  B200_CHECK(act >= 0 && act <= 2, "act_and_mul: unknown activation");
  if (num_tokens == 0 || d == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = dtype == B200_F32 ? 4 : 2;
  const int n = 16 / esz;
  const bool vec = ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(input)) & 15) == 0 &&
                   d % n == 0;
  const int grid = flat_grid((int64_t)num_tokens * (d / (vec ? n : 1)));
#define LAUNCH_ACT(T, A)                                                                          \
  do {                                                                                            \
    if (vec) act_and_mul_kernel<T, A, true><<<grid, 256, 0, st>>>((T*)out, (const T*)input, num_tokens, d);  \
    else act_and_mul_kernel<T, A, false><<<grid, 256, 0, st>>>((T*)out, (const T*)input, num_tokens, d);     \
  } while (0)
#define FN(T)                                           \
  do {                                                  \
    if (act == 0) LAUNCH_ACT(T, ACT_SILU);              \
    else if (act == 1) LAUNCH_ACT(T, ACT_GELU);         \
    else LAUNCH_ACT(T, ACT_GELU_TANH);                  \
  } while (0)
  B200_DISPATCH_T(dtype, FN);
#undef FN
#undef LAUNCH_ACT
  return check_launch("act_and_mul_kernel");
}

extern "C" int b200_activation(void* out, const void* input, int num_tokens, int d, int act,
                               int dtype, void* stream) {
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "activation: unsupported dtype");
  B200_CHECK(act >= 0 && act <= 2, "activation: unknown activation");
  const int64_t numel = (int64_t)num_tokens * d;
  if (numel == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = flat_grid(numel);
#define FN(T)                                                                                       \
  do {                                                                                              \
    if (act == 0) activation_kernel<T, ACT_GELU_NEW><<<grid, 256, 0, st>>>((T*)out, (const T*)input, numel);        \
    else if (act == 1) activation_kernel<T, ACT_GELU_FAST><<<grid, 256, 0, st>>>((T*)out, (const T*)input, numel);  \
    else activation_kernel<T, ACT_GELU_QUICK><<<grid, 256, 0, st>>>((T*)out, (const T*)input, numel);               \
  } while (0)
  B200_DISPATCH_T(dtype, FN);
#undef FN
  return check_launch("activation_kernel");
}
