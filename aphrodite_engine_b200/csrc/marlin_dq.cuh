// marlin_dq.cuh — device helpers shared by the Marlin-format GEMM kernels (tcgen05 large-batch kernel in
// marlin_gemm.cu, mma.sync small-batch kernel in marlin_gemm_small.cu): shared-memory / cp.async wrappers and the
// exact code -> 16-bit float dequantisation in Marlin's interleaved nibble / byte order.
#pragma once
#include "common.cuh"

#include <type_traits>

namespace b200 {

// zero-point flavours (C ABI `has_zp`): none (symmetric bias 8 / 128), packed integers (AWQ), 16-bit floats (HQQ)
enum { ZP_NONE = 0, ZP_INT = 1, ZP_FLOAT = 2 };

// explicit shared-state-space accesses: the 1024-B re-aligned dynamic smem pointer is a GENERIC pointer to
// the compiler (it would emit LD.E/ST.E through the generic path, which showed up as long-scoreboard stalls)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 ldg_stream128(const void* g) {    // read-once data: no L1 allocation
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

template <typename T> struct DQ;  // magic numbers of the int4 -> 16-bit float trick (exact integers)
template <> struct DQ<__nv_bfloat16> {
  static constexpr uint32_t MAGIC = 0x43004300u;          // 128.0 | 128.0 : 128 + q is exact (q < 128)
  static __device__ __forceinline__ uint32_t offset(int q) {  // bf16x2 of (128 + q)
    const uint32_t h = 0x4300u + (uint32_t)q;                 // 128+q: mantissa lsb = 1 in [128,256)
    return h | (h << 16);
  }
  static __device__ __forceinline__ uint32_t sub_mul(uint32_t x, uint32_t off, uint32_t s2) {
    __nv_bfloat162 v = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&off));
    v = __hmul2(v, *reinterpret_cast<__nv_bfloat162*>(&s2));
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct DQ<__half> {
  static constexpr uint32_t MAGIC = 0x64006400u;          // 1024.0 | 1024.0
  static __device__ __forceinline__ uint32_t offset(int q) {
    const uint32_t h = 0x6400u + (uint32_t)q;                 // 1024+q exact
    return h | (h << 16);
  }
  static __device__ __forceinline__ uint32_t sub_mul(uint32_t x, uint32_t off, uint32_t s2) {
    __half2 v = __hsub2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&off));
    v = __hmul2(v, *reinterpret_cast<__half2*>(&s2));
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {  // (a & b) | c
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// Weight-code -> 16-bit float pairs, generalised over the code width and the zero-point flavour. All variants
// produce (q - z) EXACTLY (|q - z| <= 255 fits the 8 significant bits of bf16 and the 11 of fp16) and then round
// once in the multiply by the scale — the arithmetic of the reference's dequant<> + sub_zp + scale chain
// (gptq_marlin.cu:156-360, 382-392, 366-379). With float zero points (HQQ) the subtraction rounds too, as there
// (sub_zpf, :393-403).
//   offset word `off` per output column: 4-bit / fp16-8-bit: 16-bit pair of MAGIC + z; bf16-8-bit: fp32 bits of
//   2^23 + z; float zero points: the 16-bit pair {zp, zp}.
template <typename T, int BITS, int ZP> struct WDQ {
  static __device__ __forceinline__ uint32_t offset_of(int z) {
    if constexpr (BITS == 8 && std::is_same<T, __nv_bfloat16>::value) return __float_as_uint(8388608.f + (float)z);
    else return DQ<T>::offset(z);
  }
  static __device__ __forceinline__ uint32_t finish(uint32_t x, uint32_t off, uint32_t s2) {  // x = MAGIC + q pair
    if constexpr (ZP == ZP_FLOAT) {
      const uint32_t magic = DQ<T>::MAGIC;
      __half2 v = __hsub2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<const __half2*>(&magic));
      v = __hsub2(v, *reinterpret_cast<__half2*>(&off));
      v = __hmul2(v, *reinterpret_cast<__half2*>(&s2));
      return *reinterpret_cast<uint32_t*>(&v);
    } else {
      return DQ<T>::sub_mul(x, off, s2);
    }
  }
  // 8-bit: bytes (lo, lo + 2) of `w` are the codes of k and k + 1
  static __device__ __forceinline__ uint32_t pair8(uint32_t w, int lo, uint32_t off, uint32_t s2) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      const float base = __uint_as_float(off);                       // 2^23 + z
      const float f0 = __uint_as_float(prmt(w, 0x4B000000u, lo ? 0x7651u : 0x7650u)) - base;
      const float f1 = __uint_as_float(prmt(w, 0x4B000000u, lo ? 0x7653u : 0x7652u)) - base;
      uint32_t x = prmt(__float_as_uint(f0), __float_as_uint(f1), 0x7632u);   // exact: |q - z| <= 255
      __nv_bfloat162 v = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&s2));
      return *reinterpret_cast<uint32_t*>(&v);
    } else {
      const uint32_t x = prmt(w, 0x64646464u, lo ? 0x5351u : 0x5250u);        // fp16 pair of 1024 + q
      return finish(x, off, s2);
    }
  }
};

// zero point of column e (= 2j + b) of a lane's 8 columns inside its packed zero-point word(s)
// (marlin_zero_points: aphrodite/quantization/utils/marlin_utils.py:198-217 — scale permutation, then the
// interleave [0,2,4,6,1,3,5,7] (4-bit) / [0,2,1,3] (8-bit) inside every int32)
template <int BITS> __device__ __forceinline__ int zp_code(int e, uint32_t z0, uint32_t z1) {
  if constexpr (BITS == 4) return (int)((z0 >> (4 * (((e & 1) << 2) | (e >> 1)))) & 0xFu);
  else {
    const int r = e & 3;
    return (int)((((e >> 2) ? z1 : z0) >> (8 * (((r & 1) << 1) | (r >> 1)))) & 0xFFu);
  }
}

// position of output column n (0..N) inside a Marlin-permuted scale row
// (aphrodite/quantization/utils/marlin_utils.py:172-196)
__device__ __forceinline__ int scale_pos(int n, bool grouped) {
  if (grouped) return (n & ~63) + 8 * (n & 7) + ((n & 63) >> 3);
  return (n & ~31) + 8 * ((n & 7) >> 1) + 2 * ((n & 31) >> 3) + (n & 1);
}

}  // namespace b200
