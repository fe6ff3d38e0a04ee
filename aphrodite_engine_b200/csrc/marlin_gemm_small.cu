// marlin_gemm_small.cu — W4A16 GEMM in the Marlin weight format for SMALL batches (M <= 32 tokens), sm_100a.
//
// Same contract as marlin_gemm.cu (gptq_marlin_gemm, kernels/quantization/gptq_marlin/gptq_marlin.cu:2247-2430),
// different regime: with a handful of tokens the op is a stream over the packed weights (0.5 byte per weight,
// 2*M flop per weight), and the tcgen05 kernel's fixed per-stage costs (dequantised tile through shared memory,
// proxy fence, mbarrier round trip, single-thread MMA issue: ~600 cycles per 64 x 128 weights regardless of M)
// cap it near 1 TB/s of packed weights. Here every warp is an independent stream:
//   * warp (nb, kq) of a CTA owns Marlin block nb (64 channels) of the CTA's 128-channel tile and the k-tiles
//     kq, kq+4, ... of the CTA's k range. Per k-tile each lane cp.async's ONE uint4 of packed words (its four
//     mma.sync B fragments in Marlin's layout), its 16 bytes of group scales, its zero-point word and its share of
//     the [tokens x 16] activation slice into a private DEPTH-deep ring — no barriers, no producer warp;
//   * dequant = lop3 nibble pairs + exact (q - 8 | zp) + one rounding multiply (bit-identical weights to the
//     tcgen05 kernel and to the reference's w_ref), results ARE the mma.sync.m16n8k16 B fragments (registers);
//     A fragments come from the activation slice with ldmatrix.x4 (swizzled 16-byte pieces, conflict-free);
//   * fp32 accumulators in registers; the four k-interleaved warps of a block are summed through shared memory;
//   * work = (128-channel tile, 64-k chunk) units in tile-major order, cut into gridDim.x EQUAL contiguous ranges
//     (stream-k): every CTA runs the same number of chunks whatever N / 128 is (224 tiles on 148 SMs would otherwise
//     quantise to waves). A CTA's range crosses at most a few tiles; each (tile, range) segment is reduced over the
//     four k-lanes through shared memory; tiles covered by several CTAs go to fp32 slabs [segment][M][N] and the LAST
//     CTA of a tile (atomic ticket on the reference's zeroed `workspace`, returned to zero) adds them in segment
//     order: deterministic, no spinning, no co-residency requirement, CUDA-graph capturable.
#include "common.cuh"
#include "marlin_dq.cuh"

#include <algorithm>

namespace b200 {

static constexpr int SM_WARPS = 8;
static constexpr int SM_THREADS = SM_WARPS * 32;
static constexpr int SM_DEPTH = 6;            // k-tiles in flight per warp
static constexpr int SM_KQ = 4;               // k-interleave: warps per Marlin block

struct SmallParams {
  const void* a;            // [M, K] T
  const uint32_t* b_q;      // [K/16, N*2] int32 Marlin layout
  const void* scales;       // [groups, N] T, Marlin-permuted
  const uint32_t* zeros;    // [groups, N/8] int32 or nullptr
  void* c;                  // [M, N] T
  float* c_tmp;             // [split, M, N] fp32 (split > 1)
  int* locks;               // per channel tile, zero on entry / exit
  int M, N, K;
  int grouped;              // 1: one scale row per k-group
  int ktiles_per_group;     // group_size / 16 (grouped)
  int kpg_shift;            // log2(ktiles_per_group) when a power of two, else -1
  int chunks;               // K / 64
  int tiles;                // ceil(N / 128)
};

template <int MB> struct SmallCfg {
  static constexpr int A_OFF = 1152;                         // words 512 | scales 512 | zero points 128
  static constexpr int SLOT = A_OFF + 512 * MB;              // + [16*MB tokens x 16 k] activation slice
  static constexpr int RING = SM_WARPS * SM_DEPTH * SLOT;
  static constexpr int RED_STRIDE = 132;                     // floats per reduce row (128 + pad: conflict-free stores)
  static constexpr int REDUCE = SM_KQ * 16 * MB * RED_STRIDE * 4;   // [kq][rows][128 ch] fp32, reuses the ring
  static constexpr int SMEM = RING > REDUCE ? RING : REDUCE;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* g, bool valid) {
  const uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(n) : "memory");
}

template <typename T, int ZP, int MB>
__global__ void __launch_bounds__(SM_THREADS, 2)
marlin_w4a16_small_kernel(const SmallParams p) {
  using Cfg = SmallCfg<MB>;
  using W = WDQ<T, 4, ZP>;
  extern __shared__ __align__(128) uint8_t sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = warp & 1, kq = warp >> 1;
  const int m = lane & 3, cq = lane >> 2;
  const T* sc = reinterpret_cast<const T*>(p.scales);
  const T* a = reinterpret_cast<const T*>(p.a);
  T* cptr = reinterpret_cast<T*>(p.c);
  __shared__ int s_last;

  // this CTA's contiguous range of (tile, chunk) units
  const long long U = (long long)p.tiles * p.chunks;
  const long long G = gridDim.x;
  const long long u_end = (long long)(blockIdx.x + 1) * U / G;
  auto cta_of = [&](long long u) { return (int)(((u + 1) * G - 1) / U); };   // owner of unit u

  for (long long u = (long long)blockIdx.x * U / G; u < u_end;) {
  const int tile = (int)(u / p.chunks);
  const int cb = (int)(u - (long long)tile * p.chunks);
  const int ce = (int)min((long long)p.chunks, cb + (u_end - u));
  u += ce - cb;
  const int n_base = tile * 128;
  const int nblk = min(2, (p.N - n_base) / 64);
  const int kt0 = cb * 4, kt1 = ce * 4;

  float acc[MB][8][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int f = 0; f < 8; ++f)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mb][f][i] = 0.f;

  if (nb < nblk) {
    const int col_base = n_base + nb * 64;                    // first channel of this warp's Marlin block
    uint32_t s2[8], off2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s2[e] = 0; off2[e] = W::offset_of(8); }
    if (!p.grouped) {
      uint32_t z0 = 0;
      if constexpr (ZP == ZP_INT) z0 = p.zeros[col_base / 8 + cq];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int n = col_base + cq + 16 * (e >> 1) + 8 * (e & 1);
        const T sv = sc[scale_pos(n, false)];
        const uint32_t s16 = *reinterpret_cast<const uint16_t*>(&sv);
        s2[e] = s16 | (s16 << 16);
        if constexpr (ZP == ZP_INT) off2[e] = W::offset_of(zp_code<4>(e, z0, 0));
      }
    }
    const uint32_t ring = smem_u32(sm_raw) + (uint32_t)warp * (SM_DEPTH * Cfg::SLOT);
    const int row_words = p.N * 2;
    const uint32_t* wsrc = p.b_q + (size_t)(col_base / 64) * 128 + (size_t)lane * 4;
    const int nkt = (kt1 - kt0 - kq + SM_KQ - 1) / SM_KQ;     // k-tiles of this warp (may be <= 0)

    // producer state, advanced incrementally (no divisions / 64-bit multiplies in the loop): next k-tile to fetch,
    // its ring slot, its packed-word pointer, its activation column and its scale group
    int pf_i = 0, pf_kt = kt0 + kq, pf_g = -1;
    uint32_t pf_slot = ring;
    const uint32_t* pf_w = wsrc + (size_t)pf_kt * row_words;
    const size_t w_step = (size_t)SM_KQ * row_words;
    auto group_of = [&](int kt) { return p.kpg_shift >= 0 ? (kt >> p.kpg_shift) : kt / p.ktiles_per_group; };
    auto prefetch = [&]() {
      if (pf_i < nkt) {
        cp_async16(pf_slot + lane * 16, pf_w);
        if (p.grouped) {
          const int g = group_of(pf_kt);
          if (g != pf_g) {                                     // first k-tile of this warp in a new group
            pf_g = g;
            cp_async16(pf_slot + 512 + lane * 16,
                       reinterpret_cast<const uint8_t*>(sc) + ((size_t)g * p.N + col_base + 8 * cq) * 2);
            if constexpr (ZP == ZP_INT)
              cp_async4(pf_slot + 1024 + lane * 4, p.zeros + (size_t)g * (p.N / 8) + col_base / 8 + cq);
          }
        }
#pragma unroll
        for (int j = 0; j < MB; ++j) {                        // activation slice: 16-byte piece (row r, k-half h)
          const int idx = lane + 32 * j;
          const int r = idx >> 1, h = idx & 1;
          const uint32_t dst = pf_slot + Cfg::A_OFF + (uint32_t)(((r * 2 + h) ^ ((r >> 2) & 1)) * 16);
          const bool ok = r < p.M;
          cp_async16_zfill(dst, a + (size_t)(ok ? r : 0) * p.K + pf_kt * 16 + h * 8, ok);
        }
        ++pf_i;
        pf_kt += SM_KQ;
        pf_w += w_step;
        pf_slot = (pf_slot + Cfg::SLOT == ring + SM_DEPTH * Cfg::SLOT) ? ring : pf_slot + Cfg::SLOT;
      }
      cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < SM_DEPTH; ++d) prefetch();

    int g_cur = -1;
    uint32_t slot = ring;
    int kt = kt0 + kq;
    for (int i = 0; i < nkt; ++i, kt += SM_KQ, slot = (slot + Cfg::SLOT == ring + SM_DEPTH * Cfg::SLOT) ? ring : slot + Cfg::SLOT) {
      cp_async_wait<SM_DEPTH - 1>();
      __syncwarp();                                           // A pieces were copied by other lanes of this warp
      const uint4 q = lds128(slot + lane * 16);
      if (p.grouped) {
        const int g = group_of(kt);
        if (g != g_cur) {                                     // warp-uniform: new scale row (and zero points)
          g_cur = g;
          const uint4 sv = lds128(slot + 512 + lane * 16);
          const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
          uint32_t z0 = 0;
          if constexpr (ZP == ZP_INT) z0 = lds32(slot + 1024 + lane * 4);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t s16 = (sw[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            s2[e] = s16 | (s16 << 16);
            if constexpr (ZP == ZP_INT) off2[e] = W::offset_of(zp_code<4>(e, z0, 0));
          }
        }
      }
      uint32_t af[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int mi = lane >> 3, rr = lane & 7;
        const int r = mb * 16 + (mi & 1) * 8 + rr, h = mi >> 1;
        ldmatrix_x4(af[mb][0], af[mb][1], af[mb][2], af[mb][3],
                    slot + Cfg::A_OFF + (uint32_t)(((r * 2 + h) ^ ((r >> 2) & 1)) * 16));
      }
      const uint32_t wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = wq[j];
        // B fragments of the two 8-channel halves of 16-channel sub-tile j: {k 2m,2m+1} and {k 2m+8,2m+9}
        const uint32_t b00 = W::finish(lop3_and_or(w, 0x000f000fu, DQ<T>::MAGIC), off2[2 * j], s2[2 * j]);
        const uint32_t b01 = W::finish(lop3_and_or(w >> 4, 0x000f000fu, DQ<T>::MAGIC), off2[2 * j], s2[2 * j]);
        const uint32_t b10 = W::finish(lop3_and_or(w >> 8, 0x000f000fu, DQ<T>::MAGIC), off2[2 * j + 1], s2[2 * j + 1]);
        const uint32_t b11 = W::finish(lop3_and_or(w >> 12, 0x000f000fu, DQ<T>::MAGIC), off2[2 * j + 1], s2[2 * j + 1]);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          mma_16816<T>(acc[mb][2 * j], af[mb], b00, b01);
          mma_16816<T>(acc[mb][2 * j + 1], af[mb], b10, b11);
        }
      }
      __syncwarp();                                           // every lane is done with this slot's A pieces
      prefetch();
    }
    cp_async_wait<0>();
  }
  __syncthreads();                                            // rings are dead: reuse them as the reduce buffer

  // ---- sum the four k-interleaved warps of every block: red[kq][row][128 ch] ----
  float* red = reinterpret_cast<float*>(sm_raw);
  constexpr int ROWS = 16 * MB;
  if (nb < nblk) {
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        const int ch = nb * 64 + f * 8 + 2 * m;
        float* r0 = red + ((size_t)kq * ROWS + mb * 16 + cq) * Cfg::RED_STRIDE + ch;
        *reinterpret_cast<float2*>(r0) = make_float2(acc[mb][f][0], acc[mb][f][1]);
        *reinterpret_cast<float2*>(r0 + 8 * Cfg::RED_STRIDE) = make_float2(acc[mb][f][2], acc[mb][f][3]);
      }
  }
  __syncthreads();
  const int nvalid = min(128, p.N - n_base);
  // segments of this tile: one per CTA whose range intersects it (consecutive CTAs)
  const int cf = cta_of((long long)tile * p.chunks);
  const int nseg = cta_of((long long)tile * p.chunks + p.chunks - 1) - cf + 1;
  float* slab = nseg > 1 ? p.c_tmp + (size_t)((int)blockIdx.x - cf) * p.M * p.N : nullptr;
  // thread -> (row, 4 channels): 32 threads cover a row of 128 channels
  for (int r = warp; r < p.M && r < ROWS; r += SM_WARPS) {
    const int ch = lane * 4;
    if (ch < nvalid) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < SM_KQ; ++k) {
        const float4 t = *reinterpret_cast<const float4*>(red + ((size_t)k * ROWS + r) * Cfg::RED_STRIDE + ch);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
      const size_t off = (size_t)r * p.N + n_base + ch;
      if (slab != nullptr) {
        __stcg(reinterpret_cast<float4*>(slab + off), v);
      } else {
        uint2 o;
        o.x = pack2<T>(v.x, v.y);
        o.y = pack2<T>(v.z, v.w);
        *reinterpret_cast<uint2*>(cptr + off) = o;
      }
    }
  }
  if (slab != nullptr) {
    // ticket: the last CTA through adds the tile's slabs in segment order (threadFenceReduction pattern)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.locks + tile, 1) == nseg - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int r = warp; r < p.M; r += SM_WARPS) {
        const int ch = lane * 4;
        if (ch < nvalid) {
          const size_t off = (size_t)r * p.N + n_base + ch;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int z = 0; z < nseg; ++z) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p.c_tmp + (size_t)z * p.M * p.N + off));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
          uint2 o;
          o.x = pack2<T>(v.x, v.y);
          o.y = pack2<T>(v.z, v.w);
          *reinterpret_cast<uint2*>(cptr + off) = o;
        }
      }
      if (threadIdx.x == 0) p.locks[tile] = 0;
    }
  }
  __syncthreads();                                            // the reduce buffer becomes the rings of the next segment
  }  // segments
}

// stream-k launch shape of the small-batch kernel: number of CTAs (2 per SM, at least 8 chunks each) and the largest
// number of CTAs any tile is spread over (= fp32 slabs the caller must provide; 1 = none)
static void small_partition(int N, int K, int* grid_out, int* slabs_out) {
  const long long tiles = (N + 127) / 128, chunks = K / 64, U = tiles * chunks;
  long long G = std::min<long long>(2LL * num_sms(), std::max<long long>(1, U / 8));
  int slabs = 1;
  for (long long t = 0; t < tiles; ++t) {
    const long long cf = ((t * chunks + 1) * G - 1) / U, cl = ((t * chunks + chunks) * G - 1) / U;
    slabs = std::max<int>(slabs, (int)(cl - cf + 1));
  }
  *grid_out = (int)G;
  *slabs_out = slabs;
}

int marlin_small_plan(int M, int N, int K, int group_size) {
  (void)M; (void)group_size;
  int grid, slabs;
  small_partition(N, K, &grid, &slabs);
  return slabs;
}

template <typename T, int ZP, int MB>
static int launch_small(const SmallParams& p, int grid, cudaStream_t st) {
  using Cfg = SmallCfg<MB>;
  auto kern = marlin_w4a16_small_kernel<T, ZP, MB>;
  static thread_local uint64_t attr_done = 0;
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done |= 1ull << (dev & 63);
  }
  kern<<<grid, SM_THREADS, Cfg::SMEM, st>>>(p);
  return check_launch("marlin_w4a16_small_kernel");
}

// called by b200_gptq_marlin_gemm for 4-bit weights with M <= 32
int marlin_small_gemm(const void* a, const void* b_q, const void* scales, const void* zeros, void* c, float* c_tmp,
                      int* locks, int M, int N, int K, int num_groups, int has_zp, int dtype, int split_k,
                      cudaStream_t st) {
  const int gs = num_groups > 1 ? K / num_groups : -1;
  SmallParams p{};
  p.a = a; p.b_q = (const uint32_t*)b_q; p.scales = scales; p.zeros = (const uint32_t*)zeros; p.c = c;
  p.c_tmp = c_tmp; p.locks = locks; p.M = M; p.N = N; p.K = K;
  p.grouped = (gs > 0 && gs < K) ? 1 : 0;
  p.ktiles_per_group = p.grouped ? gs / 16 : 1;
  p.kpg_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == p.ktiles_per_group) p.kpg_shift = sh;
  (void)split_k;                                    // the partition is a pure function of (N, K, SM count)
  int grid, slabs;
  small_partition(N, K, &grid, &slabs);
  B200_CHECK(slabs == 1 || (c_tmp != nullptr && locks != nullptr),
             "the small-batch kernel needs the fp32 partial buffer [b200_marlin_gemm_plan(), M, N] and the zeroed lock workspace");
  p.chunks = K / 64;
  p.tiles = (N + 127) / 128;
  const bool bf = dtype == B200_BF16;
#define B200_SM(TT, ZZ, MM) return launch_small<TT, ZZ, MM>(p, grid, st)
  if (M <= 16) {
    if (bf) { if (has_zp) B200_SM(__nv_bfloat16, ZP_INT, 1); B200_SM(__nv_bfloat16, ZP_NONE, 1); }
    if (has_zp) B200_SM(__half, ZP_INT, 1);
    B200_SM(__half, ZP_NONE, 1);
  }
  if (bf) { if (has_zp) B200_SM(__nv_bfloat16, ZP_INT, 2); B200_SM(__nv_bfloat16, ZP_NONE, 2); }
  if (has_zp) B200_SM(__half, ZP_INT, 2);
  B200_SM(__half, ZP_NONE, 2);
#undef B200_SM
}

}  // namespace b200
