// marlin_gemm.cu — W4A16 GEMM in the Marlin weight format on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces gptq_marlin_gemm (kernels/quantization/gptq_marlin/gptq_marlin.cu:2247-2430 entry, :527-1764
// kernel) for the 4-bit GPTQ (uint4b8) and AWQ (uint4 + integer zero points) formats:
//     C[M,N] = A[M,K] . W,   W[k,n] = (q[k,n] - 8 | zp[g,n]) * s[g,n]     (fp32 accumulate)
// Not a port: the reference is an Ampere design (cp.async + ldmatrix + mma.sync.m16n8k16, weights
// dequantised straight into mma.sync B-fragment registers, M tiled in 64-row sub-problems that each
// re-stream W). tcgen05 has no register operands, so the problem is TRANSPOSED and restructured:
//
//   D^T[128 out-channels, tokens<=256] += Wt[128 ch, 64 k] . A^T        one CTA per (128-channel tile,
//                                                                        256-token block, k-split)
//   * A (activations) : TMA tensor-map load (SWIZZLE_128B) of a [tokens x 64] bf16 box = the MMA "B"
//                       operand (N = tokens, K-major). All tokens ride in ONE instruction (N <= 256), so W is
//                       dequantised once per CTA, not once per 64 rows of M.
//   * W (packed int4) : the Marlin tile rows of this channel tile are contiguous 1 KB runs -> cp.async.bulk
//                       into an 8-deep ring; four dequant warps read one uint4 per lane (= a lane's four
//                       mma.sync fragments in Marlin's layout), extract nibble PAIRS with lop3 in Marlin's
//                       interleaved order, subtract the bias / zero point exactly, multiply by the group scale
//                       (one rounding: bit-identical to the reference's w = T((q-8)*s)), and st.shared the
//                       result into a SWIZZLE_128B K-major tile = the MMA "A" operand (M = 128 channels).
//                       The lane->(row,k) mapping makes those stores bank-conflict free.
//   * MMA             : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=tokens, K=16) x4
//                       per stage, accumulating in TMEM; tcgen05.commit releases the stage / publishes the tile.
//   * epilogue        : the dequant warps tcgen05.ld their TMEM lane quadrant (lane = channel, column = token),
//                       convert and store C (or red.add into the fp32 split-k buffer).
#include "common.cuh"

#include <cuda.h>

#include <stdlib.h>
#include <type_traits>

namespace b200 {

static constexpr int MG_NT = 128;        // output channels per CTA  (UMMA M)
static constexpr int MG_KC = 64;         // k per stage (one 128-byte swizzle row of 16-bit elements)
static constexpr int MG_TOK = 256;       // max tokens per CTA       (UMMA N)
static constexpr int MG_MAX_STAGES = 8;  // (activation tile + dequantised weight tile) stages: 3 at 256 tokens, more below
static constexpr int MG_TEAMS = 4;       // dequant teams (4 warps each) working on interleaved chunks (TLP for the ALU chains)
static constexpr int MG_DQ_WARPS = 4 * MG_TEAMS;
static constexpr int MG_WARP_TMA = MG_DQ_WARPS;       // highest warp ids = highest issue priority
static constexpr int MG_WARP_MMA = MG_DQ_WARPS + 1;
static constexpr int MG_THREADS = (MG_DQ_WARPS + 2) * 32;
static constexpr int MG_W_BYTES = MG_NT * 128;        // 16 KB dequantised weight tile
// per-warp cp.async ring slot: 2 x 512 B packed words | 2 x 512 B scale vectors | 2 x 128 B zero points
static constexpr int MG_SLOT_BYTES = 2304;
static constexpr int MG_RING_DEPTH = 2;               // chunks in flight per warp (x 4 teams = 8 chunks ahead)
static constexpr int MG_SMEM_TOTAL = 226 * 1024;      // opt-in dynamic shared memory available to one CTA
static constexpr int MG_SMEM_FIXED = MG_DQ_WARPS * MG_RING_DEPTH * MG_SLOT_BYTES + 1024 /*barriers*/ + 1024 /*align*/;

struct MarlinParams {
  const uint32_t* b_q;   // [K/16, N*2] int32, Marlin layout
  const void* scales;    // [groups, N] T, Marlin-permuted
  const uint32_t* zeros; // [groups, N/8] int32 (AWQ) or nullptr
  void* c;               // [M, N] T
  float* c_tmp;          // [split_k, M, N] fp32 partial slabs (split-k only; no initialisation needed)
  int* locks;            // >= tiles ints, zero on entry, returned to zero (the reference's `workspace`)
  int M, N, K;
  int group_size;        // -1 = channel-wise (single scale row)
  int chunks_per_split;  // 64-wide k chunks handled by one CTA
  int split_k;
  int box_rows;          // rows of the activation TMA box (tokens rounded up to 16, <= 256)
  int grouped;           // 1: b_scales has one row per k-group (Marlin "grouped" permutation), 0: single row
  int rows_per_chunk;    // scale rows a 64-wide chunk spans (1, or 2 when group_size == 32)
  int chunks_per_group;  // 64-wide chunks per scale group (>= 1)
  int cpg_shift;         // log2(chunks_per_group) when it is a power of two, else -1
  int stages;            // act/weight pipeline depth (2..8), chosen from the token count
  int act_bytes;         // bytes of one activation stage (box_rows * 128, rounded up to 1024)
  int debug;             // B200_MARLIN_DEBUG (timing experiments only): 1 skip dequant math, 2 skip MMAs, 4 skip act TMA
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                               uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
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

// debug-only cycle attribution (B200_MARLIN_DEBUG & 16): CTA (0,0,0) accumulates, per role, the cycles spent
// in each mbarrier wait; read back with b200_debug_marlin_prof()
__device__ unsigned long long g_mg_prof[32];
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool prof, unsigned long long& acc) {
  if (!prof) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major: 1), [32,46) SBO>>4 (8 rows x 128 B = 1024),
//   [46,48) version = 1 (Blackwell), [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
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

// position of output column n (0..N) inside a Marlin-permuted scale row
// (aphrodite/quantization/utils/marlin_utils.py:172-196)
__device__ __forceinline__ int scale_pos(int n, bool grouped) {
  if (grouped) return (n & ~63) + 8 * (n & 7) + ((n & 63) >> 3);
  return (n & ~31) + 8 * ((n & 7) >> 1) + 2 * ((n & 31) >> 3) + (n & 1);
}

template <typename T, bool HAS_ZP>
__global__ void __launch_bounds__(MG_THREADS, 1)
marlin_w4a16_tc5_kernel(const __grid_constant__ CUtensorMap tmap_a, const MarlinParams p) {
  extern __shared__ uint8_t mg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mg_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NS = p.stages;
  uint8_t* act_s = smem;                                       // [NS][box_rows x 128 B]   (TMA, SWIZZLE_128B)
  uint8_t* w_s = act_s + (size_t)NS * p.act_bytes;             // [NS][128 x 128 B]        (dequantised weights)
  uint8_t* ring_s = w_s + (size_t)NS * MG_W_BYTES;             // [DQ_WARPS][DEPTH][slot]  (per-warp cp.async rings)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_s + (size_t)MG_DQ_WARPS * MG_RING_DEPTH * MG_SLOT_BYTES);
  // one "stage ready" barrier per stage: 1 arrival + tx bytes from the activation TMA and 4 arrivals from the
  // team that dequantised the weight tile (a single wait in the MMA issuer's loop instead of two)
  uint64_t* full_w = bars;                          // [NS]
  uint64_t* full_act = full_w;                      //  (same barriers)
  // "stage free" barriers, TWO per stage used alternately (use u of a stage signals empty[(u&1)][s]): a waiter
  // is then never two phases ahead of the barrier it tests, whatever the team / stage counts (a parity test
  // two phases ahead passes vacuously — the bug class of the attention ring)
  uint64_t* empty = full_w + MG_MAX_STAGES;         // [2][NS]  tcgen05.commit
  uint64_t* accum_full = empty + 2 * MG_MAX_STAGES; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = blockIdx.x * MG_NT;
  const int tok_base = blockIdx.y * MG_TOK;
  const int toks = min(MG_TOK, p.M - tok_base);
  const int n_mma = (toks + 15) & ~15;                       // UMMA N (multiple of 16, <= 256)
  const int nblk = min(2, (p.N - n_base) / 64);              // 16x64 Marlin blocks in this channel tile
  const int total_chunks = p.K / MG_KC;
  const int chunk0 = blockIdx.z * p.chunks_per_split;
  const int nchunks = min(p.chunks_per_split, total_chunks - chunk0);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_mma) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MG_MAX_STAGES; ++i) {
      mbar_init(&full_w[i], 5);
      mbar_init(&empty[i], 1);
      mbar_init(&empty[MG_MAX_STAGES + i], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  if (warp == MG_WARP_TMA) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const bool prof = (p.debug & 16) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  const long long t_role0 = clock64();

  // Warp roles. The two single-thread roles sit in the HIGHEST warp ids: the issue arbiter favours high warp
  // ids, and with the ALU-heavy dequant warps above them the TMA / MMA issuers were starved (measured:
  // 500-1200 cycles per issued TMA / 200 per MMA).
  if (nchunks > 0 && warp == MG_WARP_TMA) {
    // ===================== activation TMA producer =====================
    if (lane == 0) {
      // stage / use counters are advanced incrementally: a runtime `c % NS`, `c / NS` costs ~100 cycles of
      // dependent integer math per stage in these single-thread roles
      int s = 0;
      uint32_t use = 0;
      for (int c = 0; c < nchunks; ++c, s = (s + 1 == NS) ? 0 : s + 1, use += (s == 0)) {
        if (use > 0) mbar_wait_t(&empty[((use - 1) & 1u) * MG_MAX_STAGES + s], ((use - 1) >> 1) & 1u, prof, w0);
        if (p.debug & 4) { mbar_arrive(&full_act[s]); continue; }
        mbar_arrive_expect_tx(&full_act[s], (uint32_t)p.box_rows * 128u);  // TMA always moves the full box
        tma_load_2d(act_s + (size_t)s * p.act_bytes, &tmap_a, &full_act[s], (chunk0 + c) * MG_KC, tok_base);
      }
      if (prof) { g_mg_prof[0] = (unsigned long long)(clock64() - t_role0); g_mg_prof[1] = w0; g_mg_prof[2] = (unsigned long long)nchunks; }
    }
  } else if (nchunks > 0 && warp == MG_WARP_MMA) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A/B = T, both K-major, N, M = 128
      const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;   // 0 = F16, 1 = BF16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                             ((uint32_t)(MG_NT >> 4) << 24);
      int s = 0;
      uint32_t use = 0;
      for (int c = 0; c < nchunks; ++c, s = (s + 1 == NS) ? 0 : s + 1, use += (s == 0)) {
        const uint32_t par = use & 1u;
        mbar_wait_t(&full_w[s], par, prof, w1);
        tc_fence_after();
        const long long tm0 = prof ? clock64() : 0;
        const uint64_t a_desc = make_sw128_desc(smem_u32(w_s + (size_t)s * MG_W_BYTES));
        const uint64_t b_desc = make_sw128_desc(smem_u32(act_s + (size_t)s * p.act_bytes));
#pragma unroll
        for (int ks = 0; ks < MG_KC / 16; ++ks) {
          if (p.debug & 2) break;
          // advancing 16 k = 32 bytes inside the 128-byte swizzle row = +2 in the descriptor's (addr >> 4) field
          umma_f16(tmem_d, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u);
        }
        const long long tm1 = prof ? clock64() : 0;
        umma_commit(&empty[(use & 1u) * MG_MAX_STAGES + s]);   // frees act/w stage s when these MMAs retire
        if (prof) { w2 += (unsigned long long)(tm1 - tm0); w3 += (unsigned long long)(clock64() - tm1); }
      }
      umma_commit(accum_full);                   // accumulator complete
      if (prof) { g_mg_prof[4] = (unsigned long long)(clock64() - t_role0); g_mg_prof[5] = w0; g_mg_prof[6] = w1; g_mg_prof[7] = w2; g_mg_prof[16] = w3; }
    }
  } else if (nchunks > 0 && warp < MG_DQ_WARPS) {
    // ===================== dequant warps =====================
    // team = warp / 4 takes chunks team, team + TEAMS, ...; inside a team warp dq = warp % 4 owns k-tile dq of
    // the chunk. Every warp streams ITS OWN packed words, scale vectors and zero points with cp.async
    // (LDGSTS, 16 B per lane) into a private 4-deep ring: no producer warp, no TMA-unit time (the unit costs
    // ~400 cycles per operation and was the bottleneck when weights came by TMA / bulk copies), no barriers —
    // each lane reads back exactly the bytes it fetched.
    const int team = warp >> 2;
    const int dq = warp & 3;
    const int m = lane & 3, cq = lane >> 2;
    const bool grouped = p.grouped != 0;
    const T* sc = reinterpret_cast<const T*>(p.scales);
    uint32_t s2[2][8];    // per Marlin block: scale pairs {s,s} for column (j, b) at index 2j+b
    uint32_t off2[2][8];  // per column: 16-bit pair of MAGIC + (8 | zero point)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int e = 0; e < 8; ++e) { s2[nb][e] = 0; off2[nb][e] = DQ<T>::offset(8); }
    if (!grouped) {
      // channel-wise: one scale row (and zero-point row) for the whole k range, read once
      for (int nb = 0; nb < nblk; ++nb) {
        const int col0 = n_base + nb * 64 + cq;
        uint32_t zword = 0;
        if (HAS_ZP) zword = p.zeros[(n_base + nb * 64) / 8 + cq];
#pragma unroll
        for (int e = 0; e < 8; ++e) {      // e = 2j + b  <->  column col0 + 16j + 8b
          const int n = col0 + 16 * (e >> 1) + 8 * (e & 1);
          const T sv = sc[scale_pos(n, false)];
          const uint32_t s16 = *reinterpret_cast<const uint16_t*>(&sv);
          s2[nb][e] = s16 | (s16 << 16);
          if (HAS_ZP) off2[nb][e] = DQ<T>::offset((int)((zword >> (4 * (((e & 1) << 2) | (e >> 1)))) & 0xFu));
        }
      }
    }
    const uint32_t ring = smem_u32(ring_s) + (uint32_t)warp * (MG_RING_DEPTH * MG_SLOT_BYTES) + (uint32_t)lane * 16u;
    const uint32_t w_addr = smem_u32(w_s);
    const int row_words = p.N * 2;                      // int32 per 16-row k-tile of the Marlin matrix
    // global sources of this lane: words 4*lane..4*lane+3 of block nb, k-tile (4*chunk + dq)
    const uint32_t* wsrc = p.b_q + (size_t)(n_base / 64) * 128 + (size_t)lane * 4;
    auto prefetch = [&](int c) {       // issue this lane's cp.async for chunk c into its ring slot, one group
      if (c < nchunks) {
        const uint32_t slot = ring + (uint32_t)((c / MG_TEAMS) % MG_RING_DEPTH) * MG_SLOT_BYTES;
        const uint32_t* src = wsrc + (size_t)((chunk0 + c) * 4 + dq) * row_words;
        cp_async16(slot, src);
        if (nblk > 1) cp_async16(slot + 512, src + 128);
        if (grouped) {
          const int g = p.rows_per_chunk == 2 ? (chunk0 + c) * 2 + (dq >> 1)
                                             : (p.cpg_shift >= 0 ? (chunk0 + c) >> p.cpg_shift : (chunk0 + c) / p.chunks_per_group);
          const uint8_t* ssrc = reinterpret_cast<const uint8_t*>(p.scales) + ((size_t)g * p.N + n_base + 8 * cq) * 2;
          cp_async16(slot + 1024, ssrc);
          if (nblk > 1) cp_async16(slot + 1536, ssrc + 128);
          if (HAS_ZP) {
            const uint32_t* zsrc = p.zeros + (size_t)g * (p.N / 8) + n_base / 8 + cq;
            cp_async4(slot + 2048 - (uint32_t)lane * 12u, zsrc);            // 4-byte slots: lane * 4
            if (nblk > 1) cp_async4(slot + 2176 - (uint32_t)lane * 12u, zsrc + 8);
          }
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < MG_RING_DEPTH; ++d) prefetch(team + d * MG_TEAMS);

    int s = team % NS;                 // stage and use count of chunk c, advanced by MG_TEAMS per iteration
    uint32_t use = (uint32_t)(team / NS);
    for (int c = team; c < nchunks; c += MG_TEAMS) {
      cp_async_wait<MG_RING_DEPTH - 1>();           // this lane's copies for chunk c have landed
      const uint32_t slot = ring + (uint32_t)((c / MG_TEAMS) % MG_RING_DEPTH) * MG_SLOT_BYTES;
      uint4 q[2];
      q[0] = lds128(slot);
      q[1] = (nblk > 1) ? lds128(slot + 512) : make_uint4(0, 0, 0, 0);
      if (grouped) {
        // Marlin's grouped scale permutation puts this lane's 8 scales (columns 16j + 8b + cq of a block) in 16
        // contiguous bytes at position 8*cq of the block's 64-entry row; the 8 zero points are one int32
        for (int nb = 0; nb < nblk; ++nb) {
          const uint4 sv = lds128(slot + 1024 + nb * 512);
          const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
          uint32_t zword = 0;
          if (HAS_ZP) zword = lds32(slot + 2048 + nb * 128 - (uint32_t)lane * 12u);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t s16 = (sw[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            s2[nb][e] = s16 | (s16 << 16);
            if (HAS_ZP) off2[nb][e] = DQ<T>::offset((int)((zword >> (4 * (((e & 1) << 2) | (e >> 1)))) & 0xFu));
          }
        }
      }
      {                                                   // stage's previous MMAs must have retired
        if (use > 0) mbar_wait_t(&empty[((use - 1) & 1u) * MG_MAX_STAGES + s], ((use - 1) >> 1) & 1u, prof && warp == 0 && lane == 0, w1);
      }
      const uint32_t wt = w_addr + (uint32_t)s * MG_W_BYTES;
      const uint32_t a0 = (uint32_t)(((2 * dq) ^ cq) << 4) + 4u * m;       // k = 16dq + 2m (+1)
      const uint32_t a1 = (uint32_t)(((2 * dq + 1) ^ cq) << 4) + 4u * m;   // k = 16dq + 8 + 2m (+1)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        if (nb < nblk && !(p.debug & 1)) {
          const uint32_t wq[4] = {q[nb].x, q[nb].y, q[nb].z, q[nb].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w = wq[j];
            const int row0 = nb * 64 + 16 * j + cq;
            const uint32_t r0 = wt + row0 * 128;
            const uint32_t r1 = r0 + 8 * 128;
            const uint32_t x00 = lop3_and_or(w, 0x000f000fu, DQ<T>::MAGIC);
            const uint32_t x01 = lop3_and_or(w >> 4, 0x000f000fu, DQ<T>::MAGIC);
            const uint32_t x10 = lop3_and_or(w >> 8, 0x000f000fu, DQ<T>::MAGIC);
            const uint32_t x11 = lop3_and_or(w >> 12, 0x000f000fu, DQ<T>::MAGIC);
            sts32(r0 + a0, DQ<T>::sub_mul(x00, off2[nb][2 * j], s2[nb][2 * j]));
            sts32(r0 + a1, DQ<T>::sub_mul(x01, off2[nb][2 * j], s2[nb][2 * j]));
            sts32(r1 + a0, DQ<T>::sub_mul(x10, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
            sts32(r1 + a1, DQ<T>::sub_mul(x11, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
          }
        }
      }
      if (!(p.debug & 8)) fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_w[s]);
      prefetch(c + MG_RING_DEPTH * MG_TEAMS);     // refill the slot just consumed (its words are in registers)
      s += MG_TEAMS;
      while (s >= NS) { s -= NS; ++use; }
    }
    cp_async_wait<0>();
    if (prof && warp == 0 && lane == 0) { g_mg_prof[12] = (unsigned long long)(clock64() - t_role0); g_mg_prof[13] = w0; g_mg_prof[14] = w1; }

    // ===================== epilogue: TMEM -> C =====================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quad = warp & 3;                            // TMEM lanes 32*quad .. +31 belong to this warp
    const int ch = n_base + quad * 32 + lane;
    const bool ch_ok = ch < p.N;
    T* cptr = reinterpret_cast<T*>(p.c);
    float* slab = p.split_k > 1 ? p.c_tmp + (size_t)blockIdx.z * p.M * p.N : nullptr;
    // the MG_TEAMS warps that share a lane quadrant interleave 32-column slabs
    for (int col0 = team * 32; col0 < n_mma; col0 += 32 * MG_TEAMS) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, v);
      if (ch_ok) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const int tok = tok_base + col0 + t;
          if (col0 + t < toks) {
            const float f = __uint_as_float(v[t]);
            if (slab != nullptr) slab[(size_t)tok * p.N + ch] = f;
            else cptr[(size_t)tok * p.N + ch] = from_f32<T>(f);
          }
        }
      }
    }
    if (slab != nullptr) {
      // split-k: every split stores its fp32 partial in its own slab, then all splits of the tile meet on the
      // tile's lock (the grid fits the GPU when the plan splits k, so every CTA is resident) and EACH reduces
      // an interleaved 1/split_k share of the token rows over the slabs in a fixed order: deterministic, no
      // atomics on data, no memset, no second kernel. The last CTA through returns the lock to zero — the
      // reference's workspace contract (kernels/torch_bindings.cpp:167-176).
      constexpr int EPI = MG_DQ_WARPS * 32;
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      int* lock = p.locks + tile;
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
      if (threadIdx.x == 0) {
        atomicAdd(lock, 1);
        while (atomicAdd(lock, 0) < p.split_k) __nanosleep(64);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
      __threadfence();
      const int nvalid = min(MG_NT, p.N - n_base);               // channels of this tile (64 or 128)
      if (lane * 4 < nvalid) {
        for (int r = blockIdx.z + p.split_k * warp; r < toks; r += p.split_k * MG_DQ_WARPS) {
          const size_t off = (size_t)(tok_base + r) * p.N + n_base + lane * 4;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int z = 0; z < p.split_k; ++z) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p.c_tmp + (size_t)z * p.M * p.N + off));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          uint2 o;
          o.x = pack2<T>(acc.x, acc.y);
          o.y = pack2<T>(acc.z, acc.w);
          *reinterpret_cast<uint2*>(cptr + off) = o;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
      if (threadIdx.x == 0) {
        if (atomicAdd(lock, 1) == 2 * p.split_k - 1) atomicExch(lock, 0);
      }
    }
    if (prof && warp == 0 && lane == 0) g_mg_prof[15] = (unsigned long long)(clock64() - t_role0);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == MG_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// ---- host -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int plan_split_k(int M, int N, int K, int group_size) {
  const int tiles = ((N + MG_NT - 1) / MG_NT) * ((M + MG_TOK - 1) / MG_TOK);
  const int chunks = K / MG_KC;
  int split = 1;
  const int sms = num_sms();
  if ((M + MG_TOK - 1) / MG_TOK > 32) return 1;   // lock workspace (N/64*16 ints) covers <= 32 token blocks
  while (tiles * split * 2 <= sms && chunks / (split * 2) >= 8) split *= 2;
  // keep every split on a group boundary
  const int gchunks = group_size > MG_KC ? group_size / MG_KC : 1;
  while (split > 1 && ((chunks + split - 1) / split) % gchunks != 0) split /= 2;
  return split;
}

template <typename T, bool HAS_ZP>
static int launch_marlin(const CUtensorMap& tmap, const MarlinParams& p, cudaStream_t st) {
  auto kern = marlin_w4a16_tc5_kernel<T, HAS_ZP>;
  static thread_local uint64_t attr_done = 0;
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM_TOTAL));
    attr_done |= 1ull << (dev & 63);
  }
  dim3 grid((p.N + MG_NT - 1) / MG_NT, (p.M + MG_TOK - 1) / MG_TOK, p.split_k);
  const size_t smem = (size_t)p.stages * (p.act_bytes + MG_W_BYTES) + MG_SMEM_FIXED;
  kern<<<grid, MG_THREADS, smem, st>>>(tmap, p);
  return check_launch("marlin_w4a16_tc5_kernel");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_debug_marlin_prof(unsigned long long* out32) {
  B200_CUDA_OK(cudaDeviceSynchronize());
  B200_CUDA_OK(cudaMemcpyFromSymbol(out32, g_mg_prof, sizeof(unsigned long long) * 32));
  return 0;
}

extern "C" int b200_marlin_gemm_plan(int size_m, int size_n, int size_k, int num_groups) {
  const int gs = num_groups > 1 ? size_k / num_groups : -1;
  if (size_m <= 0 || size_n <= 0 || size_k < MG_KC) return 1;
  return plan_split_k(size_m, size_n, size_k, gs);
}

extern "C" int b200_gptq_marlin_gemm(const void* a, const void* b_q_weight, const void* b_scales,
                                     const void* b_zeros, void* c, float* c_tmp, int32_t* workspace,
                                     int size_m, int size_n, int size_k, int num_groups, int num_bits,
                                     int has_zp, int dtype, int split_k, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "gpt_marlin_gemm only supports bfloat16 and float16");
  B200_CHECK(num_bits == 4, "b200 marlin gemm: only 4-bit weights (uint4b8 / uint4) are implemented");
  B200_CHECK(size_n % 64 == 0, "size_n = " + std::to_string(size_n) + ", is not divisible by min_thread_n = 64");
  B200_CHECK(size_k % MG_KC == 0, "size_k = " + std::to_string(size_k) + " is not divisible by 64");
  B200_CHECK(num_groups >= 1 && size_k % num_groups == 0, "size_k is not divisible by the number of scale groups");
  const int gs = num_groups > 1 ? size_k / num_groups : -1;
  B200_CHECK(gs == -1 || gs == size_k || gs == 32 || (gs % 64 == 0), "unsupported group size " + std::to_string(gs) +
             " (supported: -1 / 32 / multiples of 64)");
  B200_CHECK(!has_zp || b_zeros != nullptr, "has_zp requires b_zeros");
  B200_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b_q_weight) & 15) == 0,
             "a and b_q_weight must be 16-byte aligned");
  if (size_m == 0) return 0;
  if (split_k <= 0) split_k = plan_split_k(size_m, size_n, size_k, gs);
  B200_CHECK(split_k == 1 || (c_tmp != nullptr && workspace != nullptr),
             "split-k needs the fp32 partial buffer [split_k, M, N] and the zeroed lock workspace");
  {
    const int tiles = ((size_n + MG_NT - 1) / MG_NT) * ((size_m + MG_TOK - 1) / MG_TOK);
    while (split_k > 1 && tiles * split_k > num_sms()) --split_k;   // all splits of a tile must be co-resident
    const int chunks_total = size_k / MG_KC;
    while (split_k > 1 && (split_k - 1) * ((chunks_total + split_k - 1) / split_k) >= chunks_total) --split_k;
  }
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");

  const int toks = size_m < MG_TOK ? size_m : MG_TOK;
  const int box_rows = (toks + 15) & ~15;
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)size_k, (cuuint64_t)size_m};
  const cuuint64_t gstride[1] = {(cuuint64_t)size_k * 2};
  const cuuint32_t box[2] = {(cuuint32_t)MG_KC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         2, const_cast<void*>(a), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  MarlinParams p{};
  p.b_q = (const uint32_t*)b_q_weight; p.scales = b_scales; p.zeros = (const uint32_t*)b_zeros;
  p.c = c; p.c_tmp = c_tmp; p.locks = workspace; p.M = size_m; p.N = size_n; p.K = size_k; p.group_size = gs;
  p.split_k = split_k;
  p.box_rows = box_rows;
  {
    const char* dbg = getenv("B200_MARLIN_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.act_bytes = (box_rows * 128 + 1023) & ~1023;
  // shared-memory plan: as many MMA stages (activation tile + dequantised weight tile) as fit, 2..8
  p.stages = (MG_SMEM_TOTAL - MG_SMEM_FIXED) / (p.act_bytes + MG_W_BYTES);
  if (p.stages > MG_MAX_STAGES) p.stages = MG_MAX_STAGES;
  p.grouped = (gs > 0 && gs < size_k) ? 1 : 0;
  p.rows_per_chunk = (gs > 0 && gs < MG_KC) ? MG_KC / gs : 1;
  p.chunks_per_group = (gs > MG_KC) ? gs / MG_KC : 1;
  p.cpg_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == p.chunks_per_group) p.cpg_shift = sh;
  const int chunks = size_k / MG_KC;
  p.chunks_per_split = (chunks + split_k - 1) / split_k;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == B200_BF16)
    rc = has_zp ? launch_marlin<__nv_bfloat16, true>(tmap, p, st) : launch_marlin<__nv_bfloat16, false>(tmap, p, st);
  else
    rc = has_zp ? launch_marlin<__half, true>(tmap, p, st) : launch_marlin<__half, false>(tmap, p, st);
  return rc;
}
