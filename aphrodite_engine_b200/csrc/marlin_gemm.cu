// marlin_gemm.cu — W4A16 GEMM in the Marlin weight format on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces gptq_marlin_gemm (kernels/quantization/gptq_marlin/gptq_marlin.cu:2247-2430 entry, :527-1764
// kernel) for the GPTQ (uint4b8 / uint8b128), AWQ (uint4 / uint8 + integer zero points) and HQQ (uint4 + fp16
// zero points) formats, and marlin_gemm_moe (kernels/moe/marlin_moe_ops.cu:1482-1546) as a grouped launch:
//     C[M,N] = A[M,K] . W,   W[k,n] = (q[k,n] - bias | zp[g,n]) * s[g,n]     (fp32 accumulate)
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
#include "marlin_dq.cuh"
#include "tc5.cuh"

#include <stdlib.h>

#include <algorithm>
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
static constexpr int MG_ISSUERS = 2;                  // MMA issuer warps, interleaved over the chunks (see the issuer role)
static constexpr int MG_THREADS = (MG_DQ_WARPS + 1 + MG_ISSUERS) * 32;
static constexpr int MG_W_BYTES = MG_NT * 128;        // 16 KB dequantised weight tile
static constexpr int MG_RING_DEPTH = 2;               // chunks in flight per warp (x 4 teams = 8 chunks ahead)
static constexpr int MG_SMEM_TOTAL = 226 * 1024;      // opt-in dynamic shared memory available to one CTA

// per-warp cp.async ring slot: packed words of 2 Marlin blocks | 2 x 512 B scale vectors | zero points
//   words: 4-bit 2 x 512 B (one uint4 per lane and block), 8-bit 2 x 1024 B (two uint4 per lane and block)
//   zero points: int4 2 x 128 B (one word per lane), int8 2 x 256 B (two words per lane), float 2 x 512 B
template <int BITS, int ZP> struct MGCfg {
  static constexpr int WORD_BYTES = BITS == 4 ? 1024 : 2048;
  static constexpr int SC_OFF = WORD_BYTES;
  static constexpr int ZP_OFF = SC_OFF + 1024;
  static constexpr int ZP_BYTES = ZP == ZP_FLOAT ? 1024 : (BITS == 4 ? 256 : 512);
  static constexpr int SLOT = ZP_OFF + ZP_BYTES;
  static constexpr int RING = MG_DQ_WARPS * MG_RING_DEPTH * SLOT;
  static constexpr int FIXED = RING + 1024 /*barriers*/ + 1024 /*align*/;
};

struct MarlinParams {
  const uint32_t* b_q;   // [K/16, N*16/pack] int32, Marlin layout (x num_experts for the grouped launch)
  const void* scales;    // [groups, N] T, Marlin-permuted
  const void* zeros;     // [groups, N/pack] int32 (AWQ), [groups, N] T (HQQ) or nullptr
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
  // grouped (mixture-of-experts) launch: blockIdx.y enumerates (expert, tile of `tile_rows` sorted rows)
  const int* expert_offsets;   // [E+1] first sorted row of every expert (block-size padded), device
  const int* sorted_ids;       // [cap] row of C (and of the un-gathered A) behind every sorted position
  const float* topk_weights;   // [M*topk] or nullptr: multiply row r of C by topk_weights[r]
  int num_experts;
  int valid_rows;              // M * topk: sorted ids >= this are padding
  int tile_rows;               // sorted rows per CTA (= box_rows)
  long long expert_words;      // int32 per expert in b_q
  long long expert_scales;     // scale elements per expert
};

// debug-only cycle attribution (B200_MARLIN_DEBUG & 16): CTA (0,0,0) accumulates, per role, the cycles spent
// in each mbarrier wait; read back with b200_debug_marlin_prof()
__device__ unsigned long long g_mg_prof[32];
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool prof, unsigned long long& acc) {
  if (!prof) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

// RING = true : packed words / scales / zero points travel global -> per-warp cp.async ring -> registers
// RING = false: they are loaded straight into registers one iteration (MG_TEAMS chunks) ahead; the 72+ KB of ring
//               become one more activation + weight stage (4 instead of 3 at 256 tokens)
template <typename T, int ZP, int BITS, bool MOE, bool RING>
__global__ void __launch_bounds__(MG_THREADS, 1)
marlin_w4a16_tc5_kernel(const __grid_constant__ CUtensorMap tmap_a, const MarlinParams p) {
  using Cfg = MGCfg<BITS, ZP>;
  constexpr int SLOT = Cfg::SLOT;
  constexpr int BIAS = BITS == 4 ? 8 : 128;
  constexpr int BLK_WORDS = BITS == 4 ? 128 : 256;      // int32 of one 16x64 Marlin block
  extern __shared__ uint8_t mg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mg_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NS = p.stages;
  uint8_t* act_s = smem;                                       // [NS][box_rows x 128 B]   (TMA, SWIZZLE_128B)
  uint8_t* w_s = act_s + (size_t)NS * p.act_bytes;             // [NS][128 x 128 B]        (dequantised weights)
  uint8_t* ring_s = w_s + (size_t)NS * MG_W_BYTES;             // [DQ_WARPS][DEPTH][slot]  (per-warp cp.async rings)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_s + (size_t)(RING ? Cfg::RING : 0));
  // one "stage ready" barrier per stage: 1 arrival + tx bytes from the activation TMA and 4 arrivals from the
  // team that dequantised the weight tile (a single wait in the MMA issuer's loop instead of two)
  uint64_t* full_w = bars;                          // [NS]
  uint64_t* full_act = full_w;                      //  (same barriers)
  // "stage free" barriers, TWO per stage used alternately (use u of a stage signals empty[(u&1)][s]): a waiter
  // is then never two phases ahead of the barrier it tests, whatever the team / stage counts (a parity test
  // two phases ahead passes vacuously — the bug class of the attention ring)
  uint64_t* empty = full_w + MG_MAX_STAGES;         // [2][NS]  tcgen05.commit
  uint64_t* accum_full = empty + 2 * MG_MAX_STAGES; // [1]  one commit per issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
  uint32_t* turn = tmem_slot + 1;                   // next chunk whose MMAs may be issued (orders the two issuers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = blockIdx.x * MG_NT;
  int tok_base = blockIdx.y * MG_TOK;                        // first row of this CTA in the TMA'd activation matrix
  int toks = min(MG_TOK, p.M - tok_base);
  const uint32_t* b_q = p.b_q;
  const T* sc = reinterpret_cast<const T*>(p.scales);
  if constexpr (MOE) {
    // blockIdx.y -> (expert, tile of its sorted-row segment). Uniform over the CTA, so surplus CTAs of the
    // upper-bound grid leave before any barrier / TMEM allocation.
    int t = blockIdx.y, e = 0, seg0 = 0, len = 0;
    for (; e < p.num_experts; ++e) {
      seg0 = __ldg(p.expert_offsets + e);
      len = __ldg(p.expert_offsets + e + 1) - seg0;
      const int nt = (len + p.tile_rows - 1) / p.tile_rows;
      if (t < nt) break;
      t -= nt;
    }
    if (e >= p.num_experts) return;
    tok_base = seg0 + t * p.tile_rows;
    toks = min(p.tile_rows, len - t * p.tile_rows);
    b_q += (size_t)e * p.expert_words;
    sc += (size_t)e * p.expert_scales;
  }
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
    mbar_init(accum_full, MG_ISSUERS);
    *turn = 0;
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
    if (elect_one()) {
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
  } else if (nchunks > 0 && warp >= MG_WARP_MMA) {
    // ===================== MMA issuers =====================
    // TWO issuer threads (one per warp), issuer q taking chunks q, q + 2, ...: the per-stage serial chain of one
    // issuer (mbarrier wait -> descriptors -> 4 UTCHMMA -> commit, ~500 cycles whatever N) was the limiter up to
    // 128 tokens and above the 512-cycle tensor floor at 256; two chains overlap. All MMAs accumulate into the same
    // TMEM tile and the tensor pipe executes them in issue order, so the ISSUE order is kept equal to the chunk
    // order with a shared-memory turn counter (a spin of a few tens of cycles): the accumulator-initialising MMA is
    // first, and fp32 accumulation order — hence the result — is identical from run to run; only the waits,
    // descriptor set-up and commits of the two issuers overlap. tcgen05.commit tracks the MMAs of the executing
    // thread, so each stage is released by the issuer that consumed it, and `accum_full` collects one commit per issuer.
    const int iq = warp - MG_WARP_MMA;
    if (elect_one()) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A/B = T, both K-major, N, M = 128
      const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;   // 0 = F16, 1 = BF16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                             ((uint32_t)(MG_NT >> 4) << 24);
      const bool skip_mma = (p.debug & 2) != 0;
      const bool profq = prof && iq == 0;
      int s = iq % NS;
      uint32_t use = (uint32_t)(iq / NS);
      const uint32_t turn_addr = smem_u32(turn);
      for (int c = iq; c < nchunks; c += MG_ISSUERS) {
        const uint32_t par = use & 1u;
        mbar_wait_t(&full_w[s], par, profq, w1);
        tc_fence_after();
        const long long tm0 = profq ? clock64() : 0;
        const uint64_t a_desc = make_sw128_desc(smem_u32(w_s + (size_t)s * MG_W_BYTES));
        const uint64_t b_desc = make_sw128_desc(smem_u32(act_s + (size_t)s * p.act_bytes));
        {                                           // my turn: every earlier chunk's MMAs have been issued
          uint32_t t;
          do {
            asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(t) : "r"(turn_addr) : "memory");
          } while (t != (uint32_t)c);
        }
#pragma unroll
        for (int ks = 0; ks < MG_KC / 16; ++ks) {
          if (skip_mma) break;
          // advancing 16 k = 32 bytes inside the 128-byte swizzle row = +2 in the descriptor's (addr >> 4) field
          umma_f16(tmem_d, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, (c > 0 || ks > 0) ? 1u : 0u);
        }
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(turn_addr), "r"((uint32_t)c + 1u) : "memory");
        const long long tm1 = profq ? clock64() : 0;
        umma_commit(&empty[(use & 1u) * MG_MAX_STAGES + s]);   // frees act/w stage s when these MMAs retire
        if (profq) { w2 += (unsigned long long)(tm1 - tm0); w3 += (unsigned long long)(clock64() - tm1); }
        s += MG_ISSUERS;
        while (s >= NS) { s -= NS; ++use; }
      }
      umma_commit(accum_full);                   // this issuer's share of the accumulator is complete
      if (profq) { g_mg_prof[4] = (unsigned long long)(clock64() - t_role0); g_mg_prof[5] = w0; g_mg_prof[6] = w1; g_mg_prof[7] = w2; g_mg_prof[16] = w3; }
    }
  } else if (nchunks > 0 && warp < MG_DQ_WARPS) {
    // ===================== dequant warps =====================
    // team = warp / 4 takes chunks team, team + TEAMS, ...; inside a team warp dq = warp % 4 owns k-tile dq of
    // the chunk. Every warp streams ITS OWN packed words, scale vectors and zero points with cp.async
    // (LDGSTS, 16 B per lane) into a private 2-deep ring: no producer warp, no TMA-unit time (the unit costs
    // ~400 cycles per operation and was the bottleneck when weights came by TMA / bulk copies), no barriers —
    // each lane reads back exactly the bytes it fetched.
    using W = WDQ<T, BITS, ZP>;
    const int team = warp >> 2;
    const int dq = warp & 3;
    const int m = lane & 3, cq = lane >> 2;
    const bool grouped = p.grouped != 0;
    uint32_t s2[2][8];    // per Marlin block: scale pairs {s,s} for column (j, b) at index 2j+b
    uint32_t off2[2][8];  // per column: offset word of WDQ (MAGIC + bias | zero point, or the float zero point pair)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int e = 0; e < 8; ++e) { s2[nb][e] = 0; off2[nb][e] = ZP == ZP_FLOAT ? 0u : W::offset_of(BIAS); }
    if (!grouped) {
      // channel-wise: one scale row (and zero-point row) for the whole k range, read once
      for (int nb = 0; nb < nblk; ++nb) {
        const int col0 = n_base + nb * 64 + cq;
        uint32_t z0 = 0, z1 = 0;
        if constexpr (ZP == ZP_INT) {
          const uint32_t* zr = reinterpret_cast<const uint32_t*>(p.zeros);
          if constexpr (BITS == 4) z0 = zr[(n_base + nb * 64) / 8 + cq];
          else { z0 = zr[(n_base + nb * 64) / 4 + 2 * cq]; z1 = zr[(n_base + nb * 64) / 4 + 2 * cq + 1]; }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {      // e = 2j + b  <->  column col0 + 16j + 8b
          const int n = col0 + 16 * (e >> 1) + 8 * (e & 1);
          const T sv = sc[scale_pos(n, false)];
          const uint32_t s16 = *reinterpret_cast<const uint16_t*>(&sv);
          s2[nb][e] = s16 | (s16 << 16);
          if constexpr (ZP == ZP_INT) off2[nb][e] = W::offset_of(zp_code<BITS>(e, z0, z1));
          if constexpr (ZP == ZP_FLOAT) {
            const uint32_t z16 = reinterpret_cast<const uint16_t*>(p.zeros)[scale_pos(n, false)];
            off2[nb][e] = z16 | (z16 << 16);
          }
        }
      }
    }
    const uint32_t ring = smem_u32(ring_s) + (uint32_t)warp * (MG_RING_DEPTH * SLOT) + (uint32_t)lane * 16u;
    const uint32_t w_addr = smem_u32(w_s);
    const int row_words = (p.N / 64) * BLK_WORDS;        // int32 per 16-row k-tile of the Marlin matrix
    // global sources of this lane: its words of block nb, k-tile (4*chunk + dq)
    const uint32_t* wsrc = b_q + (size_t)(n_base / 64) * BLK_WORDS + (size_t)lane * (BLK_WORDS / 32);
    // group row of chunk c for k-tile dq
    auto group_of = [&](int c) {
      return p.rows_per_chunk == 2 ? (chunk0 + c) * 2 + (dq >> 1)
                                   : (p.cpg_shift >= 0 ? (chunk0 + c) >> p.cpg_shift : (chunk0 + c) / p.chunks_per_group);
    };
    auto prefetch = [&](int c) {       // issue this lane's cp.async for chunk c into its ring slot, one group
      if (c < nchunks) {
        const uint32_t slot = ring + (uint32_t)((c / MG_TEAMS) % MG_RING_DEPTH) * SLOT;
        const uint32_t* src = wsrc + (size_t)((chunk0 + c) * 4 + dq) * row_words;
        if constexpr (BITS == 4) {
          cp_async16(slot, src);
          if (nblk > 1) cp_async16(slot + 512, src + BLK_WORDS);
        } else {                          // 8 words per lane and block: two conflict-free 512-byte planes
          cp_async16(slot, src);
          cp_async16(slot + 512, src + 4);
          if (nblk > 1) { cp_async16(slot + 1024, src + BLK_WORDS); cp_async16(slot + 1536, src + BLK_WORDS + 4); }
        }
        if (grouped) {
          const int g = group_of(c);
          const uint8_t* ssrc = reinterpret_cast<const uint8_t*>(sc) + ((size_t)g * p.N + n_base + 8 * cq) * 2;
          cp_async16(slot + Cfg::SC_OFF, ssrc);
          if (nblk > 1) cp_async16(slot + Cfg::SC_OFF + 512, ssrc + 128);
          if constexpr (ZP == ZP_INT && BITS == 4) {
            const uint32_t* zsrc = reinterpret_cast<const uint32_t*>(p.zeros) + (size_t)g * (p.N / 8) + n_base / 8 + cq;
            cp_async4(slot + Cfg::ZP_OFF - (uint32_t)lane * 12u, zsrc);            // 4-byte slots: lane * 4
            if (nblk > 1) cp_async4(slot + Cfg::ZP_OFF + 128 - (uint32_t)lane * 12u, zsrc + 8);
          }
          if constexpr (ZP == ZP_INT && BITS == 8) {
            const uint32_t* zsrc = reinterpret_cast<const uint32_t*>(p.zeros) + (size_t)g * (p.N / 4) + n_base / 4 + 2 * cq;
            cp_async8(slot + Cfg::ZP_OFF - (uint32_t)lane * 8u, zsrc);             // 8-byte slots: lane * 8
            if (nblk > 1) cp_async8(slot + Cfg::ZP_OFF + 256 - (uint32_t)lane * 8u, zsrc + 16);
          }
          if constexpr (ZP == ZP_FLOAT) {
            const uint8_t* zsrc = reinterpret_cast<const uint8_t*>(p.zeros) + ((size_t)g * p.N + n_base + 8 * cq) * 2;
            cp_async16(slot + Cfg::ZP_OFF, zsrc);
            if (nblk > 1) cp_async16(slot + Cfg::ZP_OFF + 512, zsrc + 128);
          }
        }
      }
      cp_async_commit();
    };
    // register-prefetch variant: the same bytes, loaded with ld.global.nc one iteration ahead
    uint4 qn[2][BITS / 4], sn[2], zn[2];
    auto fetch = [&](int c) {
      if (c < nchunks) {
        const uint32_t* src = wsrc + (size_t)((chunk0 + c) * 4 + dq) * row_words;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int pl = 0; pl < BITS / 4; ++pl)
            qn[nb][pl] = (nb < nblk) ? ldg_stream128(src + nb * BLK_WORDS + pl * 4) : make_uint4(0, 0, 0, 0);
        if (grouped) {
          const int g = group_of(c);
          const uint8_t* ssrc = reinterpret_cast<const uint8_t*>(sc) + ((size_t)g * p.N + n_base + 8 * cq) * 2;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            sn[nb] = (nb < nblk) ? ldg_stream128(ssrc + nb * 128) : make_uint4(0, 0, 0, 0);
            zn[nb] = make_uint4(0, 0, 0, 0);
            if (nb < nblk) {
              if constexpr (ZP == ZP_INT && BITS == 4)
                zn[nb].x = __ldg(reinterpret_cast<const uint32_t*>(p.zeros) + (size_t)g * (p.N / 8) + n_base / 8 + cq + nb * 8);
              if constexpr (ZP == ZP_INT && BITS == 8) {
                const uint2 zz = __ldg(reinterpret_cast<const uint2*>(
                    reinterpret_cast<const uint32_t*>(p.zeros) + (size_t)g * (p.N / 4) + n_base / 4 + 2 * cq + nb * 16));
                zn[nb].x = zz.x; zn[nb].y = zz.y;
              }
              if constexpr (ZP == ZP_FLOAT)
                zn[nb] = ldg_stream128(reinterpret_cast<const uint8_t*>(p.zeros) + ((size_t)g * p.N + n_base + 8 * cq) * 2 + nb * 128);
            }
          }
        }
      }
    };
    if constexpr (RING) {
#pragma unroll
      for (int d = 0; d < MG_RING_DEPTH; ++d) prefetch(team + d * MG_TEAMS);
    } else {
      fetch(team);
    }

    int s = team % NS;                 // stage and use count of chunk c, advanced by MG_TEAMS per iteration
    uint32_t use = (uint32_t)(team / NS);
    for (int c = team; c < nchunks; c += MG_TEAMS) {
      uint4 q[2][BITS / 4];                         // [block][plane]
      uint4 sraw[2], zraw[2];
      uint32_t slot = 0;
      if constexpr (RING) {
        cp_async_wait<MG_RING_DEPTH - 1>();           // this lane's copies for chunk c have landed
        slot = ring + (uint32_t)((c / MG_TEAMS) % MG_RING_DEPTH) * SLOT;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int pl = 0; pl < BITS / 4; ++pl)
            q[nb][pl] = (nb < nblk) ? lds128(slot + (uint32_t)(nb * (BITS / 4) + pl) * 512u) : make_uint4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
          for (int pl = 0; pl < BITS / 4; ++pl) q[nb][pl] = qn[nb][pl];
          sraw[nb] = sn[nb];
          zraw[nb] = zn[nb];
        }
        fetch(c + MG_TEAMS);                          // next chunk of this team: in flight during the math below
      }
      if (grouped) {
        // Marlin's grouped scale permutation puts this lane's 8 scales (columns 16j + 8b + cq of a block) in 16
        // contiguous bytes at position 8*cq of the block's 64-entry row; the 8 zero points are one int32 (4-bit)
        // or two (8-bit); float zero points are laid out like the scales
        for (int nb = 0; nb < nblk; ++nb) {
          uint4 sv;
          uint32_t z0 = 0, z1 = 0;
          uint32_t zf[4] = {0, 0, 0, 0};
          if constexpr (RING) {
            sv = lds128(slot + Cfg::SC_OFF + nb * 512);
            if constexpr (ZP == ZP_INT && BITS == 4) z0 = lds32(slot + Cfg::ZP_OFF + nb * 128 - (uint32_t)lane * 12u);
            if constexpr (ZP == ZP_INT && BITS == 8) {
              z0 = lds32(slot + Cfg::ZP_OFF + nb * 256 - (uint32_t)lane * 8u);
              z1 = lds32(slot + Cfg::ZP_OFF + nb * 256 - (uint32_t)lane * 8u + 4u);
            }
            if constexpr (ZP == ZP_FLOAT) {
              const uint4 zv = lds128(slot + Cfg::ZP_OFF + nb * 512);
              zf[0] = zv.x; zf[1] = zv.y; zf[2] = zv.z; zf[3] = zv.w;
            }
          } else {
            sv = nb == 0 ? sraw[0] : sraw[1];
            const uint4 zv = nb == 0 ? zraw[0] : zraw[1];
            z0 = zv.x; z1 = zv.y;
            zf[0] = zv.x; zf[1] = zv.y; zf[2] = zv.z; zf[3] = zv.w;
          }
          const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const uint32_t s16 = (sw[e >> 1] >> (16 * (e & 1))) & 0xffffu;
            s2[nb][e] = s16 | (s16 << 16);
            if constexpr (ZP == ZP_INT) off2[nb][e] = W::offset_of(zp_code<BITS>(e, z0, z1));
            if constexpr (ZP == ZP_FLOAT) {
              const uint32_t z16 = (zf[e >> 1] >> (16 * (e & 1))) & 0xffffu;
              off2[nb][e] = z16 | (z16 << 16);
            }
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
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row0 = nb * 64 + 16 * j + cq;
            const uint32_t r0 = wt + row0 * 128;          // column 16j + cq      (b = 0)
            const uint32_t r1 = r0 + 8 * 128;             // column 16j + 8 + cq  (b = 1)
            if constexpr (BITS == 4) {
              const uint32_t wq[4] = {q[nb][0].x, q[nb][0].y, q[nb][0].z, q[nb][0].w};
              const uint32_t w = wq[j];
              const uint32_t x00 = lop3_and_or(w, 0x000f000fu, DQ<T>::MAGIC);
              const uint32_t x01 = lop3_and_or(w >> 4, 0x000f000fu, DQ<T>::MAGIC);
              const uint32_t x10 = lop3_and_or(w >> 8, 0x000f000fu, DQ<T>::MAGIC);
              const uint32_t x11 = lop3_and_or(w >> 12, 0x000f000fu, DQ<T>::MAGIC);
              sts32(r0 + a0, W::finish(x00, off2[nb][2 * j], s2[nb][2 * j]));
              sts32(r0 + a1, W::finish(x01, off2[nb][2 * j], s2[nb][2 * j]));
              sts32(r1 + a0, W::finish(x10, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
              sts32(r1 + a1, W::finish(x11, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
            } else {
              // lane's words of the block: index 2j + b; plane 0 holds j = 0,1, plane 1 holds j = 2,3
              const uint32_t wq[8] = {q[nb][0].x, q[nb][0].y, q[nb][0].z, q[nb][0].w,
                                      q[nb][BITS / 4 - 1].x, q[nb][BITS / 4 - 1].y, q[nb][BITS / 4 - 1].z, q[nb][BITS / 4 - 1].w};
              const uint32_t wb0 = wq[2 * j], wb1 = wq[2 * j + 1];
              sts32(r0 + a0, W::pair8(wb0, 0, off2[nb][2 * j], s2[nb][2 * j]));
              sts32(r0 + a1, W::pair8(wb0, 1, off2[nb][2 * j], s2[nb][2 * j]));
              sts32(r1 + a0, W::pair8(wb1, 0, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
              sts32(r1 + a1, W::pair8(wb1, 1, off2[nb][2 * j + 1], s2[nb][2 * j + 1]));
            }
          }
        }
      }
      if (!(p.debug & 8)) fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_w[s]);
      if constexpr (RING) prefetch(c + MG_RING_DEPTH * MG_TEAMS);   // refill the slot just consumed (its words are in registers)
      s += MG_TEAMS;
      while (s >= NS) { s -= NS; ++use; }
    }
    if constexpr (RING) cp_async_wait<0>();
    if (prof && warp == 0 && lane == 0) { g_mg_prof[12] = (unsigned long long)(clock64() - t_role0); g_mg_prof[13] = w0; g_mg_prof[14] = w1; }

    // ===================== epilogue: TMEM -> C =====================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quad = warp & 3;                            // TMEM lanes 32*quad .. +31 belong to this warp
    const int ch = n_base + quad * 32 + lane;
    const bool ch_ok = ch < p.N;
    T* cptr = reinterpret_cast<T*>(p.c);
    float* slab = (!MOE && p.split_k > 1) ? p.c_tmp + (size_t)blockIdx.z * p.M * p.N : nullptr;
    // the MG_TEAMS warps that share a lane quadrant interleave 32-column slabs
    for (int col0 = team * 32; col0 < n_mma; col0 += 32 * MG_TEAMS) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, v);
      if (ch_ok) {
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          if (col0 + t < toks) {
            const float f = __uint_as_float(v[t]);
            if constexpr (MOE) {
              // sorted position -> row of C; padding positions hold ids >= M * topk. The optional routing weight
              // multiplies the ROUNDED product in fp32, as the reference (marlin_moe_ops.cu:945-956)
              const int row = __ldg(p.sorted_ids + tok_base + col0 + t);
              if (row < p.valid_rows) {
                T o = from_f32<T>(f);
                if (p.topk_weights != nullptr) o = from_f32<T>(__ldg(p.topk_weights + row) * to_f32<T>(o));
                cptr[(size_t)row * p.N + ch] = o;
              }
            } else {
              const int tok = tok_base + col0 + t;
              if (slab != nullptr) slab[(size_t)tok * p.N + ch] = f;
              else cptr[(size_t)tok * p.N + ch] = from_f32<T>(f);
            }
          }
        }
      }
    }
    if (prof && warp == 0 && lane == 0) g_mg_prof[15] = (unsigned long long)(clock64() - t_role0);
    tc_fence_before();
  }
  if (!MOE && p.split_k > 1) {
    // split-k: the splits of a tile are the CTAs of one thread-block CLUSTER (1, 1, split_k), so the hardware
    // guarantees that they are co-resident and `barrier.cluster` is a rendezvous that cannot deadlock, whatever else
    // runs on the GPU (round 1 spun on a lock in global memory and relied on the whole grid being resident). Every
    // split has stored its fp32 partial in its own slab; after the barrier EACH split reduces an interleaved
    // 1 / split_k share of the token rows over the slabs in split order: deterministic, no atomics, no second kernel,
    // and the reduction is spread over split_k SMs. The lock workspace of the reference's contract is not touched.
    __threadfence();                                            // slab stores -> visible to the cluster's other SMs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp < MG_DQ_WARPS && nchunks > 0) {
      T* cptr = reinterpret_cast<T*>(p.c);
      const int nvalid = min(MG_NT, p.N - n_base);              // channels of this tile (64 or 128)
      if (lane * 4 < nvalid) {
        for (int r = blockIdx.z + p.split_k * warp; r < toks; r += p.split_k * MG_DQ_WARPS) {
          const size_t off = (size_t)(tok_base + r) * p.N + n_base + lane * 4;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int z0 = 0; z0 < p.split_k; z0 += 4) {           // four slab loads in flight, added in split order
            float4 v[4];
#pragma unroll
            for (int z = 0; z < 4; ++z)
              v[z] = (z0 + z < p.split_k)
                         ? __ldcg(reinterpret_cast<const float4*>(p.c_tmp + (size_t)(z0 + z) * p.M * p.N + off))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int z = 0; z < 4; ++z) { acc.x += v[z].x; acc.y += v[z].y; acc.z += v[z].z; acc.w += v[z].w; }
          }
          uint2 o;
          o.x = pack2<T>(acc.x, acc.y);
          o.y = pack2<T>(acc.z, acc.w);
          *reinterpret_cast<uint2*>(cptr + off) = o;
        }
      }
    }
  }
  __syncthreads();
  if (warp == MG_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// ---- grouped launch helpers (marlin_gemm_moe) -------------------------------------------------------------------
// expert_offsets[e] = first sorted position of expert e when every expert's rows are padded to block_size — the
// layout moe_align_block_size produced sorted_ids with (the reference recomputes it the same way,
// marlin_moe_ops.cu:237-257)
__global__ void moe_expert_offsets_kernel(const int* __restrict__ topk_ids, int* __restrict__ expert_offsets,
                                          int numel, int num_experts, int block_size) {
  extern __shared__ int moe_cnt[];
  for (int e = threadIdx.x; e < num_experts; e += blockDim.x) moe_cnt[e] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < numel; i += blockDim.x) {
    const int e = topk_ids[i];
    if (e >= 0 && e < num_experts) atomicAdd(&moe_cnt[e], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    expert_offsets[0] = 0;
    for (int e = 0; e < num_experts; ++e) {
      tot += (moe_cnt[e] + block_size - 1) / block_size * block_size;
      expert_offsets[e + 1] = tot;
    }
  }
}

// a_sorted[i, :] = A[row(i), perm_e[:]] for every sorted position i of a real (non-padding) row; one CTA per
// position, 16-byte vectors without a permutation. Padding positions are left untouched: their MMA columns are
// independent of the real ones and never stored.
template <typename T>
__global__ void moe_gather_rows_kernel(const T* __restrict__ a, const int* __restrict__ sorted_ids,
                                       const int* __restrict__ expert_offsets, const int* __restrict__ perm,
                                       T* __restrict__ a_sorted, int K, int num_experts, int valid_rows, int topk,
                                       int replicate) {
  const int i = blockIdx.x;
  if (i >= __ldg(expert_offsets + num_experts)) return;
  const int row = __ldg(sorted_ids + i);
  if (row >= valid_rows) return;
  const T* src = a + (size_t)(replicate ? row / topk : row) * K;
  T* dst = a_sorted + (size_t)i * K;
  if (perm == nullptr) {
    for (int k = threadIdx.x * 8; k < K; k += blockDim.x * 8)
      *reinterpret_cast<uint4*>(dst + k) = __ldg(reinterpret_cast<const uint4*>(src + k));
  } else {
    int e = 0;
    while (e + 1 < num_experts && i >= __ldg(expert_offsets + e + 1)) ++e;
    const int* pe = perm + (size_t)e * K;
    for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] = src[__ldg(pe + k)];
  }
}

// ---- host -----------------------------------------------------------------------------------------
static int plan_split_k(int M, int N, int K, int group_size) {
  const int tiles = ((N + MG_NT - 1) / MG_NT) * ((M + MG_TOK - 1) / MG_TOK);
  const int chunks = K / MG_KC;
  const int sms = num_sms();
  if ((M + MG_TOK - 1) / MG_TOK > 32) return 1;   // lock workspace (N/64*16 ints) covers <= 32 token blocks
  // as many k-splits as fit one wave of the SMs (one CTA per SM) with at least 8 chunks each: 48 tiles -> 3 splits =
  // 144 CTAs (a power-of-two rule left a third of the SMs idle)
  int split = std::min(std::min(sms / std::max(tiles, 1), chunks / 8), 8);   // 8 = portable cluster size
  if (split < 1) split = 1;
  // keep every split on a group boundary and none empty
  const int gchunks = group_size > MG_KC ? group_size / MG_KC : 1;
  for (; split > 1; --split) {
    int per = (chunks + split - 1) / split;
    per = (per + gchunks - 1) / gchunks * gchunks;
    if ((split - 1) * per < chunks) break;
  }
  return split;
}

template <typename T, int ZP, int BITS, bool MOE, bool RING>
static int launch_marlin_v(const CUtensorMap& tmap, MarlinParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = MGCfg<BITS, ZP>;
  constexpr int FIXED = RING ? Cfg::FIXED : 2048;
  auto kern = marlin_w4a16_tc5_kernel<T, ZP, BITS, MOE, RING>;
  static thread_local uint64_t attr_done = 0;
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM_TOTAL));
    attr_done |= 1ull << (dev & 63);
  }
  // shared-memory plan: as many MMA stages (activation tile + dequantised weight tile) as fit, 2..8
  p.stages = (MG_SMEM_TOTAL - FIXED) / (p.act_bytes + MG_W_BYTES);
  if (p.stages > MG_MAX_STAGES) p.stages = MG_MAX_STAGES;
  B200_CHECK(p.stages >= 2, "marlin gemm: shared-memory plan leaves fewer than two pipeline stages");
  const size_t smem = (size_t)p.stages * (p.act_bytes + MG_W_BYTES) + FIXED;
  if (!MOE && p.split_k > 1) {
    // the k-splits of a tile form one cluster (1, 1, split_k): co-residency guaranteed by the hardware
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(MG_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)p.split_k;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmap, p));
  } else {
    kern<<<grid, MG_THREADS, smem, st>>>(tmap, p);
  }
  return check_launch("marlin_w4a16_tc5_kernel");
}

// how many clusters of `split` CTAs of this kernel the device can hold at once (the k-splits of a tile must not be
// queued behind other tiles' clusters, or the split costs a second wave); all instantiations share the launch shape
static int max_active_clusters(int split) {
  static thread_local int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (split < 2 || split > 8) return 1 << 30;
  if (cache[split] == 0) {
    auto kern = marlin_w4a16_tc5_kernel<__nv_bfloat16, ZP_NONE, 4, false, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM_TOTAL);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, 1, (unsigned)split);
    cfg.blockDim = dim3(MG_THREADS, 1, 1);
    cfg.dynamicSmemBytes = MG_SMEM_TOTAL - 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)split;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms() / split;      // no device (plan queries on a CPU box) or no answer: the arithmetic bound
    }
    cache[split] = n;
  }
  return cache[split];
}

// weight-operand path: register prefetch by default (one more pipeline stage at 256 tokens: 794 vs 762 TFLOP/s
// on 4096 x 28672, 719 vs 639 on 14336 x 4096); B200_MARLIN_RING=1 selects the per-warp cp.async ring
static bool marlin_use_ring() {
#ifdef B200_MARLIN_DEBUG_BUILD
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_MARLIN_RING");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
#else
  return false;
#endif
}

template <typename T, int ZP, int BITS, bool MOE>
static int launch_marlin(const CUtensorMap& tmap, MarlinParams& p, dim3 grid, cudaStream_t st) {
  return marlin_use_ring() ? launch_marlin_v<T, ZP, BITS, MOE, true>(tmap, p, grid, st)
                           : launch_marlin_v<T, ZP, BITS, MOE, false>(tmap, p, grid, st);
}

template <bool MOE>
static int dispatch_marlin(const CUtensorMap& tmap, MarlinParams& p, dim3 grid, int dtype, int zp, int bits,
                           cudaStream_t st) {
#define B200_MG(TT, ZZ, BB) return launch_marlin<TT, ZZ, BB, MOE>(tmap, p, grid, st)
  if constexpr (MOE) {
    if (dtype == B200_BF16) B200_MG(__nv_bfloat16, ZP_NONE, 4);
    B200_MG(__half, ZP_NONE, 4);
  } else {
    if (dtype == B200_BF16) {
      if (bits == 4) { if (zp == ZP_INT) B200_MG(__nv_bfloat16, ZP_INT, 4); B200_MG(__nv_bfloat16, ZP_NONE, 4); }
      if (zp == ZP_INT) B200_MG(__nv_bfloat16, ZP_INT, 8);
      B200_MG(__nv_bfloat16, ZP_NONE, 8);
    }
    if (bits == 4) {
      if (zp == ZP_FLOAT) B200_MG(__half, ZP_FLOAT, 4);
      if (zp == ZP_INT) B200_MG(__half, ZP_INT, 4);
      B200_MG(__half, ZP_NONE, 4);
    }
    if (zp == ZP_FLOAT) B200_MG(__half, ZP_FLOAT, 8);
    if (zp == ZP_INT) B200_MG(__half, ZP_INT, 8);
    B200_MG(__half, ZP_NONE, 8);
  }
#undef B200_MG
}

// 2-D tensor map over a row-major [rows, K] 16-bit matrix, box [64 k x box_rows], SWIZZLE_128B
static int encode_act_map(CUtensorMap* tmap, const void* a, int64_t rows, int size_k, int box_rows, int dtype) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)size_k, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)size_k * 2};
  const cuuint32_t box[2] = {(cuuint32_t)MG_KC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tmap, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                         2, const_cast<void*>(a), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return 0;
}

// group bookkeeping shared by the dense and the grouped entry points
static void fill_group_params(MarlinParams& p, int size_k, int gs) {
  p.group_size = gs;
  p.grouped = (gs > 0 && gs < size_k) ? 1 : 0;
  p.rows_per_chunk = (gs > 0 && gs < MG_KC) ? MG_KC / gs : 1;
  p.chunks_per_group = (gs > MG_KC) ? gs / MG_KC : 1;
  p.cpg_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == p.chunks_per_group) p.cpg_shift = sh;
#ifdef B200_MARLIN_DEBUG_BUILD     // timing experiments only (bits that skip work make results wrong by construction):
  const char* dbg = getenv("B200_MARLIN_DEBUG");   // compiled out of the shipped library
  p.debug = dbg ? atoi(dbg) : 0;
#else
  p.debug = 0;
#endif
}

// small-batch mma.sync kernel (marlin_gemm_small.cu)
int marlin_small_plan(int M, int N, int K, int group_size);
int marlin_small_gemm(const void* a, const void* b_q, const void* scales, const void* zeros, void* c, float* c_tmp,
                      int* locks, int M, int N, int K, int num_groups, int has_zp, int dtype, int split_k,
                      cudaStream_t st);
static constexpr int MG_SMALL_M = 32;
static bool marlin_use_small() {
#ifdef B200_MARLIN_DEBUG_BUILD
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_MARLIN_SMALL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
#else
  return true;
#endif
}

}  // namespace b200

using namespace b200;

extern "C" int b200_debug_marlin_prof(unsigned long long* out32) {
  B200_CUDA_OK(cudaDeviceSynchronize());
  B200_CUDA_OK(cudaMemcpyFromSymbol(out32, g_mg_prof, sizeof(unsigned long long) * 32));
  return 0;
}

// upper bound of the number of fp32 partial slabs [M, N] the GEMM may use for this shape (sizes c_tmp)
extern "C" int b200_marlin_gemm_plan(int size_m, int size_n, int size_k, int num_groups) {
  const int gs = num_groups > 1 ? size_k / num_groups : -1;
  if (size_m <= 0 || size_n <= 0 || size_k < MG_KC) return 1;
  int split = plan_split_k(size_m, size_n, size_k, gs);
  if (size_m <= MG_SMALL_M) split = std::max(split, marlin_small_plan(size_m, size_n, size_k, gs));
  return split;
}

extern "C" int b200_gptq_marlin_gemm(const void* a, const void* b_q_weight, const void* b_scales,
                                     const void* b_zeros, void* c, float* c_tmp, int32_t* workspace,
                                     int size_m, int size_n, int size_k, int num_groups, int num_bits,
                                     int has_zp, int dtype, int split_k, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "gpt_marlin_gemm only supports bfloat16 and float16");
  B200_CHECK(num_bits == 4 || num_bits == 8, "num_bits must be 4 or 8. Got = " + std::to_string(num_bits));
  B200_CHECK(has_zp >= ZP_NONE && has_zp <= ZP_FLOAT, "has_zp must be 0 (none), 1 (packed integers) or 2 (floats)");
  B200_CHECK(has_zp != ZP_FLOAT || dtype == B200_F16,
             "Computation type must be float16 (half) when using float zero points.");
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
  if (num_bits == 4 && has_zp != ZP_FLOAT && size_m <= MG_SMALL_M && marlin_use_small())
    return marlin_small_gemm(a, b_q_weight, b_scales, b_zeros, c, c_tmp, workspace, size_m, size_n, size_k, num_groups,
                             has_zp, dtype, split_k, (cudaStream_t)stream);
  if (split_k <= 0) split_k = plan_split_k(size_m, size_n, size_k, gs);
  B200_CHECK(split_k == 1 || (c_tmp != nullptr && workspace != nullptr),
             "split-k needs the fp32 partial buffer [split_k, M, N] and the zeroed lock workspace");
  {
    const int tiles = ((size_n + MG_NT - 1) / MG_NT) * ((size_m + MG_TOK - 1) / MG_TOK);
    while (split_k > 1 && (tiles * split_k > num_sms() || tiles > max_active_clusters(split_k))) --split_k;   // one wave
    if (split_k > 8) split_k = 8;
    const int chunks_total = size_k / MG_KC;
    while (split_k > 1 && (split_k - 1) * ((chunks_total + split_k - 1) / split_k) >= chunks_total) --split_k;
  }
  const int toks = size_m < MG_TOK ? size_m : MG_TOK;
  const int box_rows = (toks + 15) & ~15;
  CUtensorMap tmap;
  if (int rc = encode_act_map(&tmap, a, size_m, size_k, box_rows, dtype)) return rc;
  MarlinParams p{};
  p.b_q = (const uint32_t*)b_q_weight; p.scales = b_scales; p.zeros = b_zeros;
  p.c = c; p.c_tmp = c_tmp; p.locks = workspace; p.M = size_m; p.N = size_n; p.K = size_k;
  p.split_k = split_k;
  p.box_rows = box_rows;
  p.act_bytes = (box_rows * 128 + 1023) & ~1023;
  fill_group_params(p, size_k, gs);
  const int chunks = size_k / MG_KC;
  {
    const int gchunks = gs > MG_KC ? gs / MG_KC : 1;
    int per = (chunks + split_k - 1) / split_k;
    per = (per + gchunks - 1) / gchunks * gchunks;            // splits start on group boundaries
    while (split_k > 1 && (split_k - 1) * per >= chunks) {    // never an empty split
      --split_k;
      per = ((chunks + split_k - 1) / split_k + gchunks - 1) / gchunks * gchunks;
    }
    p.split_k = split_k;
    p.chunks_per_split = per;
  }
  dim3 grid((size_n + MG_NT - 1) / MG_NT, (size_m + MG_TOK - 1) / MG_TOK, split_k);
  return dispatch_marlin<false>(tmap, p, grid, dtype, has_zp, num_bits, (cudaStream_t)stream);
}

extern "C" int b200_marlin_gemm_moe(const void* a, const void* b_q_weights, const int32_t* sorted_ids,
                                    int64_t sorted_capacity, const float* topk_weights, const int32_t* topk_ids,
                                    const void* b_scales, const int32_t* perm, void* c, void* a_sorted,
                                    int32_t* expert_offsets, int size_m, int size_n, int size_k, int num_groups,
                                    int num_experts, int topk, int moe_block_size, int replicate_input,
                                    int apply_weights, int dtype, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "marlin_gemm_moe only supports bfloat16 and float16");
  B200_CHECK(size_m > 0 && size_n > 0 && size_k > 0, "Invalid MNK = [" + std::to_string(size_m) + ", " +
             std::to_string(size_n) + ", " + std::to_string(size_k) + "]");
  B200_CHECK(size_n % 64 == 0, "prob_n = " + std::to_string(size_n) + " is not divisible by thread_n = 64");
  B200_CHECK(size_k % MG_KC == 0, "prob_k = " + std::to_string(size_k) + " is not divisible by thread_k = 64");
  B200_CHECK(num_experts >= 1 && num_experts <= 1024 && topk >= 1 && moe_block_size >= 1, "invalid expert / topk / block size");
  B200_CHECK(num_groups >= 1 && size_k % num_groups == 0, "size_k is not divisible by the number of scale groups");
  const int gs = num_groups > 1 ? size_k / num_groups : -1;
  B200_CHECK(gs == -1 || gs == size_k || gs == 32 || (gs % 64 == 0), "unsupported group size " + std::to_string(gs));
  B200_CHECK(a_sorted != nullptr && expert_offsets != nullptr && sorted_capacity > 0,
             "marlin_gemm_moe needs the gathered-activation scratch [sorted_capacity, K] and expert_offsets [E + 1]");
  B200_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a_sorted) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(b_q_weights) & 15) == 0, "a, a_sorted and b_q_weights must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int valid_rows = size_m * topk;
  moe_expert_offsets_kernel<<<1, 256, num_experts * sizeof(int), st>>>(topk_ids, expert_offsets, valid_rows,
                                                                      num_experts, moe_block_size);
  if (int rc = check_launch("moe_expert_offsets_kernel")) return rc;
  const int gthreads = perm ? 256 : std::max(32, std::min(256, size_k / 8));
  if (dtype == B200_BF16)
    moe_gather_rows_kernel<__nv_bfloat16><<<(unsigned)sorted_capacity, gthreads, 0, st>>>(
        (const __nv_bfloat16*)a, sorted_ids, expert_offsets, perm, (__nv_bfloat16*)a_sorted, size_k, num_experts,
        valid_rows, topk, replicate_input);
  else
    moe_gather_rows_kernel<__half><<<(unsigned)sorted_capacity, gthreads, 0, st>>>(
        (const __half*)a, sorted_ids, expert_offsets, perm, (__half*)a_sorted, size_k, num_experts, valid_rows, topk,
        replicate_input);
  if (int rc = check_launch("moe_gather_rows_kernel")) return rc;

  // rows of one expert <= tokens rounded up to the block size (a token routes to an expert at most once); longer
  // segments simply take more tiles
  const int blk = moe_block_size > 16 ? moe_block_size : 16;
  int tile_rows = ((size_m + blk - 1) / blk * blk + 15) & ~15;
  if (tile_rows > MG_TOK) tile_rows = MG_TOK;
  CUtensorMap tmap;
  if (int rc = encode_act_map(&tmap, a_sorted, sorted_capacity, size_k, tile_rows, dtype)) return rc;
  MarlinParams p{};
  p.b_q = (const uint32_t*)b_q_weights; p.scales = b_scales; p.zeros = nullptr;
  p.c = c; p.M = valid_rows; p.N = size_n; p.K = size_k;
  p.split_k = 1;
  p.box_rows = tile_rows;
  p.act_bytes = (tile_rows * 128 + 1023) & ~1023;
  fill_group_params(p, size_k, gs);
  p.chunks_per_split = size_k / MG_KC;
  p.expert_offsets = expert_offsets; p.sorted_ids = sorted_ids;
  p.topk_weights = apply_weights ? topk_weights : nullptr;
  p.num_experts = num_experts; p.valid_rows = valid_rows; p.tile_rows = tile_rows;
  p.expert_words = (long long)(size_k / 16) * size_n * 2;
  p.expert_scales = (long long)num_groups * size_n;
  const int max_tiles = num_experts + (int)(sorted_capacity / tile_rows);
  dim3 grid((size_n + MG_NT - 1) / MG_NT, max_tiles, 1);
  return dispatch_marlin<true>(tmap, p, grid, dtype, ZP_NONE, 4, st);
}
