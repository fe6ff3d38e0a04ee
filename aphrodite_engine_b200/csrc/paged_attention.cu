// paged_attention.cu — single-query (decode) attention over a block-table-indexed KV cache, sm_100a.
//
// Replaces the reference's paged_attention_v1 / paged_attention_v2 (+ v2 reduce)
// (kernels/attention/attention_kernels.cu:87-496 kernel, :565-669 reduce, :690-998 launchers).
// Not a port: the reference runs one 128-thread CTA per (q-head, seq) with scalar FMA dot products and
// re-reads K/V once per q-head of a GQA group. Here:
//
//   * one CTA per (kv-head, seq[, 512-token partition]) serves ALL q-heads of the GQA group, so K/V
//     cross HBM exactly once;
//   * a producer warp walks the block table and streams each (block, kv-head) chunk — a contiguous
//     head_size*block_size*sizeof(elt) run for K and for V — into a shared-memory ring with
//     cp.async.bulk (TMA engine, SASS UBLKCP) completing on mbarriers, L2 evict-first;
//   * consumer warps run both contractions on tensor cores in the TRANSPOSED form
//       S^T[tok,head] = K[tok,:] . Q^T      O^T[d,head] += V^T[d,tok] . P^T[tok,head]
//     so the paged layouts are consumed as they lie: the K chunk [D/8][tok][8] is a grid of 8x16-B
//     ldmatrix tiles (A operand, k = head dim), the V chunk [D][tok] is row-major A (k = token);
//     P is transposed in registers with movmatrix; online softmax in fp32 with exp2;
//   * fp8 KV (e4m3/e5m2) is dequantised between the shared-memory tile and the MMA operand
//     registers, following quant_utils.cuh's fp8 -> half -> float*scale -> T rounding chain.
//
// Shapes the tensor-core kernel does not cover (fp32, head 120, block 8, block-sparse, GQA group > 8
// with ALiBi ...) run on a generic SIMT kernel in this file — still on the GPU; there is no CPU path.
#include "common.cuh"

#include <algorithm>
#include <float.h>
#include <math.h>

#include <type_traits>

namespace b200 {

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;
static constexpr int kPartitionSize = 512;  // aphrodite/attention/ops/paged_attn.py:13

struct AttnParams {
  void* out;          // v1: [S,H,D]; v2: tmp_out [S,H,P,D]
  float* exp_sums;    // v2 only [S,H,P]
  float* max_logits;  // v2 only [S,H,P]
  const void* q;
  const void* k_cache;
  const void* v_cache;
  const int32_t* block_tables;
  const int32_t* seq_lens;
  const float* alibi_slopes;
  int num_seqs, num_heads, num_kv_heads, head_size, block_size;
  int max_num_blocks_per_seq;
  int max_num_partitions;  // 0 for v1
  int max_seq_len;         // caller's bound on seq_lens (v1: sizes the cluster split; <= 0: unknown)
  int64_t q_stride, kv_block_stride, kv_head_stride;
  float scale, k_scale, v_scale;
  int tp_rank, bs_local_blocks, bs_vert_stride, bs_block_size, bs_head_sliding_step;
};

// =================================================================================================
// Tensor-core kernel
// =================================================================================================
template <int KV> struct CacheElt { using type = uint8_t; };
template <> struct CacheElt<B200_KV_AUTO> { using type = uint16_t; };

static constexpr int kConsumerWarps = 4;
static constexpr int kFastThreads = (kConsumerWarps + 1) * 32;
static constexpr int kHeadsPerCta = 8;  // MMA N dimension

template <int D, int BS, int KV>
struct FastCfg {
  static constexpr int ESZ = (KV == B200_KV_AUTO) ? 2 : 1;
  static constexpr int CHUNK = D * BS * ESZ;  // bytes of one (block, kv-head) K or V chunk
  static constexpr int STAGE = 2 * CHUNK;
  static constexpr int RING_BUDGET = 100 * 1024;
  static constexpr int NSTAGES_RAW = RING_BUDGET / STAGE;
  // The stage count MUST be a multiple of the consumer-warp count: stage s is then always consumed
  // by warp s % kConsumerWarps, so the parity waits on one mbarrier are issued in program order by a
  // single warp. (With e.g. 6 stages / 4 warps a fast warp could test parity 1 of a barrier whose
  // phase 0 is still in flight — that test passes vacuously and the warp would read a stale stage.)
  static constexpr int NSTAGES_CAP = NSTAGES_RAW > 16 ? 16 : NSTAGES_RAW;
  static constexpr int NSTAGES =
      NSTAGES_CAP < kConsumerWarps ? kConsumerWarps : (NSTAGES_CAP / kConsumerWarps) * kConsumerWarps;
  static constexpr int OPAD = D + 4;  // merge buffer row pitch (floats)
  static constexpr int MERGE_BYTES = kConsumerWarps * kHeadsPerCta * OPAD * 4;
  static constexpr int RING_BYTES = NSTAGES * STAGE;
  static constexpr int DATA_BYTES = RING_BYTES > MERGE_BYTES ? RING_BYTES : MERGE_BYTES;
  // barriers + per-warp (m,l) for the merge
  static constexpr int AUX_BYTES = 2 * NSTAGES * 8 + 2 * kConsumerWarps * kHeadsPerCta * 4;
  static constexpr int SMEM_BYTES = DATA_BYTES + AUX_BYTES + 128;
};

__device__ __forceinline__ float fast_exp2(float x) {  // ex2.approx: exp2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// fp8 KV path. Every e4m3 / e5m2 value is exactly representable in fp16 and the conversion is ONE hardware
// instruction per pair (cvt.rn.f16x2.e4m3x2), so the fp8 contractions run as fp16 MMAs whatever the model dtype:
// Q is converted to fp16 once per CTA (exact for bf16 values in fp16's normal range), P is packed as fp16, and the
// per-tensor scales leave the per-element path: k_scale multiplies the logits, v_scale the final output.
// (The reference materialises T(fp8 * scale) per element, quant_utils.cuh:296-360: identical when the scales are
// 1.0 — its CPU path supports nothing else — and within the rounding noise of that per-element product otherwise;
// the first version of this kernel did exactly that and was ALU-bound at 0.74 of the HBM peak.)
template <int KV>
__device__ __forceinline__ void fp8x4_to_f16(uint32_t w, uint32_t& lo, uint32_t& hi) {
  // inline PTX on purpose: through __nv_cvt_fp8x2_to_halfraw2 the packed result is split into two 16-bit struct
  // fields and re-packed, which cost two PRMT per conversion (30 % of the kernel's instructions, ALU pipe at 67 %)
  if constexpr (KV == B200_KV_FP8_E5M2) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f16x2.e5m2x2 %0, l;\n\tcvt.rn.f16x2.e5m2x2 %1, h;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(w));
  } else {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f16x2.e4m3x2 %0, l;\n\tcvt.rn.f16x2.e4m3x2 %1, h;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(w));
  }
}
// two 16-bit values of T -> packed fp16 pair
template <typename T> __device__ __forceinline__ uint32_t pair_to_f16(uint32_t v);
template <> __device__ __forceinline__ uint32_t pair_to_f16<__half>(uint32_t v) { return v; }
template <> __device__ __forceinline__ uint32_t pair_to_f16<__nv_bfloat16>(uint32_t v) {
  const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
  return pack2<__half>(lo, hi);
}

// ---- thread-block cluster helpers: a v1 launch may split every sequence over the CTAs of a cluster (1, 1, cs) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(const float* local_ptr, uint32_t cta_rank, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(cta_rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}
static constexpr int kMaxClusterSplit = 4;

// Non-partitioned launches may run as clusters of cs CTAs along z: CTA r of a cluster takes the r-th share of the
// sequence's blocks, the partial (max, sum, O) of every CTA is written into the LEADER's shared memory through DSMEM
// and merged there in fp32 — the split-KV idea of paged_attention_v2 without its global scratch, second kernel or
// 16-bit rounding of the partial outputs. It exists for wave quantisation: (kv-heads x seqs) CTAs at 2 per SM run in
// ceil(n / 296) waves, and e.g. 1024 CTAs (BASELINE configs[3] per GPU: 1 kv-head, 1024 seqs) are 3.46 waves = 86 %
// of the HBM roofline at best; as 2048 half-length CTAs they are 6.9 waves = 99 %.
template <typename T, int D, int BS, int KV, bool PARTITIONED>
__global__ void __launch_bounds__(kFastThreads, 2)
paged_attention_tc_kernel(const AttnParams p) {
  using Cfg = FastCfg<D, BS, KV>;
  using MT = typename std::conditional<KV == B200_KV_AUTO, T, __half>::type;   // tensor-core operand type
  constexpr int NST = Cfg::NSTAGES;
  constexpr int KSTEPS = D / 16;
  constexpr int MTILES = D / 16;
  constexpr int SUBTILES = BS / 16;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* ring = smem_raw;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + Cfg::DATA_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  float* merge_m = reinterpret_cast<float*>(empty_bar + NST);  // [warps][8]
  float* merge_l = merge_m + kConsumerWarps * kHeadsPerCta;     // [warps][8]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.y;
  const int part_idx = PARTITIONED ? blockIdx.z : 0;
  const int G = p.num_heads / p.num_kv_heads;        // q-heads per kv-head
  const int groups = (G + kHeadsPerCta - 1) / kHeadsPerCta;
  const int kvh = blockIdx.x / groups;
  const int hgrp = blockIdx.x % groups;
  const int head0 = kvh * G + hgrp * kHeadsPerCta;   // first q-head of this CTA
  const int nheads = min(kHeadsPerCta, G - hgrp * kHeadsPerCta);

  const int seq_len = p.seq_lens[seq];
  if (PARTITIONED && part_idx * kPartitionSize >= seq_len) return;  // nothing to do (reference :117-120)

  const int num_seq_blocks = (seq_len + BS - 1) / BS;
  const uint32_t cs = PARTITIONED ? 1u : cluster_nctarank();     // CTAs sharing this sequence (1: plain launch)
  const uint32_t crank = PARTITIONED ? 0u : cluster_ctarank();
  int start_block = PARTITIONED ? part_idx * (kPartitionSize / BS) : 0;
  int end_block = PARTITIONED ? min(start_block + kPartitionSize / BS, num_seq_blocks) : num_seq_blocks;
  if (!PARTITIONED && cs > 1) {
    const int per = (num_seq_blocks + (int)cs - 1) / (int)cs;
    start_block = (int)crank * per;
    end_block = min(start_block + per, num_seq_blocks);
  }
  const int nblocks = max(end_block - start_block, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ===================== producer warp: block table -> bulk copies =====================
    const int32_t* bt = p.block_tables + (int64_t)seq * p.max_num_blocks_per_seq + start_block;
    const uint8_t* kc = reinterpret_cast<const uint8_t*>(p.k_cache);
    const uint8_t* vc = reinterpret_cast<const uint8_t*>(p.v_cache);
    const int64_t blk_stride_b = p.kv_block_stride * Cfg::ESZ;
    const int64_t head_off_b = (int64_t)kvh * p.kv_head_stride * Cfg::ESZ;
    const uint64_t pol = policy_evict_first();
    int32_t mine = (lane < nblocks) ? bt[lane] : 0;
    for (int base = 0; base < nblocks; base += 32) {
      const int nxt = base + 32 + lane;
      const int32_t next_mine = (nxt < nblocks) ? bt[nxt] : 0;
      const int cnt = min(32, nblocks - base);
      for (int j = 0; j < cnt; ++j) {
        const int32_t blk = __shfl_sync(0xffffffffu, mine, j);
        if (lane == 0) {
          const int i = base + j;
          const int s = i % NST;
          const uint32_t use = (uint32_t)(i / NST);
          mbar_wait(&empty_bar[s], (use & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE);
          const int64_t off = (int64_t)blk * blk_stride_b + head_off_b;
          uint8_t* dst = ring + (size_t)s * Cfg::STAGE;
          bulk_g2s(dst, kc + off, Cfg::CHUNK, &full_bar[s], pol);
          bulk_g2s(dst + Cfg::CHUNK, vc + off, Cfg::CHUNK, &full_bar[s], pol);
        }
      }
      mine = next_mine;
    }
  }

  // per-thread online-softmax state (consumer warps): this thread owns heads h0 = 2q, h1 = 2q+1
  const int g = lane >> 2;
  const int q = lane & 3;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float o[MTILES][4];
#pragma unroll
  for (int i = 0; i < MTILES; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  if (warp < kConsumerWarps) {
    // ===================== consumer warps =====================
    // Q^T B-fragments: b0 = Q[head g][16ks + kA, kA+1], b1 = Q[head g][16ks + kB, kB+1]
    //   16-bit KV: kA = 2q, kB = 2q+8 (natural order, matches ldmatrix of K)
    //   fp8 KV   : kA = 4q, kB = 4q+2 (k permuted so one 4-byte smem read feeds a0/a2)
    uint32_t qf[KSTEPS][2];
    {
      const T* qrow = reinterpret_cast<const T*>(p.q) + (int64_t)seq * p.q_stride +
                      (int64_t)(head0 + g) * D;
      const bool hv = g < nheads;
      constexpr int kA = (KV == B200_KV_AUTO) ? 2 : 4;
      constexpr int kBoff = (KV == B200_KV_AUTO) ? 8 : 2;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int d0 = ks * 16 + kA * q;
        qf[ks][0] = hv ? *reinterpret_cast<const uint32_t*>(qrow + d0) : 0u;
        qf[ks][1] = hv ? *reinterpret_cast<const uint32_t*>(qrow + d0 + kBoff) : 0u;
        if constexpr (KV != B200_KV_AUTO) {
          qf[ks][0] = pair_to_f16<T>(qf[ks][0]);
          qf[ks][1] = pair_to_f16<T>(qf[ks][1]);
        }
      }
    }
    const float sc = p.scale * kLog2e * (KV == B200_KV_AUTO ? 1.f : p.k_scale);
    float alibi[2] = {0.f, 0.f};
    if (p.alibi_slopes != nullptr) {
      if (2 * q < nheads) alibi[0] = p.alibi_slopes[head0 + 2 * q] * kLog2e;
      if (2 * q + 1 < nheads) alibi[1] = p.alibi_slopes[head0 + 2 * q + 1] * kLog2e;
    }
    const bool use_alibi = p.alibi_slopes != nullptr;

    // lane-dependent parts of the ldmatrix row addresses
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t k_lane = (uint32_t)((mi >> 1) * (BS * 16) + (mi & 1) * 128 + r8 * 16);
    const uint32_t v_lane = (uint32_t)(((mi & 1) * 8 + r8) * (BS * 2) + (mi >> 1) * 16);
    // fp8: token handled by MMA row g / g+8 (see header comment on the k/token permutation)
    const int tokA = 4 * (g >> 1) + (g & 1);
    const int tokB = tokA + 2;

    const uint32_t ring_s = smem_u32(ring);

    for (int i = warp; i < nblocks; i += kConsumerWarps) {
      const int s = i % NST;
      const uint32_t use = (uint32_t)(i / NST);
      mbar_wait(&full_bar[s], use & 1u);
      const uint32_t k_s = ring_s + (uint32_t)s * Cfg::STAGE;
      const uint32_t v_s = k_s + Cfg::CHUNK;
      const uint8_t* k_g = ring + (size_t)s * Cfg::STAGE;
      const uint8_t* v_g = k_g + Cfg::CHUNK;

#pragma unroll
      for (int sub = 0; sub < SUBTILES; ++sub) {
        const int tok0 = (start_block + i) * BS + sub * 16;  // first token of this 16-token tile
        if (tok0 >= seq_len) break;
        const int nvalid = seq_len - tok0;  // >= 1; tile fully valid when >= 16

        // ---------------- S^T = K . Q^T ----------------
        float st[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (KV == B200_KV_AUTO) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            uint32_t a[4];
            ldmatrix_x4(a[0], a[1], a[2], a[3],
                        k_s + (uint32_t)(2 * ks * BS * 16 + sub * 256) + k_lane);
            mma_16816<T>(st, a, qf[ks][0], qf[ks][1]);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint8_t* kr = k_g + ks * (BS * 16) + 4 * q;
            const uint32_t wA = *reinterpret_cast<const uint32_t*>(kr + (sub * 16 + tokA) * 16);
            const uint32_t wB = *reinterpret_cast<const uint32_t*>(kr + (sub * 16 + tokB) * 16);
            uint32_t a[4];
            fp8x4_to_f16<KV>(wA, a[0], a[2]);
            fp8x4_to_f16<KV>(wB, a[1], a[3]);
            mma_16816<__half>(st, a, qf[ks][0], qf[ks][1]);
          }
        }

        // ---------------- online softmax (log2 domain) ----------------
        // st[0] = (tok rA, head 2q), st[1] = (rA, 2q+1), st[2] = (rB, 2q), st[3] = (rB, 2q+1)
        const int rA = (KV == B200_KV_AUTO) ? g : tokA;
        const int rB = (KV == B200_KV_AUTO) ? g + 8 : tokB;
        float t[4];
        t[0] = st[0] * sc; t[1] = st[1] * sc; t[2] = st[2] * sc; t[3] = st[3] * sc;
        if (use_alibi) {
          const float pa = (float)(tok0 + rA - seq_len + 1), pb = (float)(tok0 + rB - seq_len + 1);
          t[0] += alibi[0] * pa; t[1] += alibi[1] * pa;
          t[2] += alibi[0] * pb; t[3] += alibi[1] * pb;
        }
        if (nvalid < 16) {
          if (rA >= nvalid) { t[0] = -INFINITY; t[1] = -INFINITY; }
          if (rB >= nvalid) { t[2] = -INFINITY; t[3] = -INFINITY; }
        }
        float mx0 = fmaxf(t[0], t[2]), mx1 = fmaxf(t[1], t[3]);
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, off));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, off));
        }
        const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
        if (__any_sync(0xffffffffu, (mn0 != m_run[0]) || (mn1 != m_run[1]))) {
          const float a0 = fast_exp2(m_run[0] - mn0), a1 = fast_exp2(m_run[1] - mn1);
          l_run[0] *= a0; l_run[1] *= a1;
#pragma unroll
          for (int mt = 0; mt < MTILES; ++mt) {
            o[mt][0] *= a0; o[mt][1] *= a1; o[mt][2] *= a0; o[mt][3] *= a1;
          }
          m_run[0] = mn0; m_run[1] = mn1;
        }
        const float p0 = fast_exp2(t[0] - mn0), p1 = fast_exp2(t[1] - mn1);
        const float p2 = fast_exp2(t[2] - mn0), p3 = fast_exp2(t[3] - mn1);
        l_run[0] += p0 + p2;
        l_run[1] += p1 + p3;
        // P^T B-fragments via in-register 8x8 transposes
        const uint32_t pb0 = movmatrix_trans(pack2<MT>(p0, p1));
        const uint32_t pb1 = movmatrix_trans(pack2<MT>(p2, p3));

        // ---------------- O^T += V^T . P^T ----------------
        if constexpr (KV == B200_KV_AUTO) {
          // zero V entries of out-of-context tokens (they may hold NaNs; reference :412-421)
          uint32_t msk01 = 0xffffffffu, msk23 = 0xffffffffu;
          if (nvalid < 16) {
            msk01 = (2 * q < nvalid ? 0x0000ffffu : 0u) | (2 * q + 1 < nvalid ? 0xffff0000u : 0u);
            msk23 = (2 * q + 8 < nvalid ? 0x0000ffffu : 0u) | (2 * q + 9 < nvalid ? 0xffff0000u : 0u);
          }
#pragma unroll
          for (int mt = 0; mt < MTILES; ++mt) {
            uint32_t a[4];
            ldmatrix_x4(a[0], a[1], a[2], a[3],
                        v_s + (uint32_t)(mt * 16 * BS * 2 + sub * 32) + v_lane);
            if (nvalid < 16) { a[0] &= msk01; a[1] &= msk01; a[2] &= msk23; a[3] &= msk23; }
            mma_16816<T>(o[mt], a, pb0, pb1);
          }
        } else {
          // fp8 V [D][BS] bytes: thread reads tokens 4q..4q+3 of rows d = 16mt+g and +8
          uint32_t bmask = 0xffffffffu;
          if (nvalid < 16) {
            bmask = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * q + e < nvalid) bmask |= 0xffu << (8 * e);
          }
#pragma unroll
          for (int mt = 0; mt < MTILES; ++mt) {
            const uint8_t* vr = v_g + (mt * 16 + g) * BS + sub * 16 + 4 * q;
            const uint32_t wA = *reinterpret_cast<const uint32_t*>(vr) & bmask;
            const uint32_t wB = *reinterpret_cast<const uint32_t*>(vr + 8 * BS) & bmask;
            uint32_t a[4];
            fp8x4_to_f16<KV>(wA, a[0], a[2]);
            fp8x4_to_f16<KV>(wB, a[1], a[3]);
            mma_16816<__half>(o[mt], a, pb0, pb1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    // full row sums: add the partial sums of the 8 lanes that share q
#pragma unroll
    for (int off = 4; off <= 16; off <<= 1) {
      l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], off);
      l_run[1] += __shfl_xor_sync(0xffffffffu, l_run[1], off);
    }
  }

  // ===================== merge the consumer warps' partial results =====================
  __syncthreads();  // every stage has been consumed: the ring can be reused as the merge buffer
  float* mo = reinterpret_cast<float*>(ring);  // [warps][8 heads][OPAD]
  if (warp < kConsumerWarps) {
    float* w = mo + warp * kHeadsPerCta * Cfg::OPAD;
#pragma unroll
    for (int mt = 0; mt < MTILES; ++mt) {
      const int d = mt * 16 + g;
      w[(2 * q) * Cfg::OPAD + d] = o[mt][0];
      w[(2 * q + 1) * Cfg::OPAD + d] = o[mt][1];
      w[(2 * q) * Cfg::OPAD + d + 8] = o[mt][2];
      w[(2 * q + 1) * Cfg::OPAD + d + 8] = o[mt][3];
    }
    if (g == 0) {
      merge_m[warp * kHeadsPerCta + 2 * q] = m_run[0];
      merge_m[warp * kHeadsPerCta + 2 * q + 1] = m_run[1];
      merge_l[warp * kHeadsPerCta + 2 * q] = l_run[0];
      merge_l[warp * kHeadsPerCta + 2 * q + 1] = l_run[1];
    }
  }
  __syncthreads();

  T* outp = reinterpret_cast<T*>(p.out);
  // cluster launch: every CTA's ring is free now (its own merge buffer lives there), so peers may write into the leader
  if (!PARTITIONED && cs > 1) cluster_sync_all();
  float* part = mo + kConsumerWarps * kHeadsPerCta * Cfg::OPAD;   // leader: [cs][8 heads][OPAD], cols D / D+1 = (m, l)
  for (int idx = threadIdx.x; idx < nheads * D; idx += kFastThreads) {
    const int h = idx / D, d = idx % D;
    float mg = -INFINITY;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) mg = fmaxf(mg, merge_m[w * kHeadsPerCta + h]);
    float num = 0.f, den = 0.f;
    if (mg != -INFINITY) {
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        const float f = fast_exp2(merge_m[w * kHeadsPerCta + h] - mg);
        num += f * mo[(w * kHeadsPerCta + h) * Cfg::OPAD + d];
        den += f * merge_l[w * kHeadsPerCta + h];
      }
    }
    if (!PARTITIONED && cs > 1) {
      float* slot = part + ((int)crank * kHeadsPerCta + h) * Cfg::OPAD;
      st_cluster_f32(slot + d, 0, num);
      if (d == 0) {
        st_cluster_f32(slot + D, 0, mg);
        st_cluster_f32(slot + D + 1, 0, den);
      }
      continue;
    }
    float val = num * __fdividef(1.f, den + 1e-6f);
    if (KV != B200_KV_AUTO) val *= p.v_scale;
    const int head = head0 + h;
    if (PARTITIONED) {
      const int64_t base = ((int64_t)seq * p.num_heads + head) * p.max_num_partitions + part_idx;
      outp[base * D + d] = from_f32<T>(val);
      if (d == 0) {
        p.max_logits[base] = (mg == -INFINITY) ? -FLT_MAX : mg * kLn2;
        p.exp_sums[base] = den;
      }
    } else {
      outp[((int64_t)seq * p.num_heads + head) * D + d] = from_f32<T>(val);
    }
  }
  if (!PARTITIONED && cs > 1) {
    cluster_sync_all();                       // all partials have landed in the leader's shared memory
    if (crank == 0) {
      for (int idx = threadIdx.x; idx < nheads * D; idx += kFastThreads) {
        const int h = idx / D, d = idx % D;
        float mg = -INFINITY;
        for (uint32_t r = 0; r < cs; ++r) mg = fmaxf(mg, part[(r * kHeadsPerCta + h) * Cfg::OPAD + D]);
        float num = 0.f, den = 0.f;
        if (mg != -INFINITY) {
          for (uint32_t r = 0; r < cs; ++r) {
            const float* slot = part + (r * kHeadsPerCta + h) * Cfg::OPAD;
            const float f = fast_exp2(slot[D] - mg);
            num += f * slot[d];
            den += f * slot[D + 1];
          }
        }
        float val = num * __fdividef(1.f, den + 1e-6f);
        if (KV != B200_KV_AUTO) val *= p.v_scale;
        outp[((int64_t)seq * p.num_heads + head0 + h) * D + d] = from_f32<T>(val);
      }
    }
  }
}

// =================================================================================================
// Generic SIMT kernel: any head size / block size / dtype, ALiBi, block-sparse, fp8 KV.
// One CTA per (q-head, seq[, partition]); tokens are processed in segments whose logits live in shared
// memory, with an online-softmax rescale between segments.
// =================================================================================================
static constexpr int kGenThreads = 128;
static constexpr int kGenSegment = 1024;  // tokens per segment
static constexpr int kGenMaxAcc = 4;      // head_size <= kGenThreads * kGenMaxAcc

template <typename T, int KV>
__device__ __forceinline__ float load_cache_elt(const void* base, int64_t idx, float scale) {
  if constexpr (KV == B200_KV_AUTO) {
    return to_f32<T>(reinterpret_cast<const T*>(base)[idx]);
  } else {
    // dequantise to T first (the reference materialises K/V in scalar_t), then widen
    return to_f32<T>(fp8_dequant<T, KV>(reinterpret_cast<const uint8_t*>(base)[idx], scale));
  }
}

template <typename T, int KV, bool PARTITIONED>
__global__ void __launch_bounds__(kGenThreads)
paged_attention_generic_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t gsm[];
  const int D = p.head_size, BS = p.block_size;
  float* q_s = reinterpret_cast<float*>(gsm);  // [D]
  float* lg = q_s + ((D + 3) & ~3);             // [kGenSegment]
  __shared__ float red[kGenThreads / 32];
  __shared__ float bcast;

  const int head = blockIdx.x, seq = blockIdx.y;
  const int part_idx = PARTITIONED ? blockIdx.z : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seq_len = p.seq_lens[seq];
  if (PARTITIONED && part_idx * kPartitionSize >= seq_len) return;
  const int G = p.num_heads / p.num_kv_heads;
  const int kvh = head / G;
  constexpr int ESZ = (KV == B200_KV_AUTO) ? (int)sizeof(T) : 1;
  const int x = 16 / ESZ;

  const int tok_begin = PARTITIONED ? part_idx * kPartitionSize : 0;
  const int tok_end = PARTITIONED ? min(tok_begin + kPartitionSize, seq_len) : seq_len;

  const T* qp = reinterpret_cast<const T*>(p.q) + (int64_t)seq * p.q_stride + (int64_t)head * D;
  for (int i = tid; i < D; i += kGenThreads) q_s[i] = to_f32<T>(qp[i]);
  __syncthreads();

  const float slope = p.alibi_slopes ? p.alibi_slopes[head] : 0.f;
  const int32_t* bt = p.block_tables + (int64_t)seq * p.max_num_blocks_per_seq;
  const bool sparse = p.bs_vert_stride > 1;
  int bs_off = 0, q_bs_id = 0;
  if (sparse) {
    q_bs_id = (seq_len - 1) / p.bs_block_size;
    if (p.bs_head_sliding_step >= 0)
      bs_off = (p.tp_rank * p.num_heads + head) * p.bs_head_sliding_step + 1;
    else
      bs_off = (p.tp_rank * p.num_kv_heads + kvh) * (-p.bs_head_sliding_step) + 1;
  }

  float m_run = -FLT_MAX, l_run = 0.f;
  float acc[kGenMaxAcc];
#pragma unroll
  for (int i = 0; i < kGenMaxAcc; ++i) acc[i] = 0.f;

  for (int seg0 = tok_begin; seg0 < tok_end; seg0 += kGenSegment) {
    const int seg_n = min(kGenSegment, tok_end - seg0);
    // ---- logits: one thread per token ----
    float lmax = -FLT_MAX;
    for (int j = tid; j < seg_n; j += kGenThreads) {
      const int tok = seg0 + j;
      const int bidx = tok / BS, boff = tok % BS;
      bool attend = true;
      if (sparse) {
        const int kb = bidx * BS / p.bs_block_size;
        attend = ((kb + bs_off) % p.bs_vert_stride == 0) || (kb > q_bs_id - p.bs_local_blocks);
      }
      float v = -FLT_MAX;
      if (attend) {
        const int64_t base = (int64_t)bt[bidx] * p.kv_block_stride + (int64_t)kvh * p.kv_head_stride;
        float dot = 0.f;
        for (int c = 0; c < D / x; ++c) {
          const int64_t off = base + ((int64_t)c * BS + boff) * x;
          for (int e = 0; e < x; ++e)
            dot = fmaf(q_s[c * x + e], load_cache_elt<T, KV>(p.k_cache, off + e, p.k_scale), dot);
        }
        v = dot * p.scale;
        if (slope != 0.f) v += slope * (float)(tok - seq_len + 1);
        lmax = fmaxf(lmax, v);
      }
      lg[j] = v;
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    if (tid == 0) {
      float mm = red[0];
      for (int w = 1; w < kGenThreads / 32; ++w) mm = fmaxf(mm, red[w]);
      bcast = mm;
    }
    __syncthreads();
    const float m_new = fmaxf(m_run, bcast);
    const float alpha = __expf(m_run - m_new);
    m_run = m_new;
    // ---- probabilities ----
    float lsum = 0.f;
    for (int j = tid; j < seg_n; j += kGenThreads) {
      const float l = lg[j];
      const float e = (l == -FLT_MAX) ? 0.f : __expf(l - m_new);  // skipped (block-sparse) tokens
      lsum += e;
      lg[j] = to_f32<T>(from_f32<T>(e));  // P is rounded to scalar_t before P.V (reference :395-397)
    }
    lsum = warp_sum(lsum);
    __syncthreads();  // red/bcast reuse + lg visible
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kGenThreads / 32; ++w) tot += red[w];
    l_run = l_run * alpha + tot;
    // ---- P.V : thread <-> output dims ----
#pragma unroll
    for (int a = 0; a < kGenMaxAcc; ++a) {
      const int d = tid + a * kGenThreads;
      if (d < D) {
        float sum = 0.f;
        for (int j = 0; j < seg_n;) {
          const int tok = seg0 + j;
          const int bidx = tok / BS, boff = tok % BS;
          const int n = min(BS - boff, seg_n - j);
          bool attend = true;
          if (sparse) {
            const int kb = bidx * BS / p.bs_block_size;
            attend = ((kb + bs_off) % p.bs_vert_stride == 0) || (kb > q_bs_id - p.bs_local_blocks);
          }
          if (attend) {
            const int64_t base = (int64_t)bt[bidx] * p.kv_block_stride +
                                 (int64_t)kvh * p.kv_head_stride + (int64_t)d * BS + boff;
            for (int e = 0; e < n; ++e) {
              const float pv = lg[j + e];
              if (pv != 0.f) sum = fmaf(pv, load_cache_elt<T, KV>(p.v_cache, base + e, p.v_scale), sum);
            }
          }
          j += n;
        }
        acc[a] = acc[a] * alpha + sum;
      }
    }
    __syncthreads();
  }

  const float inv = __fdividef(1.f, l_run + 1e-6f);
  T* outp = reinterpret_cast<T*>(p.out);
  int64_t obase;
  if (PARTITIONED) {
    const int64_t b = ((int64_t)seq * p.num_heads + head) * p.max_num_partitions + part_idx;
    obase = b * D;
    if (tid == 0) {
      p.max_logits[b] = m_run;
      p.exp_sums[b] = l_run;
    }
  } else {
    obase = ((int64_t)seq * p.num_heads + head) * D;
  }
#pragma unroll
  for (int a = 0; a < kGenMaxAcc; ++a) {
    const int d = tid + a * kGenThreads;
    if (d < D) outp[obase + d] = from_f32<T>(acc[a] * inv);
  }
}

// v2 reduce: merge the per-partition results (reference formula, attention_kernels.cu:565-669)
template <typename T>
__global__ void __launch_bounds__(128)
paged_attention_reduce_kernel(T* __restrict__ out, const float* __restrict__ exp_sums,
                              const float* __restrict__ max_logits, const T* __restrict__ tmp_out,
                              const int32_t* __restrict__ seq_lens, int num_heads, int D,
                              int max_num_partitions) {
  extern __shared__ float rs[];  // [num_partitions] weights
  __shared__ float red[4];
  const int head = blockIdx.x, seq = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seq_len = seq_lens[seq];
  const int np = (seq_len + kPartitionSize - 1) / kPartitionSize;
  const int64_t b = ((int64_t)seq * num_heads + head) * max_num_partitions;
  T* o = out + ((int64_t)seq * num_heads + head) * D;
  const T* t = tmp_out + b * D;
  if (np <= 1) {
    if (np == 1)
      for (int i = tid; i < D; i += blockDim.x) o[i] = t[i];
    return;
  }
  float mx = -FLT_MAX;
  for (int i = tid; i < np; i += blockDim.x) mx = fmaxf(mx, max_logits[b + i]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < np; i += blockDim.x) {
    const float w = exp_sums[b + i] * expf(max_logits[b + i] - mx);
    rs[i] = w;
    sum += w;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = red[0] + red[1] + red[2] + red[3];
  const float inv = __fdividef(1.f, sum + 1e-6f);
  for (int i = tid; i < D; i += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < np; ++j) acc += to_f32<T>(t[(int64_t)j * D + i]) * rs[j] * inv;
    o[i] = from_f32<T>(acc);
  }
}

// =================================================================================================
// host side
// =================================================================================================
static thread_local int g_force_impl = 0;
static thread_local int g_last_path = 0;
static thread_local int g_last_cluster_split = 1;

template <typename T, int D, int BS, int KV, bool PART>
static int launch_tc(const AttnParams& p, cudaStream_t st) {
  using Cfg = FastCfg<D, BS, KV>;
  auto kern = paged_attention_tc_kernel<T, D, BS, KV, PART>;
  static thread_local uint64_t attr_done = 0;  // per-device bit: opt-in smem size already set
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
    attr_done |= 1ull << (dev & 63);
  }
  const int G = p.num_heads / p.num_kv_heads;
  const int groups = (G + kHeadsPerCta - 1) / kHeadsPerCta;
  dim3 grid(p.num_kv_heads * groups, p.num_seqs, PART ? p.max_num_partitions : 1);
  int cs = 1;
  if constexpr (!PART) {
    // split every sequence over a cluster of cs CTAs when the launch would otherwise leave SM slots empty (fewer
    // CTAs than the 2-per-SM resident set: small batches / few kv-heads per GPU) and every CTA still streams >= 64
    // blocks. A grid of more than one wave is NOT split: the kernel is HBM-bound and a partially filled last wave still
    // saturates the memory system (measured: 1024 CTAs at 0.99-1.04 of the HBM peak unsplit, 0.99 split).
    // g_force_impl 2 / 4 force a split (tests)
    const long long n = (long long)grid.x * grid.y, resident = 2LL * num_sms();
    const int max_blocks = p.max_seq_len > 0 ? std::min((p.max_seq_len + BS - 1) / BS, p.max_num_blocks_per_seq)
                                             : p.max_num_blocks_per_seq;
    auto fits = [&](int c) { return Cfg::MERGE_BYTES + c * kHeadsPerCta * Cfg::OPAD * 4 <= Cfg::DATA_BYTES; };
    if (g_force_impl == 2 || g_force_impl == 4) {
      cs = fits(g_force_impl) ? g_force_impl : 1;
    } else {
      // fp8 rings hold half the bytes per CTA (48-64 KB in flight): there a thin last wave does NOT saturate HBM and
      // the split that evens out the waves pays (measured, 1024 CTAs, 1 kv-head, fp8: 0.83 -> 0.90 of the HBM peak)
      const bool thin = Cfg::RING_BYTES < 80 * 1024;
      auto eff = [&](int c) { const long long m = n * c; return (double)m / (double)(((m + resident - 1) / resident) * resident); };
      for (int c = 2; c <= kMaxClusterSplit; c *= 2)
        if (fits(c) && max_blocks / c >= 64 && (n * c <= resident || (thin && eff(c) > eff(cs) + 0.04))) cs = c;
    }
  }
  if (cs == 1) {
    kern<<<grid, kFastThreads, Cfg::SMEM_BYTES, st>>>(p);
  } else {
    grid.z = cs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kFastThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = cs;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  }
  g_last_cluster_split = cs;
  return check_launch("paged_attention_tc_kernel");
}

template <typename T, int KV, bool PART>
static int dispatch_tc(const AttnParams& p, cudaStream_t st, bool& handled) {
  handled = true;
#define B200_TC_CASE(DD, BB) \
  if (p.head_size == DD && p.block_size == BB) return launch_tc<T, DD, BB, KV, PART>(p, st);
  B200_TC_CASE(64, 16)
  B200_TC_CASE(80, 16)
  B200_TC_CASE(96, 16)
  B200_TC_CASE(112, 16)
  B200_TC_CASE(128, 16)
  B200_TC_CASE(192, 16)
  B200_TC_CASE(256, 16)
  B200_TC_CASE(64, 32)
  B200_TC_CASE(80, 32)
  B200_TC_CASE(96, 32)
  B200_TC_CASE(112, 32)
  B200_TC_CASE(128, 32)
  B200_TC_CASE(192, 32)
  B200_TC_CASE(256, 32)
#undef B200_TC_CASE
  handled = false;
  return 0;
}

template <typename T, int KV, bool PART>
static int launch_generic(const AttnParams& p, cudaStream_t st) {
  auto kern = paged_attention_generic_kernel<T, KV, PART>;
  const size_t smem = (((size_t)p.head_size + 3) & ~(size_t)3) * 4 + kGenSegment * 4;
  dim3 grid(p.num_heads, p.num_seqs, PART ? p.max_num_partitions : 1);
  kern<<<grid, kGenThreads, smem, st>>>(p);
  return check_launch("paged_attention_generic_kernel");
}

static bool tc_eligible(const AttnParams& p, int dtype) {
  if (g_force_impl == 1) return false;
  if (dtype == B200_F32) return false;
  if (p.bs_vert_stride > 1) return false;                 // block-sparse -> generic
  if (p.head_size % 16 != 0) return false;
  if (p.block_size != 16 && p.block_size != 32) return false;
  const int G = p.num_heads / p.num_kv_heads;
  if (G * p.num_kv_heads != p.num_heads) return false;
  // bulk copies need 16-B aligned chunks; 32-bit q loads need 4-B alignment
  if ((reinterpret_cast<uintptr_t>(p.k_cache) | reinterpret_cast<uintptr_t>(p.v_cache)) & 15) return false;
  if (reinterpret_cast<uintptr_t>(p.q) & 3) return false;
  if (p.q_stride & 1) return false;
  return true;
}

template <bool PART>
static int run_attention(const AttnParams& p, int dtype, int kv_dtype, cudaStream_t st) {
  B200_CHECK(dtype == B200_F32 || dtype == B200_F16 || dtype == B200_BF16,
             "Unsupported data type of query");
  B200_CHECK(kv_dtype >= B200_KV_AUTO && kv_dtype <= B200_KV_FP8_E5M2,
             "Unsupported data type of kv cache");
  B200_CHECK(p.block_size == 8 || p.block_size == 16 || p.block_size == 32,
             "Unsupported block size: " + std::to_string(p.block_size));
  switch (p.head_size) {
    case 64: case 80: case 96: case 112: case 120: case 128: case 192: case 256: break;
    default: return fail("Unsupported head size: " + std::to_string(p.head_size));
  }
  B200_CHECK(p.num_kv_heads > 0 && p.num_heads % p.num_kv_heads == 0,
             "num_heads must be a multiple of num_kv_heads");
  if (p.num_seqs == 0) return 0;
  const int esz = kv_dtype == B200_KV_AUTO ? (dtype == B200_F32 ? 4 : 2) : 1;
  const int64_t chunk = (int64_t)p.head_size * p.block_size;
  bool tc = tc_eligible(p, dtype);
  // the chunk of one (block, kv-head) must be contiguous and 16-B aligned for the bulk copies
  if (tc && ((p.kv_head_stride * esz) % 16 != 0 || (p.kv_block_stride * esz) % 16 != 0 ||
             p.kv_head_stride < chunk))
    tc = false;
  g_last_path = tc ? 1 : 0;

#define B200_DISPATCH_KV(T)                                                                \
  switch (kv_dtype) {                                                                      \
    case B200_KV_AUTO:                                                                     \
      if (tc) { rc = dispatch_tc<T, B200_KV_AUTO, PART>(p, st, handled); }                 \
      if (!tc || !handled) { g_last_path = 0; rc = launch_generic<T, B200_KV_AUTO, PART>(p, st); } \
      break;                                                                               \
    case B200_KV_FP8_E4M3:                                                                 \
      if (tc) { rc = dispatch_tc<T, B200_KV_FP8_E4M3, PART>(p, st, handled); }             \
      if (!tc || !handled) { g_last_path = 0; rc = launch_generic<T, B200_KV_FP8_E4M3, PART>(p, st); } \
      break;                                                                               \
    default:                                                                               \
      if (tc) { rc = dispatch_tc<T, B200_KV_FP8_E5M2, PART>(p, st, handled); }             \
      if (!tc || !handled) { g_last_path = 0; rc = launch_generic<T, B200_KV_FP8_E5M2, PART>(p, st); } \
      break;                                                                               \
  }

  int rc = 0;
  bool handled = false;
  if (dtype == B200_BF16) {
    B200_DISPATCH_KV(__nv_bfloat16)
  } else if (dtype == B200_F16) {
    B200_DISPATCH_KV(__half)
  } else {
    g_last_path = 0;
    switch (kv_dtype) {
      case B200_KV_AUTO: rc = launch_generic<float, B200_KV_AUTO, PART>(p, st); break;
      case B200_KV_FP8_E4M3: rc = launch_generic<float, B200_KV_FP8_E4M3, PART>(p, st); break;
      default: rc = launch_generic<float, B200_KV_FP8_E5M2, PART>(p, st); break;
    }
  }
#undef B200_DISPATCH_KV
  return rc;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_set_attention_impl(int impl) {
  int prev = g_force_impl;
  g_force_impl = impl;
  return prev;
}
extern "C" int b200_last_attention_path(void) { return g_last_path; }
extern "C" int b200_last_attention_cluster_split(void) { return g_last_cluster_split; }

extern "C" int b200_paged_attention_v1(
    void* out, const void* query, const void* key_cache, const void* value_cache, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, int block_size, float scale,
    const int32_t* block_tables, const int32_t* seq_lens, int max_num_blocks_per_seq,
    int max_seq_len, const float* alibi_slopes, int64_t q_stride, int64_t kv_block_stride,
    int64_t kv_head_stride, int dtype, int kv_dtype, float k_scale, float v_scale, int tp_rank,
    int blocksparse_local_blocks, int blocksparse_vert_stride, int blocksparse_block_size,
    int blocksparse_head_sliding_step, void* stream) {
  AttnParams p{};
  p.max_seq_len = max_seq_len;  // the reference sizes its logits buffer with it; here it only guides the cluster split
  p.out = out; p.q = query; p.k_cache = key_cache; p.v_cache = value_cache;
  p.block_tables = block_tables; p.seq_lens = seq_lens; p.alibi_slopes = alibi_slopes;
  p.num_seqs = num_seqs; p.num_heads = num_heads; p.num_kv_heads = num_kv_heads;
  p.head_size = head_size; p.block_size = block_size;
  p.max_num_blocks_per_seq = max_num_blocks_per_seq; p.max_num_partitions = 0;
  p.q_stride = q_stride; p.kv_block_stride = kv_block_stride; p.kv_head_stride = kv_head_stride;
  p.scale = scale; p.k_scale = k_scale; p.v_scale = v_scale;
  p.tp_rank = tp_rank; p.bs_local_blocks = blocksparse_local_blocks;
  p.bs_vert_stride = blocksparse_vert_stride; p.bs_block_size = blocksparse_block_size;
  p.bs_head_sliding_step = blocksparse_head_sliding_step;
  return run_attention<false>(p, dtype, kv_dtype, (cudaStream_t)stream);
}

extern "C" int b200_paged_attention_v2(
    void* out, float* exp_sums, float* max_logits, void* tmp_out, const void* query,
    const void* key_cache, const void* value_cache, int num_seqs, int num_heads, int num_kv_heads,
    int head_size, int block_size, float scale, const int32_t* block_tables,
    const int32_t* seq_lens, int max_num_blocks_per_seq, int max_seq_len, int max_num_partitions,
    const float* alibi_slopes, int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
    int dtype, int kv_dtype, float k_scale, float v_scale, int tp_rank,
    int blocksparse_local_blocks, int blocksparse_vert_stride, int blocksparse_block_size,
    int blocksparse_head_sliding_step, void* stream) {
  (void)max_seq_len;
  B200_CHECK(max_num_partitions >= 1, "max_num_partitions must be >= 1");
  AttnParams p{};
  p.out = tmp_out; p.exp_sums = exp_sums; p.max_logits = max_logits;
  p.q = query; p.k_cache = key_cache; p.v_cache = value_cache;
  p.block_tables = block_tables; p.seq_lens = seq_lens; p.alibi_slopes = alibi_slopes;
  p.num_seqs = num_seqs; p.num_heads = num_heads; p.num_kv_heads = num_kv_heads;
  p.head_size = head_size; p.block_size = block_size;
  p.max_num_blocks_per_seq = max_num_blocks_per_seq; p.max_num_partitions = max_num_partitions;
  p.q_stride = q_stride; p.kv_block_stride = kv_block_stride; p.kv_head_stride = kv_head_stride;
  p.scale = scale; p.k_scale = k_scale; p.v_scale = v_scale;
  p.tp_rank = tp_rank; p.bs_local_blocks = blocksparse_local_blocks;
  p.bs_vert_stride = blocksparse_vert_stride; p.bs_block_size = blocksparse_block_size;
  p.bs_head_sliding_step = blocksparse_head_sliding_step;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run_attention<true>(p, dtype, kv_dtype, st);
  if (rc != 0 || num_seqs == 0) return rc;
  dim3 grid(num_heads, num_seqs);
  const size_t smem = (size_t)max_num_partitions * sizeof(float);
  if (dtype == B200_BF16) {
    paged_attention_reduce_kernel<__nv_bfloat16><<<grid, 128, smem, st>>>(
        (__nv_bfloat16*)out, exp_sums, max_logits, (const __nv_bfloat16*)tmp_out, seq_lens,
        num_heads, head_size, max_num_partitions);
  } else if (dtype == B200_F16) {
    paged_attention_reduce_kernel<__half><<<grid, 128, smem, st>>>(
        (__half*)out, exp_sums, max_logits, (const __half*)tmp_out, seq_lens, num_heads,
        head_size, max_num_partitions);
  } else {
    paged_attention_reduce_kernel<float><<<grid, 128, smem, st>>>(
        (float*)out, exp_sums, max_logits, (const float*)tmp_out, seq_lens, num_heads, head_size,
        max_num_partitions);
  }
  return check_launch("paged_attention_reduce_kernel");
}
