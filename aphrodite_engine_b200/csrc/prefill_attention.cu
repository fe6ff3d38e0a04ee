// prefill_attention.cu — prefix-aware prefill attention over the paged KV cache, sm_100a.
//
// Replaces context_attention_fwd (aphrodite/attention/ops/prefix_prefill.py:696-858; Triton kernels _fwd_kernel :13-255
// and _fwd_kernel_alibi :448-694), called by PagedAttention.forward_prefix (attention/ops/paged_attn.py:192-228) from the
// xformers / flash backends when a prefill batch has cached context (backends/xformers.py:588-625):
//   for sequence b, query token m (position ctx_len[b] + m), head h:
//     out = softmax( scale * q . [K_cache(context) ; k_new(0..m)] (+ alibi, sliding window) ) . [V_cache ; v_new]
//   context keys / values come from the PAGED cache through the block table (layouts of paged_attention.cu:
//   key_cache [NB, Hkv, D/x, BS, x], value_cache [NB, Hkv, D, BS]), optionally fp8 (dequantised with k_scale / v_scale
//   exactly as the reference: fp8 -> fp32 * scale -> T, prefix_prefill.py:126-129,166-169); the new keys / values are the
//   contiguous k / v of this step, causal among themselves.
// Masks as the reference: positions beyond the context / the causal frontier get -inf; sliding-window violations get
// the finite -10000 of the reference (:147-149) — same softmax weights (exp underflows to 0 either way).
//
// Design: FlashAttention-2 style on mma.sync.m16n8k16 (fp32 accumulate), 4 warps x 16 query rows per CTA, 64-key tiles,
// two shared-memory stages filled with cp.async (16-byte pieces straight out of the paged layouts: a K piece is one
// x-run of one token, a V piece is 8 consecutive tokens of one head-dim row), both cache layouts consumed AS THEY LIE:
//   S = Q.K^T : B fragments by ldmatrix from K rows [token][d]            (paged K runs land as rows)
//   O += P.V  : B fragments by ldmatrix from V^T rows [d][token] (paged V, no transpose needed) or by
//               ldmatrix.trans from new V rows [token][d]
// Rows are padded by 16 bytes, so every 8-row ldmatrix phase touches 8 distinct bank groups for every head size.
// Online softmax in the log2 domain, P rounded to T before P.V, normalisation deferred to the end (the reference
// renormalises P every tile, :152-162 — same value up to rounding; tolerances in tests/tolerances.py).
// This is the tensor-core baseline of the row; a tcgen05 / TMEM version (S and O accumulators in TMEM, K/V tiles by
// TMA gather) is the next step recorded in DESIGN.md.
#include "common.cuh"

#include <type_traits>

namespace b200 {

static constexpr int PF_BM = 64;      // query rows per CTA (16 per warp)
static constexpr int PF_BN = 64;      // keys per tile
static constexpr int PF_THREADS = 128;
static constexpr int PF_VT_STRIDE = PF_BN + 8;   // elements per row of the transposed V tile ([d][token], +16 B pad)

struct PrefillParams {
  const void* q; const void* k; const void* v; void* out;
  const void* key_cache; const void* value_cache;
  const int32_t* block_tables; const int32_t* start_loc; const int32_t* seq_lens; const int32_t* ctx_lens;
  const float* alibi_slopes;
  int64_t q_stride_t, q_stride_h, k_stride_t, k_stride_h, v_stride_t, v_stride_h, o_stride_t, o_stride_h;
  int64_t kc_block_stride, kc_head_stride, vc_block_stride, vc_head_stride;   // elements of the cache type
  int64_t bt_stride;
  int num_heads, num_kv_heads, block_size, x;
  int bs_shift, x_shift;     // log2 of block_size / x (both powers of two): the tile loaders never divide
  int sliding_window;
  float scale, k_scale, v_scale;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t saddr, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void pf_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void pf_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(saddr));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 8 cache elements -> 8 T (16 bytes). 16-bit cache: a copy. fp8: fp8 -> half (exact) -> fp32 * scale -> T
template <typename T, int KV>
__device__ __forceinline__ uint4 load_cache8(const void* src, float scale) {
  if constexpr (KV == B200_KV_AUTO) {
    return *reinterpret_cast<const uint4*>(src);
  } else {
    const uint2 raw = *reinterpret_cast<const uint2*>(src);
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&raw);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      ow[i] = pack2<T>(fp8_to_f32<KV>(b[2 * i]) * scale, fp8_to_f32<KV>(b[2 * i + 1]) * scale);
    return o;
  }
}

template <typename T, int DT, int KV>
__global__ void __launch_bounds__(PF_THREADS, (DT == 8 || DT == 4 ? 3 : 1)) prefill_attention_kernel(const PrefillParams p) {
  constexpr int D = DT * 16;
  constexpr int KSTR = D + 8;                                  // elements per K / new-V row (+16 B pad)
  constexpr int K_TILE = PF_BN * KSTR;                         // elements
  constexpr int V_TILE = (D * PF_VT_STRIDE > K_TILE) ? D * PF_VT_STRIDE : K_TILE;
  constexpr int CH = D / 8;                                    // 16-byte chunks per row
  using CT = typename std::conditional<KV == B200_KV_AUTO, T, uint8_t>::type;
  extern __shared__ __align__(16) uint8_t pf_smem[];
  T* sm = reinterpret_cast<T*>(pf_smem);                       // [2 stages][K tile | V tile]

  const int mt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int ctx_len = p.ctx_lens[b];
  const int q_len = p.seq_lens[b] - ctx_len;
  if (mt * PF_BM >= q_len) return;
  const int start = p.start_loc[b];
  const int kvh = head / (p.num_heads / p.num_kv_heads);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int row0 = mt * PF_BM + warp * 16 + g, row1 = row0 + 8;

  // ---- Q fragments (A operand, 16 rows x D), straight from global ------------------------------------------
  uint32_t qf[DT][4];
  {
    const T* qb = reinterpret_cast<const T*>(p.q) + (size_t)head * p.q_stride_h;
    const T* q0 = qb + (size_t)(start + row0) * p.q_stride_t;
    const T* q1 = qb + (size_t)(start + row1) * p.q_stride_t;
    const bool v0 = row0 < q_len, v1 = row1 < q_len;
#pragma unroll
    for (int kt = 0; kt < DT; ++kt) {
      const int d0 = kt * 16 + 2 * t4;
      qf[kt][0] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + d0) : 0u;
      qf[kt][1] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + d0) : 0u;
      qf[kt][2] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + d0 + 8) : 0u;
      qf[kt][3] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + d0 + 8) : 0u;
    }
  }

  float o[2 * DT][4];
#pragma unroll
  for (int i = 0; i < 2 * DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  constexpr float LOG2E = 1.4426950408889634f;
  const float scale2 = p.scale * LOG2E;
  const float slope2 = p.alibi_slopes ? p.alibi_slopes[head] * LOG2E : 0.f;
  const bool has_alibi = p.alibi_slopes != nullptr;
  const int sw = p.sliding_window;
  const int BS = p.block_size;
  const int n_ctx = (ctx_len + PF_BN - 1) / PF_BN;
  const int n_tiles = n_ctx + mt + 1;                          // new-token tiles 0..mt (causal frontier)
  const int32_t* bt = p.block_tables + (size_t)b * p.bt_stride;
  const CT* kc = reinterpret_cast<const CT*>(p.key_cache) + (size_t)kvh * p.kc_head_stride;
  const CT* vc = reinterpret_cast<const CT*>(p.value_cache) + (size_t)kvh * p.vc_head_stride;
  const T* kn = reinterpret_cast<const T*>(p.k) + (size_t)kvh * p.k_stride_h;
  const T* vn = reinterpret_cast<const T*>(p.v) + (size_t)kvh * p.v_stride_h;

  // Tile loaders. All per-thread piece coordinates are fixed outside the tile loop (128 threads, pieces strided by
  // 128): a thread always serves the same token (K) / the same 8-token group (V) of a tile, so a tile costs it ONE
  // block-table lookup per operand and shift / add addressing (the first version recomputed div / mod by the runtime
  // block size for every 16-byte piece: 27 % IMAD + 14 % ISETP + the IABS / MUFU.RCP of emulated division against 5 % HMMA
  // in the executed-instruction mix, profiles/r02_prefill_attention_first_ncu.txt).
  const int bs_shift = p.bs_shift, bs_mask = BS - 1, x_shift = p.x_shift, x_mask = p.x - 1;
  const int k_tl = tid & (PF_BN - 1), k_c0 = tid >> 6;          // ctx K: token of the tile, first chunk (chunks step 2)
  const int v_tg = tid & 7, v_d0 = tid >> 3;                    // ctx V: 8-token group of the tile, first head-dim row (rows step 16)
  auto load_tile = [&](int tile, int stage) {
    T* ks = sm + (size_t)stage * (K_TILE + V_TILE);
    T* vs = ks + K_TILE;
    if (tile < n_ctx) {
      const int tile_start = tile * PF_BN;
      {
        // K: piece (chunk c of 8 head-dim elements, token) = half or all of one x-run of the paged layout
        const int pos = tile_start + k_tl;
        const bool valid = pos < ctx_len;
        const int blk = valid ? bt[pos >> bs_shift] : 0;
        // chunk c = k_c0 + 2r covers elements 8c .. 8c+7: run (8c >> x_shift), offset (8c & x_mask); stepping c by 2 advances
        // 16 elements = 16 << bs_shift cache elements whatever x is
        const int e00 = k_c0 * 8;
        const CT* src = kc + (size_t)blk * p.kc_block_stride + (size_t)((valid ? pos & bs_mask : 0) << x_shift) +
                        ((size_t)(e00 >> x_shift) << (bs_shift + x_shift)) + (e00 & x_mask);
        const size_t kstep = (size_t)16 << bs_shift;
        T* dst = ks + k_tl * KSTR + k_c0 * 8;
#pragma unroll
        for (int r = 0; r < CH / 2; ++r) {
          if constexpr (KV == B200_KV_AUTO) {
            cp_async16_zfill(smem_u32(dst + r * 16), src + r * kstep, valid);
          } else {
            *reinterpret_cast<uint4*>(dst + r * 16) = valid ? load_cache8<T, KV>(src + r * kstep, p.k_scale) : make_uint4(0, 0, 0, 0);
          }
        }
      }
      {
        // V: piece = 8 consecutive tokens of one head-dim row inside one cache block
        const int pos0 = tile_start + v_tg * 8;
        const int nvalid = ctx_len - pos0;                      // tokens of the group inside the context
        const int blk = nvalid > 0 ? bt[pos0 >> bs_shift] : 0;
        const CT* vb = vc + (size_t)blk * p.vc_block_stride + (pos0 & bs_mask) + ((size_t)v_d0 << bs_shift);
        const size_t vstep = (size_t)16 << bs_shift;
        T* dst = vs + v_d0 * PF_VT_STRIDE + v_tg * 8;
#pragma unroll
        for (int r = 0; r < D / 16; ++r) {
          const CT* src = vb + r * vstep;
          T* d8 = dst + r * 16 * PF_VT_STRIDE;
          if (nvalid >= 8) {
            if constexpr (KV == B200_KV_AUTO) cp_async16_zfill(smem_u32(d8), src, true);
            else *reinterpret_cast<uint4*>(d8) = load_cache8<T, KV>(src, p.v_scale);
          } else if (nvalid <= 0) {
            *reinterpret_cast<uint4*>(d8) = make_uint4(0, 0, 0, 0);
          } else {
            // partially valid group: the slots beyond the context may hold anything (NaNs included) and 0 * NaN = NaN in
            // the MMA, so they are zeroed (the reference's masked load, prefix_prefill.py:163-165)
            uint4 v = load_cache8<T, KV>(src, p.v_scale);
            uint16_t* e = reinterpret_cast<uint16_t*>(&v);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j >= nvalid) e[j] = 0;
            *reinterpret_cast<uint4*>(d8) = v;
          }
        }
      }
    } else {
      const int tile_start = (tile - n_ctx) * PF_BN;
#pragma unroll
      for (int r = 0; r < PF_BN * CH / PF_THREADS; ++r) {
        const int i = tid + r * PF_THREADS;
        const int tl = i / CH, c = i % CH;                       // CH is a compile-time constant
        const int pos = tile_start + tl;
        const bool valid = pos < q_len;
        const size_t tok = (size_t)(start + (valid ? pos : 0));
        cp_async16_zfill(smem_u32(ks + tl * KSTR + c * 8), kn + tok * p.k_stride_t + c * 8, valid);
        cp_async16_zfill(smem_u32(vs + tl * KSTR + c * 8), vn + tok * p.v_stride_t + c * 8, valid);
      }
    }
    pf_commit();
  };

  load_tile(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int stage = tile & 1;
    if (tile + 1 < n_tiles) {
      load_tile(tile + 1, stage ^ 1);
      pf_wait<1>();
    } else {
      pf_wait<0>();
    }
    __syncthreads();
    const T* ks = sm + (size_t)stage * (K_TILE + V_TILE);
    const T* vs = ks + K_TILE;
    const bool is_ctx = tile < n_ctx;
    const int tile_start = is_ctx ? tile * PF_BN : (tile - n_ctx) * PF_BN;

    // ---- S = Q . K^T ------------------------------------------------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    {
      const int mi = lane >> 3, r = lane & 7;
      const uint32_t kbase = smem_u32(ks) + (uint32_t)(((mi >> 1) * 8 + r) * KSTR + (mi & 1) * 8) * 2u;
#pragma unroll
      for (int kt = 0; kt < DT; ++kt) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t r0, r1, r2, r3;
          ldmatrix_x4(r0, r1, r2, r3, kbase + (uint32_t)(np * 16 * KSTR + kt * 16) * 2u);
          mma_16816<T>(s[2 * np], qf[kt], r0, r1);
          mma_16816<T>(s[2 * np + 1], qf[kt], r2, r3);
        }
      }
    }

    // ---- scale, bias, masks, online softmax ---------------------------------------------------------------------
    float mx[2] = {-INFINITY, -INFINITY};
    // tiles that lie entirely inside the context / strictly below the causal diagonal need no mask at all
    const bool plain = !has_alibi && sw <= 0 && (is_ctx ? (tile_start + PF_BN <= ctx_len) : (tile - n_ctx < mt));
    if (plain) {
      // the scale folds into the exponent's FMA below; the row maximum of the scaled logits is the scaled maximum
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
      mx[0] *= scale2;
      mx[1] *= scale2;
    } else {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int row = (e < 2) ? row0 : row1;
          const int col = tile_start + nt * 8 + 2 * t4 + (e & 1);
          const int qpos = ctx_len + row;
          const int kpos = is_ctx ? col : ctx_len + col;
          const bool valid = is_ctx ? (col < ctx_len) : (col <= row && col < q_len);
          float v = s[nt][e] * scale2;
          if (has_alibi) v += slope2 * (float)(kpos - qpos);
          if (sw > 0 && qpos - kpos >= sw) v = -10000.f * LOG2E;
          if (!valid) v = -INFINITY;
          s[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
    }
    float alpha[2], msafe[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m_run[h], mx[h]);
      msafe[h] = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[h] = ex2(m_run[h] - msafe[h]);                     // m_run = -inf -> 0
      m_run[h] = m_new;
      l_run[h] *= alpha[h];
    }
    uint32_t pa[4][4];
    const float ps = plain ? scale2 : 1.f;                     // masked tiles hold scaled logits already
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ex2(fmaf(s[nt][0], ps, -msafe[0])), p1 = ex2(fmaf(s[nt][1], ps, -msafe[0]));
      const float p2 = ex2(fmaf(s[nt][2], ps, -msafe[1])), p3 = ex2(fmaf(s[nt][3], ps, -msafe[1]));
      l_run[0] += p0 + p1;
      l_run[1] += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack2<T>(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2<T>(p2, p3);
    }
    // the running maximum stops moving after the first few tiles of a row: rescale the output accumulators only when
    // some row of this warp needs it (warp-uniform branch)
    if (__any_sync(0xffffffffu, alpha[0] != 1.f || alpha[1] != 1.f)) {
#pragma unroll
      for (int i = 0; i < 2 * DT; ++i) {
        o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
        o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
      }
    }

    // ---- O += P . V ------------------------------------------------------------------------------------------------
    {
      const int mi = lane >> 3, r = lane & 7;
      if (is_ctx) {
        // V^T tile [d][token]: matrices (d rows, token columns), plain ldmatrix
        const uint32_t vbase = smem_u32(vs) + (uint32_t)(((mi >> 1) * 8 + r) * PF_VT_STRIDE + (mi & 1) * 8) * 2u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int dp = 0; dp < DT; ++dp) {
            uint32_t r0, r1, r2, r3;
            ldmatrix_x4(r0, r1, r2, r3, vbase + (uint32_t)(dp * 16 * PF_VT_STRIDE + j * 16) * 2u);
            mma_16816<T>(o[2 * dp], pa[j], r0, r1);
            mma_16816<T>(o[2 * dp + 1], pa[j], r2, r3);
          }
        }
      } else {
        // new V tile [token][d]: matrices (token rows, d columns), ldmatrix.trans
        const uint32_t vbase = smem_u32(vs) + (uint32_t)(((mi & 1) * 8 + r) * KSTR + (mi >> 1) * 8) * 2u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int dp = 0; dp < DT; ++dp) {
            uint32_t r0, r1, r2, r3;
            ldmatrix_x4_trans(r0, r1, r2, r3, vbase + (uint32_t)(j * 16 * KSTR + dp * 16) * 2u);
            mma_16816<T>(o[2 * dp], pa[j], r0, r1);
            mma_16816<T>(o[2 * dp + 1], pa[j], r2, r3);
          }
        }
      }
    }
    __syncthreads();                                           // stage may be refilled by the next iteration's loads
  }

  // ---- normalise and store -------------------------------------------------------------------------------------------
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  T* ob = reinterpret_cast<T*>(p.out) + (size_t)head * p.o_stride_h;
  if (row0 < q_len) {
    T* o0 = ob + (size_t)(start + row0) * p.o_stride_t;
#pragma unroll
    for (int i = 0; i < 2 * DT; ++i)
      *reinterpret_cast<uint32_t*>(o0 + i * 8 + 2 * t4) = pack2<T>(o[i][0] * inv0, o[i][1] * inv0);
  }
  if (row1 < q_len) {
    T* o1 = ob + (size_t)(start + row1) * p.o_stride_t;
#pragma unroll
    for (int i = 0; i < 2 * DT; ++i)
      *reinterpret_cast<uint32_t*>(o1 + i * 8 + 2 * t4) = pack2<T>(o[i][2] * inv1, o[i][3] * inv1);
  }
}

template <typename T, int DT, int KV>
static int launch_prefill(const PrefillParams& p, dim3 grid, cudaStream_t st) {
  constexpr int D = DT * 16;
  constexpr int KSTR = D + 8;
  constexpr int K_TILE = PF_BN * KSTR;
  constexpr int V_TILE = (D * PF_VT_STRIDE > K_TILE) ? D * PF_VT_STRIDE : K_TILE;
  constexpr size_t smem = (size_t)2 * (K_TILE + V_TILE) * 2;
  auto kern = prefill_attention_kernel<T, DT, KV>;
  static thread_local uint64_t attr_done = 0;
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done |= 1ull << (dev & 63);
  }
  kern<<<grid, PF_THREADS, smem, st>>>(p);
  return check_launch("prefill_attention_kernel");
}

template <typename T, int KV>
static int dispatch_prefill_head(const PrefillParams& p, int head_size, dim3 grid, cudaStream_t st) {
  switch (head_size) {
    case 64: return launch_prefill<T, 4, KV>(p, grid, st);
    case 80: return launch_prefill<T, 5, KV>(p, grid, st);
    case 96: return launch_prefill<T, 6, KV>(p, grid, st);
    case 112: return launch_prefill<T, 7, KV>(p, grid, st);
    case 128: return launch_prefill<T, 8, KV>(p, grid, st);
    case 192: return launch_prefill<T, 12, KV>(p, grid, st);
    case 256: return launch_prefill<T, 16, KV>(p, grid, st);
  }
  return fail("context_attention_fwd: unsupported head size " + std::to_string(head_size) +
              " (supported: 64, 80, 96, 112, 128, 192, 256)");
}

}  // namespace b200

using namespace b200;

extern "C" int b200_context_attention_fwd(
    const void* q, const void* k, const void* v, void* out, const void* key_cache, const void* value_cache,
    const int32_t* block_tables, const int32_t* start_loc, const int32_t* seq_lens, const int32_t* ctx_lens,
    const float* alibi_slopes, int batch, int num_heads, int num_kv_heads, int head_size, int block_size, int x,
    int max_query_len, int64_t q_stride_t, int64_t q_stride_h, int64_t k_stride_t, int64_t k_stride_h,
    int64_t v_stride_t, int64_t v_stride_h, int64_t o_stride_t, int64_t o_stride_h, int64_t kc_block_stride,
    int64_t kc_head_stride, int64_t vc_block_stride, int64_t vc_head_stride, int64_t bt_stride, float scale,
    float k_scale, float v_scale, int sliding_window, int dtype, int kv_dtype, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "context_attention_fwd: query dtype must be float16 or bfloat16");
  B200_CHECK(kv_dtype >= B200_KV_AUTO && kv_dtype <= B200_KV_FP8_E5M2, "context_attention_fwd: unsupported kv cache dtype");
  B200_CHECK(num_kv_heads > 0 && num_heads % num_kv_heads == 0, "num_heads must be a multiple of num_kv_heads");
  B200_CHECK(block_size == 8 || block_size == 16 || block_size == 32 || block_size == 64,
             "context_attention_fwd: block_size must be 8, 16, 32 or 64");
  B200_CHECK(x == 8 || x == 16, "context_attention_fwd: key cache x (innermost run) must be 8 or 16 elements");
  B200_CHECK(q_stride_t % 8 == 0 && q_stride_h % 8 == 0 && k_stride_t % 8 == 0 && k_stride_h % 8 == 0 &&
             v_stride_t % 8 == 0 && v_stride_h % 8 == 0 && o_stride_t % 2 == 0 && o_stride_h % 2 == 0,
             "context_attention_fwd: q / k / v strides must be multiples of 8 elements (16-byte rows)");
  B200_CHECK(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
               reinterpret_cast<uintptr_t>(key_cache) | reinterpret_cast<uintptr_t>(value_cache)) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(out) & 3) == 0, "context_attention_fwd: pointers must be 16-byte aligned");
  const int cache_elt = kv_dtype == B200_KV_AUTO ? 2 : 1;
  B200_CHECK((kc_block_stride * cache_elt) % 16 == 0 && (kc_head_stride * cache_elt) % 16 == 0 &&
             (vc_block_stride * cache_elt) % 16 == 0 && (vc_head_stride * cache_elt) % 16 == 0,
             "context_attention_fwd: cache block / head strides must be multiples of 16 bytes");
  if (batch == 0 || max_query_len <= 0) return 0;
  PrefillParams p{};
  p.q = q; p.k = k; p.v = v; p.out = out; p.key_cache = key_cache; p.value_cache = value_cache;
  p.block_tables = block_tables; p.start_loc = start_loc; p.seq_lens = seq_lens; p.ctx_lens = ctx_lens;
  p.alibi_slopes = alibi_slopes;
  p.q_stride_t = q_stride_t; p.q_stride_h = q_stride_h; p.k_stride_t = k_stride_t; p.k_stride_h = k_stride_h;
  p.v_stride_t = v_stride_t; p.v_stride_h = v_stride_h; p.o_stride_t = o_stride_t; p.o_stride_h = o_stride_h;
  p.kc_block_stride = kc_block_stride; p.kc_head_stride = kc_head_stride;
  p.vc_block_stride = vc_block_stride; p.vc_head_stride = vc_head_stride; p.bt_stride = bt_stride;
  p.num_heads = num_heads; p.num_kv_heads = num_kv_heads; p.block_size = block_size; p.x = x;
  p.bs_shift = block_size == 8 ? 3 : block_size == 16 ? 4 : block_size == 32 ? 5 : 6;
  p.x_shift = x == 8 ? 3 : 4;
  p.sliding_window = sliding_window > 0 ? sliding_window : 0;
  p.scale = scale; p.k_scale = k_scale; p.v_scale = v_scale;
  dim3 grid((max_query_len + PF_BM - 1) / PF_BM, num_heads, batch);
  cudaStream_t st = (cudaStream_t)stream;
#define B200_PF(TT) \
  switch (kv_dtype) { \
    case B200_KV_AUTO: return dispatch_prefill_head<TT, B200_KV_AUTO>(p, head_size, grid, st); \
    case B200_KV_FP8_E4M3: return dispatch_prefill_head<TT, B200_KV_FP8_E4M3>(p, head_size, grid, st); \
    default: return dispatch_prefill_head<TT, B200_KV_FP8_E5M2>(p, head_size, grid, st); \
  }
  if (dtype == B200_BF16) { B200_PF(__nv_bfloat16) }
  B200_PF(__half)
#undef B200_PF
}
