// scaled_mm.cu — W8A8 GEMM with per-tensor / per-token / per-channel scales on 5th-gen tensor cores, sm_100a.
//
// Replaces cutlass_scaled_mm (kernels/quantization/cutlass_w8a8/scaled_mm_entry.cu:92-140; the CUTLASS 2.x / 3.x
// kernels behind it, scaled_mm_c2x.cu / scaled_mm_c3x.cu, stop at sm_90 `wgmma` and do not exist for sm_100):
//     out[M,N] = T( a_scales[m|0] * ( b_scales[n|0] * sum_k a[m,k] * b[k,n] ) + bias[n] )
//   a  [M,K] row-major, fp8-e4m3 or int8        b [K,N] COLUMN-major (= the checkpoint's [N,K] weight, K contiguous)
//   a_scales fp32 [1] or [M], b_scales fp32 [1] or [N], bias T [N] or none, out fp16 / bf16
// Epilogue arithmetic in fp32 in the reference's order (ScaledEpilogue, scaled_mm_c3x.cu: Compute0 = b_scales * acc,
// Compute1 = a_scales * Compute0, with bias: fma(a_scales, Compute0, bias)); one rounding to T.
//
// Design (the transposed formulation of marlin_gemm.cu, without the dequant stage: both operands are already in a
// tensor-core format and K-major, so they go global -> TMA -> shared memory -> tcgen05 untouched):
//   D^T[128 channels, tokens <= 256] += W[128 ch, 128 k] . A[tokens, 128 k]^T     one CTA per (128-channel tile,
//                                                                                  256-token block, k-split)
//   * both tiles by cp.async.bulk.tensor (SWIZZLE_128B; one 128-byte swizzle row = 128 k of 8-bit data), OOB rows and
//     the k tail zero-filled by the TMA unit; a 3..8-stage full/empty mbarrier ring
//   * one elected thread issues 4 x tcgen05.mma.cta_group::1.kind::f8f6f4 (or kind::i8) M=128, N=tokens, K=32 per stage;
//     fp32 (s32) accumulator in TMEM; tcgen05.commit frees the stage
//   * epilogue warps tcgen05.ld their TMEM lane quadrant (lane = channel, column = token), apply the scales / bias,
//     round once and store
//   * k-split (decode shapes have few tiles: 4096 channels = 32 tiles on 148 SMs): the splits of a tile are the CTAs
//     of one thread-block cluster (1, 1, S) — co-resident by construction, so `barrier.cluster` is a rendezvous that
//     cannot deadlock. Every split stores its 32-bit partial tile in its own slab of a caller-provided scratch
//     [S, M, N] (L2-resident: 2 MB per slab at 256 x 2048), the barrier publishes the slabs, and each split reduces an
//     interleaved 1/S share of the token rows over the slabs IN SPLIT ORDER, applies the scales and writes the rows as
//     256-byte coalesced stores: deterministic, no atomics, no second kernel (the scheme of marlin_gemm.cu). A first
//     version pushed the partials into the owner's shared memory with 4-byte st.shared::cluster stores: 4096 x 6144 at
//     256 tokens took 112 us against 69 us for the un-split 4096 x 28672 (profiles/r02_bench_f_rows_first.json) —
//     scattered 4-byte DSMEM stores are issued per thread, not per warp.
#include "common.cuh"
#include "tc5.cuh"

#include <algorithm>

namespace b200 {

static constexpr int SM_NT = 128;            // output channels per CTA (UMMA M)
static constexpr int SM_KC = 128;            // k per stage: one 128-byte swizzle row of 8-bit elements
static constexpr int SM_TOK = 256;           // tokens per CTA (UMMA N)
static constexpr int SM_MAX_STAGES = 8;
static constexpr int SM_W_BYTES = SM_NT * SM_KC;          // 16 KB
static constexpr int SM_THREADS = 192;       // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
static constexpr int SM_SMEM_TOTAL = 224 * 1024;
static constexpr int SM_FIXED = 2048;        // barriers + alignment slack

enum { SMK_FP8 = 0, SMK_INT8 = 1 };

struct ScaledMMParams {
  void* c;                  // [M, N] T, row stride ldc
  const float* a_scales;    // [1] or [M]
  const float* b_scales;    // [1] or [N]
  const void* bias;         // [N] T or nullptr
  int M, N, K;
  int64_t ldc;
  int a_scale_per_token, b_scale_per_channel;
  int box_rows;             // rows of the activation TMA box (tokens rounded up to 16)
  int act_bytes;            // box_rows * 128
  int stages;
  int split_k, chunks_per_split;
  uint32_t* scratch;        // [split_k, M, N] 32-bit partials (fp32, or s32 for int8); used when split_k > 1
};

__device__ __forceinline__ void sm_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// NSUB = 128-channel sub-tiles per CTA (1 or 2). With 2, one activation tile feeds two weight sub-tiles (two TMEM
// accumulators): the activation tile is the larger half of a stage at 256 tokens and every CTA re-reads it from L2, and
// the first version (NSUB = 1 only) ran at the L2 -> SM fabric rate, not at the tensor or DRAM rate (ncu: 352 MB through
// the crossbar for 119 MB of DRAM reads, 5.7 TB/s, tensor pipe 22 % of elapsed; profiles/r02_scaled_mm_fp8_m256_first_ncu.txt).
template <typename T, int KIND, int NSUB>
__global__ void __launch_bounds__(SM_THREADS, 1)
scaled_mm_tc5_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                     const ScaledMMParams p) {
  constexpr int CTA_CH = SM_NT * NSUB;
  constexpr int W_BYTES = CTA_CH * SM_KC;
  extern __shared__ uint8_t sm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NS = p.stages;
  const int stage_bytes = W_BYTES + p.act_bytes;
  uint8_t* tiles = smem;                                     // [NS][ W tile (NSUB x 16 KB) | activation tile act_bytes ]
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)NS * stage_bytes);
  uint64_t* empty = full + SM_MAX_STAGES;
  uint64_t* accum_full = empty + SM_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = blockIdx.x * CTA_CH;
  const int tok_base = blockIdx.y * SM_TOK;
  const int toks = min(SM_TOK, p.M - tok_base);
  const int n_mma = max(16, (toks + 15) & ~15);
  const int total_chunks = (p.K + SM_KC - 1) / SM_KC;
  const int chunk0 = blockIdx.z * p.chunks_per_split;
  const int nchunks = min(p.chunks_per_split, total_chunks - chunk0);     // >= 1 by construction of the split plan
  uint32_t acc_stride = 32;                                   // TMEM columns between the sub-tiles' accumulators
  while ((int)acc_stride < n_mma) acc_stride <<= 1;
  const uint32_t tmem_cols = acc_stride * NSUB;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SM_MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(accum_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t use = 0;
      const uint32_t tx_bytes = (uint32_t)W_BYTES + (uint32_t)p.box_rows * 128u;   // TMA always moves the full boxes
      for (int c = 0; c < nchunks; ++c) {
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1u);
        uint8_t* st = tiles + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], tx_bytes);
        tma_load_2d(st, &tmap_w, &full[s], (chunk0 + c) * SM_KC, n_base);
        tma_load_2d(st + W_BYTES, &tmap_a, &full[s], (chunk0 + c) * SM_KC, tok_base);
        if (++s == NS) { s = 0; ++use; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6) 1 = F32 / 2 = S32; a/b format [7,10),[10,13)
      // 0 = E4M3 (kind::f8f6f4) / 1 = signed 8-bit (kind::i8); both operands K-major; N >> 3 at [17,23); M >> 4 at [24,29)
      const uint32_t idesc = KIND == SMK_FP8
                                 ? ((1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(SM_NT >> 4) << 24))
                                 : ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma >> 3) << 17) |
                                    ((uint32_t)(SM_NT >> 4) << 24));
      int s = 0;
      uint32_t use = 0;
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&full[s], use & 1u);
        tc_fence_after();
        const uint32_t st = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint64_t b_desc = make_sw128_desc(st + W_BYTES);         // activations: UMMA "B" (N = tokens)
#pragma unroll
        for (int ks = 0; ks < SM_KC / 32; ++ks) {
#pragma unroll
          for (int j = 0; j < NSUB; ++j) {
            // weights of sub-tile j: UMMA "A" (M = 128 channels), rows j*128.. of the stage's W tile (128 B per row)
            const uint64_t a_desc = make_sw128_desc(st + (uint32_t)j * (SM_NT * SM_KC));
            // 32 k = 32 bytes further inside the 128-byte swizzle row = +2 in the descriptor's (address >> 4) field
            if constexpr (KIND == SMK_FP8)
              umma_f8(tmem_d + (uint32_t)j * acc_stride, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                      (c > 0 || ks > 0) ? 1u : 0u);
            else
              umma_i8(tmem_d + (uint32_t)j * acc_stride, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc,
                      (c > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);          // frees stage s when these MMAs retire
        if (++s == NS) { s = 0; ++use; }
      }
      umma_commit(accum_full);
    }
  }

  const int S = p.split_k;
  if (warp >= 2) {
    // ===================== epilogue =====================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quad = warp & 3;                                   // TMEM lanes 32*quad .. +31 belong to this warp
    T* cptr = reinterpret_cast<T*>(p.c);
    const bool has_bias = p.bias != nullptr;
    uint32_t* slab = S > 1 ? p.scratch + (size_t)blockIdx.z * p.M * p.N : nullptr;
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      const int ch = n_base + j * SM_NT + quad * 32 + lane;
      const bool ch_ok = ch < p.N;
      float bs = 0.f, bv = 0.f;
      if (ch_ok && S == 1) {
        bs = p.b_scale_per_channel ? __ldg(p.b_scales + ch) : __ldg(p.b_scales);
        if (has_bias) bv = to_f32<T>(reinterpret_cast<const T*>(p.bias)[ch]);
      }
      for (int col0 = 0; col0 < n_mma; col0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)j * acc_stride + (uint32_t)col0, v);
        if (S == 1) {
          float as_l = 1.f;
          if (col0 + lane < toks) as_l = p.a_scale_per_token ? __ldg(p.a_scales + tok_base + col0 + lane) : __ldg(p.a_scales);
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float as = __shfl_sync(0xffffffffu, as_l, t);
            if (ch_ok && col0 + t < toks) {
              const float acc = KIND == SMK_FP8 ? __uint_as_float(v[t]) : (float)(int)v[t];
              const float tmp = bs * acc;
              const float o = has_bias ? fmaf(as, tmp, bv) : as * tmp;
              cptr[(size_t)(tok_base + col0 + t) * p.ldc + ch] = from_f32<T>(o);
            }
          }
        } else if (ch_ok) {
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (col0 + t < toks) slab[(size_t)(tok_base + col0 + t) * p.N + ch] = v[t];
        }
      }
    }
    tc_fence_before();
  }

  if (S > 1) {
    __threadfence();                   // slab stores -> visible to the cluster's other SMs
    __syncthreads();
    sm_cluster_sync();
    // Reduction: this split owns token rows z, z + S, ...; every warp of the CTA takes rows of its own (row loads of
    // different warps overlap — the first version walked the rows with all threads in lock step, one L2 round trip per
    // row: 98 us on 4096 x 6144), a lane owns 4 adjacent channels (16-byte slab loads, 8-byte output stores), slabs are
    // added in split order.
    T* cptr = reinterpret_cast<T*>(p.c);
    const bool has_bias = p.bias != nullptr;
    const size_t slab_elems = (size_t)p.M * p.N;
    constexpr int NWARPS = SM_THREADS / 32;
#pragma unroll
    for (int cb = 0; cb < CTA_CH; cb += 128) {
      const int ch = n_base + cb + lane * 4;
      if (ch >= p.N) continue;
      float bsv[4], bvv[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.b_scale_per_channel) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.b_scales + ch));
        bsv[0] = t.x; bsv[1] = t.y; bsv[2] = t.z; bsv[3] = t.w;
      } else {
        bsv[0] = bsv[1] = bsv[2] = bsv[3] = __ldg(p.b_scales);
      }
      if (has_bias) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const T*>(p.bias) + ch));
        const T* tb = reinterpret_cast<const T*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) bvv[i] = to_f32<T>(tb[i]);
      }
      for (int tok = (int)blockIdx.z + S * warp; tok < toks; tok += S * NWARPS) {
        const uint32_t* src = p.scratch + (size_t)(tok_base + tok) * p.N + ch;
        float acc[4];
        if constexpr (KIND == SMK_FP8) {
          acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
          for (int z0 = 0; z0 < S; z0 += 4) {                   // four slab loads in flight, added in split order
            uint4 v4[4];
#pragma unroll
            for (int z = 0; z < 4; ++z)
              v4[z] = (z0 + z < S) ? __ldcg(reinterpret_cast<const uint4*>(src + (size_t)(z0 + z) * slab_elems)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int z = 0; z < 4; ++z) {
              acc[0] += __uint_as_float(v4[z].x); acc[1] += __uint_as_float(v4[z].y);
              acc[2] += __uint_as_float(v4[z].z); acc[3] += __uint_as_float(v4[z].w);
            }
          }
        } else {
          int ia[4] = {0, 0, 0, 0};
          for (int z = 0; z < S; ++z) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)z * slab_elems));
            ia[0] += (int)v.x; ia[1] += (int)v.y; ia[2] += (int)v.z; ia[3] += (int)v.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = (float)ia[i];
        }
        const float as = p.a_scale_per_token ? __ldg(p.a_scales + tok_base + tok) : __ldg(p.a_scales);
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float tmp = bsv[i] * acc[i];
          o[i] = has_bias ? fmaf(as, tmp, bvv[i]) : as * tmp;
        }
        uint2 packed;
        packed.x = pack2<T>(o[0], o[1]);
        packed.y = pack2<T>(o[2], o[3]);
        *reinterpret_cast<uint2*>(cptr + (size_t)(tok_base + tok) * p.ldc + ch) = packed;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// 2-D tensor map over a K-contiguous 8-bit matrix [rows, K] with row stride `ld` bytes; box [128 k x box_rows]
static int encode_u8_map(CUtensorMap* tmap, const void* base, int64_t rows, int K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld};
  const cuuint32_t box[2] = {(cuuint32_t)SM_KC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return 0;
}

// channels per CTA: 256 (two sub-tiles share one activation tile) when the 128-channel tiling would need more than one
// wave of the SMs — there the kernel runs at the L2 -> SM fabric rate and the shared activation tile cuts that traffic
// by a third (4096 x 28672 at 256 tokens: 71.7 -> 61.9 us); with fewer tiles the k-split fills the GPU and wider tiles
// only mean more slabs to reduce (4096 x 4096: 23.6 us at 128 channels, 53.2 us at 256; profiles/r02_bench_f_rows.json).
// b200_scaled_mm_set_tile() overrides for A-B measurement (0 = auto, 1 = 128, 2 = 256).
static thread_local int g_smm_tile = 0;
static int scaled_mm_nsub(int M, int N) {
  if (g_smm_tile == 1) return 1;
  if (g_smm_tile == 2) return 2;
  const int tiles = ((N + SM_NT - 1) / SM_NT) * ((M + SM_TOK - 1) / SM_TOK);
  return tiles > num_sms() ? 2 : 1;
}

// how many clusters of `split` CTAs of this kernel (one CTA per SM: ~200 KB of shared memory) the device holds at once.
// A cluster lives inside one GPC, so the answer is sum over GPCs of floor(SMs / split), not SMs / split: 48 tiles x 3
// splits = 144 CTAs "fit" 148 SMs but not their GPCs, and the clusters left over ran as a second wave.
static int smm_max_active_clusters(int split) {
  static thread_local int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (split < 2 || split > 8) return 1 << 30;
  if (cache[split] == 0) {
    auto kern = scaled_mm_tc5_kernel<__nv_bfloat16, SMK_FP8, 1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_TOTAL);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1, 1, (unsigned)split);
    cfg.blockDim = dim3(SM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SM_SMEM_TOTAL - 1024;
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
      n = num_sms() / split;        // no device (plan queries on a CPU box) or no answer: the arithmetic bound
    }
    cache[split] = n;
  }
  return cache[split];
}

// k-splits of one tile: enough to fill one wave of the SMs (and of the GPCs' cluster slots), at least 4 chunks (512 k)
// each, at most 8 (portable cluster)
static int plan_scaled_mm_split(int M, int N, int K) {
  const int cta_ch = SM_NT * scaled_mm_nsub(M, N);
  const int tiles = ((N + cta_ch - 1) / cta_ch) * ((M + SM_TOK - 1) / SM_TOK);
  const int chunks = (K + SM_KC - 1) / SM_KC;
  int split = std::min(std::min(num_sms() / std::max(tiles, 1), chunks / 4), 8);
  if (split < 1) split = 1;
  while (split > 1 && tiles > smm_max_active_clusters(split)) --split;
  while (split > 1 && (split - 1) * ((chunks + split - 1) / split) >= chunks) --split;     // never an empty split
  return split;
}

template <typename T, int KIND, int NSUB>
static int launch_scaled_mm_n(const CUtensorMap& tw, const CUtensorMap& ta, ScaledMMParams& p, dim3 grid, cudaStream_t st) {
  auto kern = scaled_mm_tc5_kernel<T, KIND, NSUB>;
  static thread_local uint64_t attr_done = 0;
  int dev = 0;
  B200_CUDA_OK(cudaGetDevice(&dev));
  if (!(attr_done >> (dev & 63) & 1)) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_TOTAL));
    attr_done |= 1ull << (dev & 63);
  }
  const int stage_bytes = SM_W_BYTES * NSUB + p.act_bytes;
  p.stages = std::min(SM_MAX_STAGES, (SM_SMEM_TOTAL - SM_FIXED) / stage_bytes);
  B200_CHECK(p.stages >= 2, "scaled_mm: shared-memory plan leaves fewer than two pipeline stages");
  const size_t smem = (size_t)p.stages * stage_bytes + SM_FIXED;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(SM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)p.split_k;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tw, ta, p));
  return check_launch("scaled_mm_tc5_kernel");
}

template <typename T, int KIND>
static int launch_scaled_mm(const CUtensorMap& tw, const CUtensorMap& ta, ScaledMMParams& p, dim3 grid, int nsub,
                            cudaStream_t st) {
  return nsub == 2 ? launch_scaled_mm_n<T, KIND, 2>(tw, ta, p, grid, st) : launch_scaled_mm_n<T, KIND, 1>(tw, ta, p, grid, st);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_scaled_mm_set_tile(int tile) {
  const int prev = g_smm_tile;
  g_smm_tile = (tile == 1 || tile == 2) ? tile : 0;
  return prev;
}

extern "C" int b200_scaled_mm_plan(int size_m, int size_n, int size_k) {
  if (size_m <= 0 || size_n <= 0 || size_k <= 0) return 1;
  return plan_scaled_mm_split(size_m, size_n, size_k);
}

extern "C" int b200_cutlass_scaled_mm_supports_fp8(int cuda_device_capability) {
  // the reference answers for its CUTLASS kernels (scaled_mm_entry.cu:66-80: sm_89 with CUDA >= 12.4, sm_90+); this
  // library is the sm_100a implementation
  return cuda_device_capability >= 100 ? 1 : 0;
}

extern "C" int b200_cutlass_scaled_mm(void* out, const void* a, const void* b, const float* a_scales,
                                      const float* b_scales, const void* bias, int size_m, int size_n, int size_k,
                                      int64_t lda, int64_t ldb, int64_t ldc, int a_scales_numel, int b_scales_numel,
                                      int ab_dtype, int out_dtype, int split_k, void* workspace, void* stream) {
  B200_CHECK(ab_dtype == B200_AB_FP8_E4M3 || ab_dtype == B200_AB_INT8, "cutlass_scaled_mm: a and b must be float8_e4m3fn or int8");
  B200_CHECK(out_dtype == B200_F16 || out_dtype == B200_BF16, "cutlass_scaled_mm: out must be float16 or bfloat16");
  B200_CHECK(size_m >= 0 && size_n > 0 && size_k > 0, "cutlass_scaled_mm: invalid problem size");
  B200_CHECK(a_scales_numel == 1 || a_scales_numel == size_m, "a_scales must hold 1 or M values");
  B200_CHECK(b_scales_numel == 1 || b_scales_numel == size_n, "b_scales must hold 1 or N values");
  B200_CHECK(size_k % 16 == 0 && lda % 16 == 0 && ldb % 16 == 0, "K, a.stride(0) and b.stride(1) must be multiples of 16 (16-byte alignment)");
  B200_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0,
             "a and b must be 16-byte aligned");
  B200_CHECK(lda >= size_k && ldb >= size_k && ldc >= size_n, "leading dimensions smaller than the row length");
  if (size_m == 0) return 0;
  const int toks = std::min(size_m, SM_TOK);
  const int box_rows = std::max(16, (toks + 15) & ~15);
  const int nsub = scaled_mm_nsub(size_m, size_n);
  CUtensorMap tw, ta;
  if (int rc = encode_u8_map(&tw, b, size_n, size_k, ldb, SM_NT * nsub)) return rc;
  if (int rc = encode_u8_map(&ta, a, size_m, size_k, lda, box_rows)) return rc;
  ScaledMMParams p{};
  p.c = out; p.a_scales = a_scales; p.b_scales = b_scales; p.bias = bias;
  p.M = size_m; p.N = size_n; p.K = size_k; p.ldc = ldc;
  p.a_scale_per_token = (a_scales_numel > 1 || size_m == 1) ? 1 : 0;
  p.b_scale_per_channel = (b_scales_numel > 1 || size_n == 1) ? 1 : 0;
  if (a_scales_numel == 1) p.a_scale_per_token = 0;
  if (b_scales_numel == 1) p.b_scale_per_channel = 0;
  p.box_rows = box_rows;
  p.act_bytes = box_rows * 128;
  const int chunks = (size_k + SM_KC - 1) / SM_KC;
  if (split_k <= 0) split_k = plan_scaled_mm_split(size_m, size_n, size_k);
  split_k = std::max(1, std::min(std::min(split_k, 8), chunks));
  if (workspace == nullptr) split_k = 1;                    // no scratch: one CTA per tile walks the whole k range
  // the split reduction moves 4 channels per lane: 16-byte slab loads, 8-byte output stores
  if (size_n % 4 != 0 || ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 7) != 0 ||
      (reinterpret_cast<uintptr_t>(workspace) & 15) != 0 || (reinterpret_cast<uintptr_t>(b_scales) & 15) != 0 ||
      (bias != nullptr && (reinterpret_cast<uintptr_t>(bias) & 7) != 0))
    split_k = 1;
  while (split_k > 1 && (split_k - 1) * ((chunks + split_k - 1) / split_k) >= chunks) --split_k;
  p.split_k = split_k;
  p.chunks_per_split = (chunks + split_k - 1) / split_k;
  p.scratch = reinterpret_cast<uint32_t*>(workspace);
  B200_CHECK(split_k == 1 || (reinterpret_cast<uintptr_t>(workspace) & 3) == 0, "scaled_mm workspace must be 4-byte aligned");
  const int cta_ch = SM_NT * nsub;
  dim3 grid((size_n + cta_ch - 1) / cta_ch, (size_m + SM_TOK - 1) / SM_TOK, split_k);
  cudaStream_t st = (cudaStream_t)stream;
  if (ab_dtype == B200_AB_FP8_E4M3) {
    if (out_dtype == B200_BF16) return launch_scaled_mm<__nv_bfloat16, SMK_FP8>(tw, ta, p, grid, nsub, st);
    return launch_scaled_mm<__half, SMK_FP8>(tw, ta, p, grid, nsub, st);
  }
  if (out_dtype == B200_BF16) return launch_scaled_mm<__nv_bfloat16, SMK_INT8>(tw, ta, p, grid, nsub, st);
  return launch_scaled_mm<__half, SMK_INT8>(tw, ta, p, grid, nsub, st);
}
