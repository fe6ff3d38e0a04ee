// tp_fused.cu — the tensor-parallel exchange step of a decode layer as ONE kernel over NVSwitch, sm_100a.
//
// In the reference a row-parallel linear is followed by three launches: the custom all-reduce kernel
// (kernels/all_reduce/custom_all_reduce.cuh:183-249, call site aphrodite/distributed/parallel_state.py:353-379), then
// `fused_add_rms_norm` (kernels/layernorm_kernels.cu:204-286) on the replicated result. Every rank adds the same residual
// and normalises the same [T, H] rows. Here the exchange and the row work are one kernel and the rows are SHARDED:
//
//   start barrier  (every rank's partial sums are in its symmetric buffer)
//   rank r, for its T/N rows:   s   = multimem.ld_reduce(X)            NVSwitch sums the N partials in flight (fp32
//                                                                      accumulate, one 16-bit rounding = the reference
//                                                                      kernel's `downcast(sum fp32)`, .cuh:150-168)
//                               z   = T(s + residual); residual = z    (layernorm_kernels.cu:231-234 rounding point)
//                               out = T(T(z * rsqrt(mean z^2 + eps)) * w)
//                               multimem.st(H, out)                    NVSwitch replicates the row into every rank's H
//   end barrier    (all rows of H have landed everywhere; nobody still reads X)
//
// Three algorithms (argument `algo`), same barriers, same row ownership:
//   B200_TP_P2P        unicast everything: N peer loads summed in fp32 in rank order, N peer stores, N flag adds.
//   B200_TP_MC_STORE   (default on NVSwitch) unicast loads + fp32 rank-order sum — BIT-IDENTICAL to the reference kernel's
//                      arithmetic — and the switch used for what has no arithmetic in it: multimem.st replicates the
//                      result (1 store instead of N) and multimem.red signals all peers with one instruction.
//   B200_TP_MC_REDUCE  multimem.ld_reduce does the sum inside the switch (1 load instead of N). Measured on the B200
//                      box: the switch does NOT round the fp32 sum to nearest-even when it narrows to bf16/f16
//                      (tools/debug_tp_fused.py prints the census), so results differ from the reference's by up to an
//                      ulp per reduction, with a bias; opt-in only.
// With the row work laid out exactly like rms_norm_vec_kernel (256 threads, same vector-to-thread map, same reduction
// tree) the P2P and MC_STORE variants reproduce all-reduce + fused_add_rms_norm BIT-EXACTLY, not just within tolerance.
//
// so the residual stream is touched by ONE rank per row (rank r keeps rows [r*ceil(T/N), ...) of `residual`, the other
// rows of its buffer are never read), NVLink carries 2 x T*H*2 bytes per GPU per call (the algorithmic minimum of an
// all-reduce), and the two flag barriers are one `multimem.red` each (a single instruction signals all peers; every rank
// polls its OWN memory). NORM = false gives the plain two-shot all-reduce (reduce-scatter by ld_reduce + all-gather by
// multimem.st). MC = false is the same algorithm over unicast peer pointers (N loads / N stores per packet) for boxes
// without a multicast mapping; it sums in rank order so all ranks see bit-identical results, like the reference.
//
// Memory contract (host side: aphrodite_engine_b200/distributed/nvls.py): one symmetric allocation per rank, mapped at
// `peer_bases[r]` on every rank and (MC) bound to one multicast object mapped at `mc_base`; the kernel is given byte
// offsets of X, H and the flag area inside it. Flags are monotonically increasing u32 counters (wrap-safe compares), so
// there is nothing to reset between launches and the kernel is CUDA-graph capturable (no host-side epoch).
#include "common.cuh"

namespace b200 {

static constexpr int kTpThreads = 256;   // as rms_norm_vec_kernel: the row reduction must add in the same order
static constexpr int kTpMaxBlocks = 128;   // flag slots per direction
static constexpr int kTpMaxRanks = 8;

struct TpParams {
  char* mc_base;                     // multicast VA of the symmetric block (MC only)
  char* local_base;                  // this rank's unicast VA of the block
  char* peer_base[kTpMaxRanks];      // unicast VAs of every rank's block (peer_base[rank] == local_base)
  int64_t in_off, out_off, flag_off; // byte offsets inside the block
  void* residual;                    // local [T, H]; only this rank's rows are read / written (NORM only)
  const void* weight;                // [H] (NORM only)
  float eps;
  int num_tokens, hidden, rank, world;
  unsigned long long* stamps;        // optional [5] globaltimer stamps of CTA 0 (bring-up / profiling), else nullptr
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flag area: u32 counter[kTpMaxBlocks] (signalled by every rank, twice per launch) | u32 epoch[kTpMaxBlocks] (local)
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// RELEASE = false: the start barrier. Nothing written by this kernel has to be published (the partial sums were
// written by the previous kernel on the stream and are visible at kernel start), so the signal is a relaxed add.
// RELEASE = true: the end barrier. The signalling thread's release-add follows a CTA barrier, so by cumulativity it
// publishes every thread's multimem.st / peer stores of this CTA — no per-thread system fence (65k MEMBAR.SYS cost
// more than the whole exchange: measured 25 us -> see profiles/).
enum { TP_P2P = 0, TP_MC_STORE = 1, TP_MC_REDUCE = 2 };

template <int ALGO, bool RELEASE>
__device__ __forceinline__ void tp_signal(const TpParams& p, int slot) {
  const int64_t off = p.flag_off + (int64_t)slot * 4;
  if (ALGO != TP_P2P) {
    if (threadIdx.x == 0) {
      if (RELEASE) asm volatile("multimem.red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.mc_base + off) : "memory");
      else asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.mc_base + off) : "memory");
    }
  } else {
    if ((int)threadIdx.x < p.world) {
      if (RELEASE) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.peer_base[threadIdx.x] + off) : "memory");
      else asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.peer_base[threadIdx.x] + off) : "memory");
    }
  }
}
__device__ __forceinline__ void tp_wait(const TpParams& p, int slot, uint32_t target) {
  if (threadIdx.x == 0) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p.local_base + p.flag_off) + slot;
    while ((int32_t)(ld_relaxed_sys_u32(f) - target) < 0) {
    }
    asm volatile("fence.acquire.sys;" ::: "memory");     // one acquire after the spin instead of one per poll
  }
  __syncthreads();
}

template <typename T> struct Pk {   // 16 bytes of T
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T e[N];
  };
};

template <typename T> __device__ __forceinline__ uint4 mc_ld_reduce(const void* mc_addr);
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const void* a) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}
template <> __device__ __forceinline__ uint4 mc_ld_reduce<__half>(const void* a) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(void* mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w) : "memory");
}

// peer data was written by the peer's previous kernel and published by the start barrier: a plain (non-.nc) load that
// the compiler may not hoist or cache across the barrier
__device__ __forceinline__ uint4 ld_peer_v4(const void* a) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a) : "memory");
  return v;
}

// The VPT packets of one thread: every load of every packet is issued before the first add, so a row costs ONE
// NVLink round trip whatever VPT and the world size are.
template <typename T, int ALGO, int VPT>
__device__ __forceinline__ void tp_reduce_packets(const TpParams& p, int64_t row_off, int nvec, Pk<T> (&z)[VPT]) {
  if constexpr (ALGO == TP_MC_REDUCE) {
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kTpThreads;
      if (v < nvec) z[k].raw = mc_ld_reduce<T>(p.mc_base + p.in_off + row_off + (int64_t)v * 16);
    }
  } else {
    Pk<T> in[VPT][kTpMaxRanks];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kTpThreads;
#pragma unroll
      for (int r = 0; r < kTpMaxRanks; ++r)
        if (r < p.world && v < nvec) in[k][r].raw = ld_peer_v4(p.peer_base[r] + p.in_off + row_off + (int64_t)v * 16);
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      float acc[Pk<T>::N];
#pragma unroll
      for (int e = 0; e < Pk<T>::N; ++e) acc[e] = 0.f;
#pragma unroll
      for (int r = 0; r < kTpMaxRanks; ++r)          // rank order, fp32: the reference kernel's arithmetic
        if (r < p.world) {
#pragma unroll
          for (int e = 0; e < Pk<T>::N; ++e) acc[e] += to_f32<T>(in[k][r].e[e]);
        }
#pragma unroll
      for (int e = 0; e < Pk<T>::N; ++e) z[k].e[e] = from_f32<T>(acc[e]);
    }
  }
}
template <int ALGO>
__device__ __forceinline__ void tp_broadcast_packet(const TpParams& p, int64_t byte_off, uint4 v) {
  if (ALGO != TP_P2P) {
    mc_st(p.mc_base + p.out_off + byte_off, v);
  } else {
#pragma unroll
    for (int r = 0; r < kTpMaxRanks; ++r)
      if (r < p.world) {
        const int dst = (p.rank + r) % p.world;   // stagger the peers
        *reinterpret_cast<uint4*>(p.peer_base[dst] + p.out_off + byte_off) = v;
      }
  }
}

// grid = min(rows per rank, kTpMaxBlocks) CTAs (identical on every rank); CTA b owns rows b, b+grid, ... of the slice.
template <typename T, int ALGO, bool NORM, int VPT>
__global__ void __launch_bounds__(kTpThreads) tp_allreduce_rows_kernel(const TpParams p) {
  __shared__ float red[kTpThreads / 32];
  constexpr int N = Pk<T>::N;
  uint32_t* flags = reinterpret_cast<uint32_t*>(p.local_base + p.flag_off);
  const uint32_t epoch = flags[kTpMaxBlocks + blockIdx.x];           // launches this slot has seen (local counter)
  const uint32_t t_start = (2u * epoch + 1u) * (uint32_t)p.world;
  const uint32_t t_end = (2u * epoch + 2u) * (uint32_t)p.world;

  const int rows_per = (p.num_tokens + p.world - 1) / p.world;
  const int row0 = p.rank * rows_per;
  const int row1 = min(p.num_tokens, row0 + rows_per);
  const int nvec = p.hidden / N;
  const int64_t row_bytes = (int64_t)p.hidden * sizeof(T);

  const bool stamp = p.stamps != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) p.stamps[0] = globaltimer_ns();
  tp_signal<ALGO, false>(p, blockIdx.x);
  // the first row's residual does not depend on the peers: fetch it while the start barrier is in flight
  Pk<T> rpre[VPT];
  if (NORM && row0 + (int)blockIdx.x < row1) {
    const T* res = reinterpret_cast<const T*>(p.residual) + (int64_t)(row0 + blockIdx.x) * p.hidden;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kTpThreads;
      if (v < nvec) rpre[k].raw = *reinterpret_cast<const uint4*>(res + (int64_t)v * N);
    }
  }
  tp_wait(p, blockIdx.x, t_start);
  if (stamp) p.stamps[1] = globaltimer_ns();

  for (int row = row0 + blockIdx.x; row < row1; row += gridDim.x) {
    Pk<T> z[VPT];
    tp_reduce_packets<T, ALGO, VPT>(p, row * row_bytes, nvec, z);
    if (stamp && row == row0) p.stamps[2] = globaltimer_ns() + 0 * (unsigned long long)z[0].raw.x;   // after the loads landed
    if (NORM) {
      float ss = 0.f;
      T* res = reinterpret_cast<T*>(p.residual) + (int64_t)row * p.hidden;
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int v = threadIdx.x + k * kTpThreads;
        if (v < nvec) {
          Pk<T> r;
          if (row == row0 + (int)blockIdx.x) r.raw = rpre[k].raw;
          else r.raw = *reinterpret_cast<const uint4*>(res + (int64_t)v * N);
#pragma unroll
          for (int e = 0; e < N; ++e) {
            z[k].e[e] = from_f32<T>(__fadd_rn(to_f32<T>(z[k].e[e]), to_f32<T>(r.e[e])));
            const float f = to_f32<T>(z[k].e[e]);
            ss = fmaf(f, f, ss);
          }
          *reinterpret_cast<uint4*>(res + (int64_t)v * N) = z[k].raw;
        }
      }
      // the reduction tree of rms_norm_vec_kernel (block_sum_256): xor-shuffle warp sums, then the 8 warp sums in order
      ss = warp_sum(ss);
      __syncthreads();                       // `red` reuse across rows
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
      __syncthreads();
      float tot = 0.f;
      for (int i = 0; i < kTpThreads / 32; ++i) tot += red[i];
      const float s = rsqrtf(tot / (float)p.hidden + p.eps);
#pragma unroll
      for (int k = 0; k < VPT; ++k) {
        const int v = threadIdx.x + k * kTpThreads;
        if (v < nvec) {
          Pk<T> w;
          w.raw = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.weight) + (int64_t)v * N));
#pragma unroll
          for (int e = 0; e < N; ++e)
            z[k].e[e] = from_f32<T>(__fmul_rn(to_f32<T>(from_f32<T>(__fmul_rn(to_f32<T>(z[k].e[e]), s))),
                                             to_f32<T>(w.e[e])));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = threadIdx.x + k * kTpThreads;
      if (v < nvec) tp_broadcast_packet<ALGO>(p, row * row_bytes + (int64_t)v * 16, z[k].raw);
    }
  }
  __syncthreads();                 // every thread's stores happen-before the release-add of the signalling thread(s)
  if (stamp) p.stamps[3] = globaltimer_ns();
  tp_signal<ALGO, true>(p, blockIdx.x);
  tp_wait(p, blockIdx.x, t_end);
  if (stamp) p.stamps[4] = globaltimer_ns();
  if (threadIdx.x == 0) flags[kTpMaxBlocks + blockIdx.x] = epoch + 1u;
}

template <typename T, int ALGO, bool NORM>
static int launch_tp(const TpParams& p, cudaStream_t st) {
  const int nvec = p.hidden / Pk<T>::N;
  const int rows_per = (p.num_tokens + p.world - 1) / p.world;
  const int grid = std::max(1, std::min(rows_per, kTpMaxBlocks));
  if (nvec <= kTpThreads)
    tp_allreduce_rows_kernel<T, ALGO, NORM, 1><<<grid, kTpThreads, 0, st>>>(p);
  else if (nvec <= 2 * kTpThreads)
    tp_allreduce_rows_kernel<T, ALGO, NORM, 2><<<grid, kTpThreads, 0, st>>>(p);
  else
    tp_allreduce_rows_kernel<T, ALGO, NORM, 4><<<grid, kTpThreads, 0, st>>>(p);
  return check_launch("tp_allreduce_rows_kernel");
}

}  // namespace b200

using namespace b200;

static thread_local unsigned long long* g_tp_stamps = nullptr;
// profiling hook: device buffer of 5 x u64 that CTA 0 of the following launches fills with %globaltimer stamps
// (kernel start, start barrier passed, first row's loads landed, stores issued, end barrier passed); nullptr = off
extern "C" void b200_tp_set_stamp_buffer(void* dev_u64x5) { g_tp_stamps = static_cast<unsigned long long*>(dev_u64x5); }

extern "C" int64_t b200_tp_flag_bytes(void) { return (int64_t)(2 * kTpMaxBlocks * sizeof(uint32_t)); }

extern "C" int b200_tp_allreduce_rows(void* mc_base, void* local_base, const int64_t* peer_bases, int64_t in_off,
                                      int64_t out_off, int64_t flag_off, void* residual, const void* weight,
                                      float epsilon, int num_tokens, int hidden, int rank, int world, int dtype,
                                      int algo, void* stream) {
  B200_CHECK(dtype == B200_F16 || dtype == B200_BF16, "tp_allreduce_rows: float16 / bfloat16 only");
  B200_CHECK(world >= 2 && world <= kTpMaxRanks && rank >= 0 && rank < world, "tp_allreduce_rows: bad rank / world");
  B200_CHECK(algo >= TP_P2P && algo <= TP_MC_REDUCE, "tp_allreduce_rows: algo must be 0 (p2p), 1 (multicast store) or 2 (multicast reduce)");
  B200_CHECK(local_base != nullptr, "tp_allreduce_rows: no symmetric block");
  B200_CHECK(algo == TP_P2P || mc_base != nullptr, "tp_allreduce_rows: this algorithm needs the multicast mapping");
  B200_CHECK(algo == TP_MC_REDUCE || peer_bases != nullptr, "tp_allreduce_rows: this algorithm needs the peer table");
  B200_CHECK(hidden % 8 == 0 && hidden / 8 <= 4 * kTpThreads, "tp_allreduce_rows: hidden must be a multiple of 8, <= 8192");
  B200_CHECK(((in_off | out_off) & 15) == 0 && (flag_off & 127) == 0, "tp_allreduce_rows: misaligned offsets");
  B200_CHECK((residual == nullptr) == (weight == nullptr), "tp_allreduce_rows: residual and weight go together");
  B200_CHECK(((reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0,
             "tp_allreduce_rows: residual / weight must be 16-byte aligned");
  if (num_tokens == 0) return 0;
  TpParams p;
  p.mc_base = static_cast<char*>(mc_base);
  p.local_base = static_cast<char*>(local_base);
  for (int r = 0; r < kTpMaxRanks; ++r)
    p.peer_base[r] = (peer_bases != nullptr && r < world) ? reinterpret_cast<char*>(peer_bases[r]) : nullptr;
  if (peer_bases != nullptr)
    B200_CHECK(p.peer_base[rank] == p.local_base, "tp_allreduce_rows: peer_bases[rank] must be the local mapping");
  p.in_off = in_off; p.out_off = out_off; p.flag_off = flag_off;
  p.residual = residual; p.weight = weight; p.eps = epsilon;
  p.num_tokens = num_tokens; p.hidden = hidden; p.rank = rank; p.world = world;
  p.stamps = g_tp_stamps;
  cudaStream_t st = (cudaStream_t)stream;
  const bool norm = residual != nullptr;
#define B200_TP_A(T, A) (norm ? launch_tp<T, A, true>(p, st) : launch_tp<T, A, false>(p, st))
#define B200_TP(T) (algo == TP_P2P ? B200_TP_A(T, TP_P2P) : (algo == TP_MC_STORE ? B200_TP_A(T, TP_MC_STORE) : B200_TP_A(T, TP_MC_REDUCE)))
  return dtype == B200_BF16 ? B200_TP(__nv_bfloat16) : B200_TP(__half);
#undef B200_TP
#undef B200_TP_A
}
