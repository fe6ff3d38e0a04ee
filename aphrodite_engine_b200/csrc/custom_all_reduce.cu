// custom_all_reduce.cu — tensor-parallel sum all-reduce over NVLink peer memory, sm_100a.
//
// Replaces the reference's custom all-reduce (kernels/all_reduce/custom_all_reduce.cuh:29-484, entry points
// kernels/all_reduce/custom_all_reduce.cu:15-141; schemas kernels/torch_bindings.cpp:510-535). Same contract:
//   * every rank owns a `meta` allocation = [Signal | scratch of max_size bytes] and a registered input buffer;
//     peers exchange cudaIpcMemHandle's (the Python side gathers them) and map each other's allocations;
//   * inputs must be IPC-registered (eager: the caller copies into the registered buffer; CUDA graphs: addresses
//     seen during capture are collected and registered afterwards — get_graph_buffer_ipc_meta /
//     register_graph_buffers — which is why the kernels read their peer pointers from a DEVICE-side table);
//   * fp32 accumulation in rank order 0..n-1 on every rank, so all ranks produce bit-identical sums.
// B200 design: NVSwitch gives every peer the same 900 GB/s, so there is no ring — one-shot (every rank reads all
// peers; world 2 or small messages) or two-shot (reduce-scatter into the owner's scratch + all-gather) with 16-byte
// peer loads, enough CTAs in flight to cover the ~1 us NVLink load latency (the reference caps at 36 blocks, tuned
// on A100 PCIe/NVLink3), and flag barriers in peer memory with system-scope release/acquire.
#include "common.cuh"

#include <cuda.h>

#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace b200 {

static constexpr int kMaxRanks = 8;
static constexpr int kMaxBlocks = 128;
static constexpr int kArThreads = 512;

// one per rank, at the start of its `meta` allocation; peers write into it
struct Signal {
  alignas(128) uint32_t start[kMaxBlocks][kMaxRanks];
  alignas(128) uint32_t end[kMaxBlocks][kMaxRanks];
  alignas(128) uint32_t epoch[kMaxBlocks];  // this rank's launch counter per block (local use only)
};

struct RankData {
  const void* ptrs[kMaxRanks];
};
struct RankSignals {
  Signal* signals[kMaxRanks];
};

__device__ __forceinline__ void st_flag_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag_volatile(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All ranks' block `blockIdx.x` meet. START: nothing before it needs to be published (inputs were written by
// earlier stream work, visible at kernel start). Thread r tells peer r "I am here", then waits for peer r.
template <int NGPUS, bool END, bool NEED_FENCE>
__device__ __forceinline__ void barrier(const RankSignals& sg, Signal* self, int rank, uint32_t flag) {
  if (END) {                  // every thread of this block has finished its reads / writes
    if (NEED_FENCE) __threadfence_system();   // ... and its scratch writes are visible to the peers
    __syncthreads();
  }
  if (threadIdx.x < NGPUS) {
    uint32_t* peer = END ? &sg.signals[threadIdx.x]->end[blockIdx.x][rank] : &sg.signals[threadIdx.x]->start[blockIdx.x][rank];
    const uint32_t* mine = END ? &self->end[blockIdx.x][threadIdx.x] : &self->start[blockIdx.x][threadIdx.x];
    if (NEED_FENCE) {
      st_flag_release(peer, flag);
      while (ld_flag_acquire(mine) != flag) {}
    } else {
      st_flag_volatile(peer, flag);
      while (ld_flag_volatile(mine) != flag) {}
    }
  }
  if (!END || NEED_FENCE) __syncthreads();
}

template <typename T> struct Packed {  // 16 bytes of T
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T e[N];
  };
};

template <typename T, int NGPUS>
__device__ __forceinline__ uint4 reduce_packets(const RankData& rd, int64_t idx) {
  float acc[Packed<T>::N];
#pragma unroll
  for (int e = 0; e < Packed<T>::N; ++e) acc[e] = 0.f;
  Packed<T> v[NGPUS];
#pragma unroll
  for (int r = 0; r < NGPUS; ++r)  // all loads first: NGPUS independent 16-byte peer loads in flight
    v[r].raw = __ldg(reinterpret_cast<const uint4*>(rd.ptrs[r]) + idx);   // inputs are read-only during the kernel
#pragma unroll
  for (int r = 0; r < NGPUS; ++r)
#pragma unroll
    for (int e = 0; e < Packed<T>::N; ++e) acc[e] += to_f32<T>(v[r].e[e]);
  Packed<T> o;
#pragma unroll
  for (int e = 0; e < Packed<T>::N; ++e) o.e[e] = from_f32<T>(acc[e]);
  return o.raw;
}

// one-shot: every rank reduces the whole message from all peers
template <typename T, int NGPUS>
__global__ void __launch_bounds__(kArThreads, 1)
all_reduce_1shot_kernel(const RankData* __restrict__ rdp, RankSignals sg, Signal* self, T* __restrict__ out,
                        int rank, int64_t packets) {
  const RankData rd = *rdp;
  uint32_t flag = 0;
  if (threadIdx.x < NGPUS) flag = self->epoch[blockIdx.x] + 1;
  barrier<NGPUS, false, false>(sg, self, rank, flag);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < packets; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(out)[i] = reduce_packets<T, NGPUS>(rd, i);
  barrier<NGPUS, true, false>(sg, self, rank, flag);   // peers may not overwrite their inputs before I have read them
  if (threadIdx.x == 0) self->epoch[blockIdx.x] = flag;
}

// two-shot: rank r reduces slice r into its scratch (after its Signal), then everyone gathers all slices
template <typename T, int NGPUS>
__global__ void __launch_bounds__(kArThreads, 1)
all_reduce_2shot_kernel(const RankData* __restrict__ rdp, RankSignals sg, Signal* self, T* __restrict__ out,
                        int rank, int64_t packets) {
  const RankData rd = *rdp;
  const int64_t part = packets / NGPUS;
  const int64_t start = rank * part;
  const int64_t mine = (rank == NGPUS - 1) ? packets - start : part;
  uint32_t flag = 0;
  if (threadIdx.x < NGPUS) flag = self->epoch[blockIdx.x] + 1;
  uint4* scratch[NGPUS];
#pragma unroll
  for (int r = 0; r < NGPUS; ++r) scratch[r] = reinterpret_cast<uint4*>(sg.signals[r] + 1);
  barrier<NGPUS, false, false>(sg, self, rank, flag);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < mine; i += stride) scratch[rank][i] = reduce_packets<T, NGPUS>(rd, start + i);
  // slice results must be visible to the peers that gather them: release/acquire barrier
  barrier<NGPUS, true, true>(sg, self, rank, flag);
  // gather: thread (block b, thread t) reads exactly what peer's (block b, thread t) wrote, so the per-block
  // barrier above is sufficient
#pragma unroll
  for (int r = 0; r < NGPUS; ++r) {
    const int src = (rank + r) % NGPUS;              // stagger peers
    const int64_t sstart = src * part;
    const int64_t scount = (src == NGPUS - 1) ? packets - sstart : part;
    for (int64_t i = tid; i < scount; i += stride)
      reinterpret_cast<uint4*>(out)[sstart + i] = scratch[src][i];
  }
  if (threadIdx.x == 0) self->epoch[blockIdx.x] = flag;
  // the next launch's start barrier uses a new flag value, and nobody can pass it before every rank has
  // finished this kernel (same stream order on every rank), so scratch reuse is safe
}

// ------------------------------------------------------------------------------------------------------------
using IpcKey = std::array<char, sizeof(cudaIpcMemHandle_t)>;

class CustomAllreduce {
 public:
  int rank_, world_size_;
  bool full_nvlink_;
  RankSignals sg_;
  Signal* self_sg_;
  RankData* d_rank_data_base_;
  RankData* d_rank_data_end_;
  std::unordered_map<const void*, RankData*> buffers_;   // registered input pointer -> device table entry
  std::vector<void*> graph_unreg_buffers_;               // addresses seen during capture, registered afterwards
  std::map<IpcKey, char*> ipc_cache_;

  CustomAllreduce(Signal* meta, void* rank_data, size_t rank_data_sz, const cudaIpcMemHandle_t* handles,
                  const int64_t* offsets, int world_size, int rank, bool full_nvlink)
      : rank_(rank), world_size_(world_size), full_nvlink_(full_nvlink), self_sg_(meta),
        d_rank_data_base_(reinterpret_cast<RankData*>(rank_data)),
        d_rank_data_end_(reinterpret_cast<RankData*>(rank_data) + rank_data_sz / sizeof(RankData)) {
    for (int i = 0; i < world_size_; ++i) {
      if (i == rank_) {
        sg_.signals[i] = meta;
      } else {
        char* base = open_ipc(handles[i]);
        sg_.signals[i] = reinterpret_cast<Signal*>(base + offsets[i]);
      }
    }
  }

  char* open_ipc(const cudaIpcMemHandle_t& h) {
    IpcKey key;
    memcpy(key.data(), &h, sizeof(h));
    auto it = ipc_cache_.find(key);
    if (it != ipc_cache_.end()) return it->second;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    ipc_cache_[key] = static_cast<char*>(p);
    return static_cast<char*>(p);
  }

  RankData* alloc_entry() {
    if (d_rank_data_base_ + 1 > d_rank_data_end_) throw std::runtime_error("rank_data buffer is overflowed");
    return d_rank_data_base_++;
  }

  void register_buffer(const cudaIpcMemHandle_t* handles, const int64_t* offsets, void* self) {
    RankData data;
    for (int i = 0; i < world_size_; ++i)
      data.ptrs[i] = (i == rank_) ? self : open_ipc(handles[i]) + offsets[i];
    RankData* d = alloc_entry();
    cudaError_t e = cudaMemcpy(d, &data, sizeof(RankData), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) throw std::runtime_error(std::string("register_buffer: ") + cudaGetErrorString(e));
    buffers_[self] = d;
  }

  ~CustomAllreduce() {
    for (auto& kv : ipc_cache_) cudaIpcCloseMemHandle(kv.second);
  }
};

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetRangeFn get_range_fn() {
  static GetRangeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GetRangeFn>(p);
  }
  return fn;
}

template <typename T>
static int launch_ar(CustomAllreduce* fa, const RankData* rd, T* out, int64_t numel, cudaStream_t st) {
  const int64_t packets = numel / Packed<T>::N;
  const int64_t bytes = numel * (int64_t)sizeof(T);
  // one-shot when there are two ranks or the message is small (latency-bound); two-shot otherwise
  const bool one_shot = fa->world_size_ == 2 || bytes <= 256 * 1024;
  int64_t want = (packets / (one_shot ? 1 : fa->world_size_) + kArThreads - 1) / kArThreads;
  int blocks = (int)std::min<int64_t>(std::max<int64_t>(want, 1), kMaxBlocks);
  blocks = std::min(blocks, num_sms());
#define B200_AR(NG)                                                                                          \
  if (one_shot)                                                                                              \
    all_reduce_1shot_kernel<T, NG><<<blocks, kArThreads, 0, st>>>(rd, fa->sg_, fa->self_sg_, out, fa->rank_, packets); \
  else                                                                                                       \
    all_reduce_2shot_kernel<T, NG><<<blocks, kArThreads, 0, st>>>(rd, fa->sg_, fa->self_sg_, out, fa->rank_, packets);
  switch (fa->world_size_) {
    case 2: B200_AR(2) break;
    case 4: B200_AR(4) break;
    case 6: B200_AR(6) break;
    case 8: B200_AR(8) break;
    default: return fail("custom allreduce only supports num gpus in (2,4,6,8)");
  }
#undef B200_AR
  return check_launch("all_reduce kernel");
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_car_meta_size(void) { return (int64_t)sizeof(Signal); }

extern "C" int64_t b200_car_init(void* meta, void* rank_data, int64_t rank_data_bytes, const void* handles,
                                 const int64_t* offsets, int world_size, int rank, int full_nvlink) {
  if (world_size > kMaxRanks || world_size % 2 != 0 || rank < 0 || rank >= world_size) {
    fail(world_size > kMaxRanks ? "world size > 8 is not supported"
                                : (world_size % 2 ? "Odd num gpus is not supported for now" : "invalid rank passed in"));
    return 0;
  }
  try {
    return (int64_t) new CustomAllreduce(reinterpret_cast<Signal*>(meta), rank_data, (size_t)rank_data_bytes,
                                         reinterpret_cast<const cudaIpcMemHandle_t*>(handles), offsets, world_size,
                                         rank, full_nvlink != 0);
  } catch (const std::exception& e) {
    fail(e.what());
    return 0;
  }
}

extern "C" void b200_car_dispose(int64_t fa) { delete reinterpret_cast<CustomAllreduce*>(fa); }

extern "C" int b200_car_register_buffer(int64_t fa_, void* self_ptr, const void* handles, const int64_t* offsets) {
  auto fa = reinterpret_cast<CustomAllreduce*>(fa_);
  try {
    fa->register_buffer(reinterpret_cast<const cudaIpcMemHandle_t*>(handles), offsets, self_ptr);
  } catch (const std::exception& e) {
    return fail(e.what());
  }
  return 0;
}

// all-reduce of a registered (or, under stream capture, to-be-registered) input. numel elements of dtype.
extern "C" int b200_car_all_reduce(int64_t fa_, const void* inp, void* out, int64_t numel, int dtype, void* stream) {
  auto fa = reinterpret_cast<CustomAllreduce*>(fa_);
  cudaStream_t st = (cudaStream_t)stream;
  const int esz = dtype == B200_F32 ? 4 : 2;
  B200_CHECK(dtype >= B200_F32 && dtype <= B200_BF16, "custom allreduce only supports float32, float16 and bfloat16");
  B200_CHECK((numel * esz) % 16 == 0, "custom allreduce currently requires input length to be multiple of 16 bytes");
  B200_CHECK((reinterpret_cast<uintptr_t>(inp) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
             "custom allreduce requires 16-byte aligned buffers");
  const RankData* rd = nullptr;
  cudaStreamCaptureStatus status;
  B200_CUDA_OK(cudaStreamIsCapturing(st, &status));
  if (status == cudaStreamCaptureStatusActive) {
    // the table entry is filled by register_graph_buffers() after capture (reference custom_all_reduce.cuh:417-430)
    rd = fa->d_rank_data_base_ + fa->graph_unreg_buffers_.size();
    B200_CHECK(rd + 1 <= fa->d_rank_data_end_, "rank_data buffer is overflowed");
    fa->graph_unreg_buffers_.push_back(const_cast<void*>(inp));
  } else {
    auto it = fa->buffers_.find(inp);
    B200_CHECK(it != fa->buffers_.end(),
               "buffer address is not registered! (call register_buffer, or capture inside the graph-buffer context)");
    rd = it->second;
  }
  if (dtype == B200_BF16) return launch_ar<__nv_bfloat16>(fa, rd, (__nv_bfloat16*)out, numel, st);
  if (dtype == B200_F16) return launch_ar<__half>(fa, rd, (__half*)out, numel, st);
  return launch_ar<float>(fa, rd, (float*)out, numel, st);
}

// After capture: IPC handles (64 B each, concatenated into `handles_out`, capacity in buffers) and offsets of every
// address recorded during capture. Returns the number of buffers, or -1 on error.
extern "C" int b200_car_get_graph_buffer_ipc_meta(int64_t fa_, void* handles_out, int64_t* offsets_out, int capacity) {
  auto fa = reinterpret_cast<CustomAllreduce*>(fa_);
  const int n = (int)fa->graph_unreg_buffers_.size();
  if (handles_out == nullptr) return n;
  if (n > capacity) { fail("get_graph_buffer_ipc_meta: capacity too small"); return -1; }
  GetRangeFn range = get_range_fn();
  if (!range) { fail("cuMemGetAddressRange is not available"); return -1; }
  for (int i = 0; i < n; ++i) {
    void* ptr = fa->graph_unreg_buffers_[i];
    CUdeviceptr base = 0;
    size_t sz = 0;
    if (range(&base, &sz, (CUdeviceptr)ptr) != CUDA_SUCCESS) { fail("failed to get pointer attr"); return -1; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, (void*)base) != cudaSuccess) { fail("cudaIpcGetMemHandle failed"); return -1; }
    memcpy(static_cast<char*>(handles_out) + (size_t)i * sizeof(h), &h, sizeof(h));
    offsets_out[i] = (int64_t)((char*)ptr - (char*)base);
  }
  return n;
}

// handles: world_size blobs of n*64 bytes (rank-major: handles[r*n + i]); offsets: world_size * n
extern "C" int b200_car_register_graph_buffers(int64_t fa_, const void* handles, const int64_t* offsets, int n) {
  auto fa = reinterpret_cast<CustomAllreduce*>(fa_);
  B200_CHECK(n == (int)fa->graph_unreg_buffers_.size(), "register_graph_buffers: buffer count mismatch");
  if (n == 0) return 0;
  std::vector<RankData> data(n);
  const cudaIpcMemHandle_t* hs = reinterpret_cast<const cudaIpcMemHandle_t*>(handles);
  try {
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < fa->world_size_; ++r)
        data[i].ptrs[r] = (r == fa->rank_) ? fa->graph_unreg_buffers_[i]
                                           : fa->open_ipc(hs[(size_t)r * n + i]) + offsets[(size_t)r * n + i];
  } catch (const std::exception& e) {
    return fail(e.what());
  }
  B200_CUDA_OK(cudaMemcpy(fa->d_rank_data_base_, data.data(), sizeof(RankData) * n, cudaMemcpyHostToDevice));
  fa->d_rank_data_base_ += n;
  fa->graph_unreg_buffers_.clear();
  return 0;
}
