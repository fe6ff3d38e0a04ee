// marlin_repack.cu — checkpoint weight layouts -> Marlin tile layout (integer re-tiling, bit-exact).
//
// Replaces gptq_marlin_repack (kernels/quantization/gptq_marlin/gptq_marlin_repack.cu:271-343) and
// awq_marlin_repack (.../awq_marlin_repack.cu:208-268). The target layout is defined by the reference's
// Python `marlin_weights` / `get_weight_perm` (aphrodite/quantization/utils/marlin_utils_test.py:30-92):
// the [K,N] matrix is cut into 16x64 (k x n) blocks; inside a block, "thread" i (0..31) owns, for each of
// the four 16x16 sub-tiles j, rows {2(i%4), +1, +8, +9} x columns {i/4, i/4+8} (its mma.sync B fragment);
// those 8 values are nibble-interleaved [0,2,4,6,1,3,5,7] into one int32 (two int32, [0,2,1,3], for 8-bit).
// Load-time only, so the kernel is output-oriented: one thread builds one output word by gathering its
// 8 (4) source values — no shared memory, fully coalesced stores.
#include "common.cuh"

namespace b200 {

template <int BITS, bool AWQ>
__global__ void __launch_bounds__(256)
marlin_repack_kernel(const uint32_t* __restrict__ src, const int32_t* __restrict__ perm,
                     uint32_t* __restrict__ out, int K, int N) {
  constexpr int PF = 32 / BITS;
  constexpr uint32_t MASK = (1u << BITS) - 1u;
  constexpr int WPB = 1024 / PF;  // words per 16x64 block
  const int words_per_row = N * 16 / PF;
  const int64_t total = (int64_t)(K / 16) * words_per_row;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int kt = (int)(w / words_per_row);
    const int c = (int)(w % words_per_row);
    const int nb = c / WPB, wi = c % WPB;
    const int grp = (BITS == 4) ? wi : (wi >> 1);  // (i, j) group: i = grp / 4, j = grp % 4
    const int wsub = (BITS == 4) ? 0 : (wi & 1);
    const int i = grp >> 2, j = grp & 3, m = i & 3;
    uint32_t word = 0;
#pragma unroll
    for (int e = 0; e < PF; ++e) {
      int t;
      if (BITS == 4) t = ((e & 3) << 1) | (e >> 2);               // [0,2,4,6,1,3,5,7][e]
      else t = wsub * 4 + (((e & 1) << 1) | (e >> 1));            // [0,2,1,3][e]
      const int tt = t & 3;
      const int r = 2 * m + (tt & 1) + 8 * (tt >> 1);
      const int col = (i >> 2) + 8 * (t >> 2);
      int k = kt * 16 + r;
      const int n = nb * 64 + j * 16 + col;
      uint32_t v;
      if (AWQ) {
        // AWQ [K, N/PF], column-interleaved: original column m' of a PF-group sits at nibble inv[m']
        const int g = n / PF, mm = n % PF;
        const int pos = (BITS == 4) ? (((mm & 1) << 2) | (mm >> 1)) : (((mm & 1) << 1) | (mm >> 1));
        v = (src[(int64_t)k * (N / PF) + g] >> (BITS * pos)) & MASK;
      } else {
        if (perm != nullptr) k = perm[k];
        v = (src[(int64_t)(k / PF) * N + n] >> (BITS * (k % PF))) & MASK;
      }
      word |= v << (BITS * e);
    }
    out[w] = word;
  }
}

template <bool AWQ>
static int launch_repack(const uint32_t* src, const int32_t* perm, uint32_t* out, int K, int N,
                         int bits, cudaStream_t st) {
  B200_CHECK(bits == 4 || bits == 8, "num_bits must be 4 or 8. Got = " + std::to_string(bits));
  B200_CHECK(K % 16 == 0, "size_k = " + std::to_string(K) + " is not divisible by tile_k_size = 16");
  B200_CHECK(N % 64 == 0, "size_n = " + std::to_string(N) + " is not divisible by tile_n_size = 64");
  if (K == 0 || N == 0) return 0;
  const int64_t total = (int64_t)(K / 16) * (N * 16 / (32 / bits));
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (bits == 4)
    marlin_repack_kernel<4, AWQ><<<(int)blocks, 256, 0, st>>>(src, perm, out, K, N);
  else
    marlin_repack_kernel<8, AWQ><<<(int)blocks, 256, 0, st>>>(src, perm, out, K, N);
  return check_launch("marlin_repack_kernel");
}

}  // namespace b200

extern "C" int b200_gptq_marlin_repack(const void* b_q_weight, const int32_t* perm, void* out,
                                       int size_k, int size_n, int num_bits, void* stream) {
  return b200::launch_repack<false>((const uint32_t*)b_q_weight, perm, (uint32_t*)out, size_k,
                                    size_n, num_bits, (cudaStream_t)stream);
}

extern "C" int b200_awq_marlin_repack(const void* b_q_weight, void* out, int size_k, int size_n,
                                      int num_bits, void* stream) {
  return b200::launch_repack<true>((const uint32_t*)b_q_weight, nullptr, (uint32_t*)out, size_k,
                                   size_n, num_bits, (cudaStream_t)stream);
}
