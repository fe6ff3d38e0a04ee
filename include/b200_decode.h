/*
 * b200_decode.h — C ABI of libb200decode.so: the sm_100a decode hot path that drops in behind
 * aphrodite-engine's custom-op boundary (aphrodite/_custom_ops.py -> torch.ops._C.* registered by
 * kernels/torch_bindings.cpp in the reference).
 *
 * Every entry point takes plain device pointers, sizes/strides in ELEMENTS (as the reference
 * launchers compute them from tensor strides) and a `cudaStream_t` passed as `void*`.
 * No torch types cross this boundary; the torch-op shim (csrc/torch_shim.cpp) and the ctypes
 * binding (aphrodite_engine_b200/_native.py) are both thin marshalling layers over it.
 *
 * Return value: 0 on success, non-zero on error (unsupported configuration, bad alignment, CUDA
 * launch failure). b200_last_error() returns a thread-local, human-readable description — the
 * shim turns it into the same RuntimeError the reference raises through TORCH_CHECK.
 * All launches are asynchronous on `stream`, allocate nothing and never synchronise, so every op
 * is CUDA-graph capturable (reference: decode runs under torch.cuda.graph,
 * aphrodite/worker/model_runner.py:1682+).
 *
 * Each declaration cites the reference interface it replaces (path:line under the reference tree).
 */
#ifndef B200_DECODE_H_
#define B200_DECODE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation / query dtype codes (torch scalar types Float / Half / BFloat16) */
enum { B200_F32 = 0, B200_F16 = 1, B200_BF16 = 2 };
/* kv-cache dtype codes: "auto" | "fp8"/"fp8_e4m3" | "fp8_e5m2"
 * (kernels/quantization/fp8/nvidia/quant_utils.cuh:531-569 DISPATCH_BY_KV_CACHE_DTYPE) */
enum { B200_KV_AUTO = 0, B200_KV_FP8_E4M3 = 1, B200_KV_FP8_E5M2 = 2 };

const char* b200_last_error(void);
/* library/ABI version, bumped when a signature changes */
int b200_abi_version(void);
/* parses the reference's kv_cache_dtype string; returns -1 for an unsupported string */
int b200_parse_kv_cache_dtype(const char* s);

/* ---- paged attention ------------------------------------------------------------------------
 * replaces paged_attention_v1  kernels/attention/attention_kernels.cu:809-830 (schema
 *          kernels/torch_bindings.cpp:25-35, prototype kernels/ops.h:6-16)
 *          paged_attention_v2  kernels/attention/attention_kernels.cu:974-998 (schema :38-49)
 * out        [num_seqs, num_heads, head_size]            (contiguous)
 * query      [num_seqs, num_heads, head_size]            row stride q_stride (may be a qkv view)
 * key_cache  [num_blocks, num_kv_heads, head_size/x, block_size, x]   x = 16 / sizeof(cache elt)
 * value_cache[num_blocks, num_kv_heads, head_size, block_size]
 * block_tables int32 [num_seqs, max_num_blocks_per_seq]; seq_lens int32 [num_seqs]
 * alibi_slopes float32 [num_heads] or NULL
 * v2 only: exp_sums,max_logits float32 [num_seqs,num_heads,max_num_partitions],
 *          tmp_out [num_seqs,num_heads,max_num_partitions,head_size]; partition = 512 tokens
 *          (aphrodite/attention/ops/paged_attn.py:13)
 * blocksparse_* : vert_stride <= 1 disables block-sparse attention (attention_kernels.cu:773-789)
 */
int b200_paged_attention_v1(
    void* out, const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, int block_size,
    float scale, const int32_t* block_tables, const int32_t* seq_lens,
    int max_num_blocks_per_seq, int max_seq_len, const float* alibi_slopes,
    int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
    int dtype, int kv_dtype, float k_scale, float v_scale,
    int tp_rank, int blocksparse_local_blocks, int blocksparse_vert_stride,
    int blocksparse_block_size, int blocksparse_head_sliding_step, void* stream);

int b200_paged_attention_v2(
    void* out, float* exp_sums, float* max_logits, void* tmp_out,
    const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, int block_size,
    float scale, const int32_t* block_tables, const int32_t* seq_lens,
    int max_num_blocks_per_seq, int max_seq_len, int max_num_partitions,
    const float* alibi_slopes,
    int64_t q_stride, int64_t kv_block_stride, int64_t kv_head_stride,
    int dtype, int kv_dtype, float k_scale, float v_scale,
    int tp_rank, int blocksparse_local_blocks, int blocksparse_vert_stride,
    int blocksparse_block_size, int blocksparse_head_sliding_step, void* stream);

/* Selects the implementation for the two calls above (tests / A-B measurement):
 *   0 = auto (tensor-core bulk-copy kernel when the shape allows, generic SIMT kernel otherwise; v1 launches split
 *       every sequence over a thread-block cluster when that fills the GPU's waves visibly better)
 *   1 = force the generic SIMT kernel;  2 / 4 = force a cluster split of 2 / 4 CTAs per sequence (v1).
 * Returns the previous value. */
int b200_set_attention_impl(int impl);
/* 1 if the last paged_attention call on this thread took the tensor-core path, else 0 */
int b200_last_attention_path(void);
/* CTAs per sequence (cluster size) of the last tensor-core v1 launch on this thread */
int b200_last_attention_cluster_split(void);

/* ---- cache ops ------------------------------------------------------------------------------
 * replaces reshape_and_cache       kernels/cache_kernels.cu:263-289 (schema torch_bindings.cpp:467-473)
 *          reshape_and_cache_flash kernels/cache_kernels.cu:304-330 (schema :476-484)
 *          copy_blocks             kernels/cache_kernels.cu:101-148 (schema :461-464)
 *          swap_blocks             kernels/cache_kernels.cu:24-63   (schema :456-458)
 *          convert_fp8             kernels/cache_kernels.cu:356-410 (schema :487-490)
 * slot_mapping int64 [num_tokens], negative = padding token (cache_kernels.cu:166)
 */
int b200_reshape_and_cache(
    const void* key, const void* value, void* key_cache, void* value_cache,
    const int64_t* slot_mapping, int num_tokens, int num_heads, int head_size,
    int block_size, int x, int64_t key_stride, int64_t value_stride,
    int dtype, int kv_dtype, float k_scale, float v_scale, void* stream);

int b200_reshape_and_cache_flash(
    const void* key, const void* value, void* key_cache, void* value_cache,
    const int64_t* slot_mapping, int num_tokens, int num_heads, int head_size,
    int block_size, int64_t block_stride, int64_t key_stride, int64_t value_stride,
    int dtype, int kv_dtype, float k_scale, float v_scale, void* stream);

/* key_cache_ptrs / value_cache_ptrs: DEVICE arrays of num_layers device pointers;
 * block_mapping: DEVICE int64 [num_pairs, 2] (src, dst); block_bytes = bytes of one block of one
 * layer's key (== value) cache. Byte-exact copy. */
int b200_copy_blocks(
    const int64_t* key_cache_ptrs, const int64_t* value_cache_ptrs,
    const int64_t* block_mapping, int num_layers, int num_pairs,
    int64_t block_bytes, void* stream);

/* block_mapping: HOST int64 [num_pairs, 2]; kind: 0 = D2D, 1 = D2H, 2 = H2D
 * (one cudaMemcpyAsync per pair, as the reference: cache_kernels.cu:55-62) */
int b200_swap_blocks(
    const void* src, void* dst, const int64_t* block_mapping_host, int num_pairs,
    int64_t block_bytes, int kind, void* stream);

/* src_dtype/dst_dtype: B200_F32/F16/BF16 or -1 for the uint8 fp8 side; exactly one side is fp8.
 * kv_dtype "auto" means e4m3, as in the reference (cache_kernels.cu:375-389). */
int b200_convert_fp8(
    void* dst, const void* src, int64_t numel, int src_dtype, int dst_dtype,
    int kv_dtype, float scale, void* stream);

/* ---- normalisation / rotary / activation ------------------------------------------------------
 * replaces rms_norm            kernels/layernorm_kernels.cu:290-307 (schema torch_bindings.cpp:97-100)
 *          fused_add_rms_norm  kernels/layernorm_kernels.cu:319-352 (schema :103-106)
 *          rotary_embedding    kernels/pos_encoding_kernels.cu:128-165 (schema :110-114)
 *          batched_rotary_embedding  kernels/pos_encoding_kernels.cu:171-206 (schema :118-124)
 *          silu_and_mul / gelu_and_mul / gelu_tanh_and_mul  kernels/activation_kernels.cu:66-82
 *          gelu_new / gelu_fast / gelu_quick                kernels/activation_kernels.cu:143-160
 */
int b200_rms_norm(void* out, const void* input, const void* weight, float epsilon,
                  int num_tokens, int hidden_size, int dtype, void* stream);
int b200_fused_add_rms_norm(void* input, void* residual, const void* weight, float epsilon,
                            int num_tokens, int hidden_size, int dtype, void* stream);
/* cos_sin_cache [max_position, rot_dim]; positions int64 [num_tokens];
 * cos_sin_cache_offsets int64 [num_tokens] or NULL (plain rotary_embedding) */
int b200_rotary_embedding(const int64_t* positions, void* query, void* key,
                          const void* cos_sin_cache, const int64_t* cos_sin_cache_offsets,
                          int num_tokens, int num_heads, int num_kv_heads, int head_size,
                          int rot_dim, int64_t query_stride, int64_t key_stride,
                          int is_neox, int dtype, void* stream);
/* rotary_embedding followed by reshape_and_cache as one launch (extension; torch op _C_b200::rotary_embedding_and_cache).
 * Outputs are those of b200_rotary_embedding(query, key) then b200_reshape_and_cache(key, value, ...) — bit-exact:
 * q and k rotated in place, rotated k and v written to the paged cache at slot_mapping[t] (negative = skip the write).
 * The fused kernel serves NeoX style, rot_dim == head_size, f16/bf16, 16-byte aligned rows; anything else runs the two
 * kernels back to back inside this call. */
int b200_rotary_embedding_and_cache(const int64_t* positions, void* query, void* key, const void* value,
                                    const void* cos_sin_cache, void* key_cache, void* value_cache,
                                    const int64_t* slot_mapping, int num_tokens, int num_heads, int num_kv_heads,
                                    int head_size, int rot_dim, int64_t query_stride, int64_t key_stride,
                                    int64_t value_stride, int is_neox, int block_size, int x, int dtype,
                                    int kv_dtype, float k_scale, float v_scale, void* stream);
/* act: 0 silu_and_mul, 1 gelu_and_mul, 2 gelu_tanh_and_mul  (input [T, 2d] -> out [T, d]) */
int b200_act_and_mul(void* out, const void* input, int num_tokens, int d, int act,
                     int dtype, void* stream);
/* act: 0 gelu_new, 1 gelu_fast, 2 gelu_quick (input [T, d] -> out [T, d]) */
int b200_activation(void* out, const void* input, int num_tokens, int d, int act,
                    int dtype, void* stream);

/* ---- Marlin-format weight-only quantised GEMM -------------------------------------------------
 * replaces gptq_marlin_gemm   kernels/quantization/gptq_marlin/gptq_marlin.cu:2247-2430
 *                             (schema kernels/torch_bindings.cpp:195-201, prototype quant_ops.h:74-94)
 *          gptq_marlin_repack kernels/quantization/gptq_marlin/gptq_marlin_repack.cu:271-343 (schema :204-208)
 *          awq_marlin_repack  kernels/quantization/gptq_marlin/awq_marlin_repack.cu:208-268  (schema :211-215)
 * a [size_m, size_k] f16/bf16 contiguous; b_q_weight int32 [size_k/16, size_n*16/pack] (Marlin tiles);
 * b_scales [num_groups, size_n] (Marlin-permuted, num_groups == 1: channel-wise); c [size_m, size_n].
 * num_bits 4 or 8. has_zp: 0 = symmetric codes (uint4b8 / uint8b128, b_zeros NULL), 1 = b_zeros int32
 * [num_groups, size_n/pack] (AWQ integer zero points, Marlin layout: uint4 / uint8), 2 = b_zeros f16
 * [num_groups, size_n] (HQQ float zero points, permuted like the scales; float16 only — the reference's
 * is_zp_float, gptq_marlin.cu:2266-2270).
 * b200_marlin_gemm_plan(...) returns an UPPER BOUND of the number of fp32 partial slabs the GEMM may use for a
 * shape; when it is > 1 the caller provides c_tmp: fp32 [plan, size_m, size_n] scratch (no
 * initialisation; the reference's use_fp32_reduce buffer, gptq_marlin.cu:2313-2327) and `workspace`:
 * int32 [>= size_n/64*16], ZERO on entry and returned to zero (the reference's lock workspace,
 * torch_bindings.cpp:167-176). Partial slabs are summed in a fixed order (deterministic). split_k <= 0 = choose
 * internally (what the torch op does); size_m <= 32 with 4-bit codes runs the streaming kernel, whose stream-k
 * partition ignores split_k. Act-order with the full k range is handled by the caller
 * permuting A's columns (b200_permute_cols), as the reference does with a_tmp (gptq_marlin.cu:2145-2158). */
int b200_marlin_gemm_plan(int size_m, int size_n, int size_k, int num_groups);
/* debug only: per-role cycle attribution of the last GEMM launched with B200_MARLIN_DEBUG & 16 (32 x u64) */
int b200_debug_marlin_prof(unsigned long long* out32);
int b200_gptq_marlin_gemm(const void* a, const void* b_q_weight, const void* b_scales,
                          const void* b_zeros, void* c, float* c_tmp, int32_t* workspace,
                          int size_m, int size_n, int size_k, int num_groups, int num_bits,
                          int has_zp, int dtype, int split_k, void* stream);
/* Grouped (mixture-of-experts) W4A16 GEMM.
 * replaces marlin_gemm_moe    kernels/moe/marlin_moe_ops.cu:1482-1546 (schema kernels/moe/torch_bindings.cpp:17-23)
 * a: [size_m, size_k] (replicate_input = 1: row r of the output reads a[r / topk]) or [size_m*topk, size_k];
 * b_q_weights int32 [E, size_k/16, size_n*2] (uint4b8 Marlin tiles); b_scales [E, num_groups, size_n];
 * sorted_ids int32 [sorted_capacity]: output of moe_align_block_size(topk_ids, moe_block_size, E) (values >=
 * size_m*topk are padding); topk_ids int32 [size_m*topk]; topk_weights f32 [size_m*topk]; perm int32 [E, size_k]
 * act-order column permutation per expert (is_k_full) or NULL; c [size_m*topk, size_n]: row r = a_row(r) . W[expert(r)]
 * (* topk_weights[r] if apply_weights; rows of tokens routed to no valid expert are left untouched).
 * Scratch owned by the caller: a_sorted [sorted_capacity, size_k] of dtype, expert_offsets int32 [E + 1].
 * Launches are stream-ordered and CUDA-graph capturable (no host synchronisation: the grid is an upper bound). */
int b200_marlin_gemm_moe(const void* a, const void* b_q_weights, const int32_t* sorted_ids,
                         int64_t sorted_capacity, const float* topk_weights, const int32_t* topk_ids,
                         const void* b_scales, const int32_t* perm, void* c, void* a_sorted,
                         int32_t* expert_offsets, int size_m, int size_n, int size_k, int num_groups,
                         int num_experts, int topk, int moe_block_size, int replicate_input,
                         int apply_weights, int dtype, void* stream);
/* b_q_weight: GPTQ int32 [size_k/pack, size_n]; perm: int32 [size_k] act-order sort indices or NULL;
 * out: int32 [size_k/16, size_n*16/pack]. Bit-exact integer re-tiling. num_bits 4 or 8. */
int b200_gptq_marlin_repack(const void* b_q_weight, const int32_t* perm, void* out, int size_k,
                            int size_n, int num_bits, void* stream);
/* b_q_weight: AWQ int32 [size_k, size_n/pack] (column-interleaved) */
int b200_awq_marlin_repack(const void* b_q_weight, void* out, int size_k, int size_n, int num_bits,
                           void* stream);

/* ---- MoE routing --------------------------------------------------------------------------------
 * replaces moe_align_block_size  kernels/moe/align_block_size_kernel.cu:111-133 (schema torch_bindings.cpp:394-399)
 *          topk_softmax          kernels/moe/softmax.cu:496-518 (schema kernels/moe/torch_bindings.cpp:11-14)
 * topk_ids: int32 or int64 [numel] expert id per (token, k) slot; sorted_token_ids int32 (pre-filled with
 * numel by the caller, fused_moe.py:214-228), expert_ids int32 [max_blocks], num_tokens_post_pad int32 [1]. */
int b200_moe_align_block_size(const void* topk_ids, int ids_are_int64, int64_t numel, int num_experts,
                              int block_size, int32_t* sorted_token_ids, int32_t* expert_ids,
                              int32_t* num_tokens_post_pad, void* stream);
/* gating_output fp32 [num_tokens, num_experts]; outputs [num_tokens, topk] */
int b200_topk_softmax(float* topk_weights, int32_t* topk_indices, int32_t* token_expert_indices,
                      const float* gating_output, int num_tokens, int num_experts, int topk, void* stream);

/* Expert-loop body of the reference's dense-per-expert quantised MoE (aphrodite/modeling/models/mixtral_quant.py:141-152)
 * as one pass (extension; torch op _C_b200::moe_expert_scale_add):
 *   final[t,:] = first ? T(cur[t,:] * w[t]) : T(final[t,:] + T(cur[t,:] * w[t])),
 *   w[t] = sum_k topk_weights[t,k] * (topk_ids[t,k] == expert)        (fp32)
 * final / cur [num_tokens, hidden] f16/bf16; topk_weights f32 and topk_ids int32 [num_tokens, topk]. */
int b200_moe_expert_scale_add(void* final_out, const void* cur, const float* topk_weights, const int32_t* topk_ids,
                              int num_tokens, int hidden, int topk, int expert, int first, int dtype, void* stream);

/* ---- custom all-reduce over NVLink peer memory ---------------------------------------------------
 * replaces init_custom_ar / all_reduce_reg / all_reduce_unreg / dispose / meta_size / register_buffer /
 *          get_graph_buffer_ipc_meta / register_graph_buffers
 *          kernels/all_reduce/custom_all_reduce.cu:15-141 (schemas kernels/torch_bindings.cpp:510-535)
 * `meta`: this rank's device allocation [b200_car_meta_size() bytes of signals | scratch >= max message bytes],
 * zero-initialised; `rank_data`: device scratch for peer-pointer tables; `handles`: world_size x 64-byte
 * cudaIpcMemHandle_t (rank order; the own entry is ignored), `offsets`: byte offset of each rank's buffer inside
 * its IPC allocation. Returns an opaque handle (0 on error). world_size in {2,4,6,8}. */
int64_t b200_car_meta_size(void);
int64_t b200_car_init(void* meta, void* rank_data, int64_t rank_data_bytes, const void* handles,
                      const int64_t* offsets, int world_size, int rank, int full_nvlink);
void b200_car_dispose(int64_t fa);
int b200_car_register_buffer(int64_t fa, void* self_ptr, const void* handles, const int64_t* offsets);
/* inp must be a registered buffer, or the call must happen under stream capture (its address is then recorded
 * and registered afterwards with the two calls below). out may be any device buffer. */
int b200_car_all_reduce(int64_t fa, const void* inp, void* out, int64_t numel, int dtype, void* stream);
/* handles_out == NULL: returns the number of recorded graph buffers; otherwise fills n x 64-byte handles + offsets */
int b200_car_get_graph_buffer_ipc_meta(int64_t fa, void* handles_out, int64_t* offsets_out, int capacity);
/* handles: [world_size][n] x 64 bytes, offsets: [world_size][n] */
int b200_car_register_graph_buffers(int64_t fa, const void* handles, const int64_t* offsets, int n);

/* ---- fused tensor-parallel exchange over NVSwitch multicast (B200 design) ----------------
 * replaces the SEQUENCE  all_reduce (kernels/all_reduce/custom_all_reduce.cuh:183-249, called from
 *          aphrodite/distributed/parallel_state.py:353-379 after every row-parallel linear)
 *          -> fused_add_rms_norm (kernels/layernorm_kernels.cu:204-286, aphrodite/modeling/models/llama.py:250-256)
 * with one kernel. Every rank owns one SYMMETRIC allocation (same layout on all ranks) that is mapped on every peer
 * (`peer_bases[world]`, unicast VAs as int64; peer_bases[rank] == local_base) and, when the fabric supports it, bound
 * to one multicast object mapped at `mc_base` (NULL when the fabric has none: only B200_TP_P2P is available).
 * Inside the allocation, at byte offsets: `in_off` the rank's partial sums X [num_tokens, hidden] (the row-parallel
 * GEMM's output), `out_off` the result H [num_tokens, hidden] (identical on all ranks afterwards), `flag_off`
 * b200_tp_flag_bytes() bytes of barrier state, ZERO before the first call (and a host barrier after zeroing).
 * residual/weight != NULL: H = rms_norm(sum_r X_r + residual) * weight with the reference's rounding points, and
 * residual <- sum + residual for THIS RANK'S rows only, rows [rank*ceil(T/world), ...) — the residual stream is
 * token-sharded across ranks (each row is owned by one rank for the whole forward pass).
 * residual == weight == NULL: H = sum_r X_r (plain all-reduce, fp32 accumulate, one rounding).
 * fp16 / bf16, hidden % 8 == 0, hidden <= 8192. Stream-ordered, no host state: CUDA-graph capturable. Every rank
 * must issue the same sequence of calls (the barriers pair launches by order).
 * algo: B200_TP_P2P (unicast loads / stores / flags: needs only peer_bases), B200_TP_MC_STORE (unicast loads summed
 * in fp32 in rank order — the reference kernel's arithmetic, bit for bit — then multimem.st / multimem.red through the
 * switch: needs both mappings), B200_TP_MC_REDUCE (multimem.ld_reduce: the switch sums; its narrowing to 16 bits is
 * not round-to-nearest-even, so results are NOT bit-identical to the reference: opt-in). */
enum { B200_TP_P2P = 0, B200_TP_MC_STORE = 1, B200_TP_MC_REDUCE = 2 };
int64_t b200_tp_flag_bytes(void);
/* profiling hook: CTA 0 of the following b200_tp_allreduce_rows launches (this host thread) writes 5 %globaltimer
 * stamps (ns) into the device buffer: kernel start, start barrier passed, first row's loads landed, stores issued,
 * end barrier passed. NULL switches it off. */
void b200_tp_set_stamp_buffer(void* dev_u64x5);
int b200_tp_allreduce_rows(void* mc_base, void* local_base, const int64_t* peer_bases, int64_t in_off,
                           int64_t out_off, int64_t flag_off, void* residual, const void* weight,
                           float epsilon, int num_tokens, int hidden, int rank, int world, int dtype,
                           int algo, void* stream);

/* ---- small adjacent ops ------------------------------------------------------------------------------
 * replaces permute_cols            kernels/permute_cols.cu (schema torch_bindings.cpp:218-219): out[m,k] = a[m,perm[k]]
 *          awq_dequantize          kernels/quantization/awq/gemm_kernels.cu:720-780 (schema :147-151), fp16 only
 *          advance_step_flashattn  kernels/prepare_inputs/advance_step.cu:13-52 (schema torch_bindings.cpp:77-82) */
int b200_permute_cols(const void* a, const int32_t* perm, void* out, int64_t size_m, int size_k, void* stream);
int b200_awq_dequantize(const void* qweight, const void* scales, const void* zeros, void* out, int64_t in_c,
                        int qout_c, int group_size, void* stream);
int b200_advance_step_flashattn(int num_seqs, int num_queries, int block_size, int64_t* input_tokens,
                                const int64_t* sampled_token_ids, int64_t* input_positions, int32_t* seq_lens,
                                int64_t* slot_mapping, const int32_t* block_tables, int64_t block_tables_stride,
                                void* stream);

/* ---- fp8 activation quantisation ---------------------------------------------------------
 * replaces static_scaled_fp8_quant            kernels/quantization/fp8/common.cu:260-277 (schema torch_bindings.cpp:375)
 *          dynamic_scaled_fp8_quant           :279-299 (schema :380-382)  — *scale must be <= 0 on entry (the caller zeroes it)
 *          dynamic_per_token_scaled_fp8_quant :301-321 (schema :386-390)  — scale_ub fp32 [1] or NULL
 * out is float8_e4m3fn (1 byte / element), input float / half / bfloat16 (dtype code), contiguous. Results are
 * bit-identical to the reference kernels (multiply by 1/scale per tensor, divide by the scale per token). */
int b200_static_scaled_fp8_quant(void* out, const void* input, const float* scale, int64_t numel, int dtype,
                                 void* stream);
int b200_dynamic_scaled_fp8_quant(void* out, const void* input, float* scale, int64_t numel, int dtype,
                                  void* stream);
int b200_dynamic_per_token_scaled_fp8_quant(void* out, const void* input, float* scales, const float* scale_ub,
                                            int num_tokens, int hidden_size, int dtype, void* stream);

/* ---- W8A8 GEMM with scales ---------------------------------------------------------------
 * replaces cutlass_scaled_mm               kernels/quantization/cutlass_w8a8/scaled_mm_entry.cu:92-140 (schema
 *                                          torch_bindings.cpp:235-239, prototype kernels/ops.h)
 *          cutlass_scaled_mm_supports_fp8  scaled_mm_entry.cu:66-80 (schema torch_bindings.cpp:243-244)
 * out[M,N] = T(a_scales[m|0] * (b_scales[n|0] * sum_k a[m,k] b[k,n]) + bias[n]); a [M,K] row-major (row stride lda),
 * b [K,N] COLUMN-major (column stride ldb, i.e. the [N,K] weight), both float8_e4m3fn or both int8 (ab_dtype);
 * out fp16 / bf16 (out_dtype), row stride ldc; a_scales / b_scales fp32 with 1 or M / N entries; bias T [N] or NULL.
 * split_k <= 0 lets the library choose; workspace = fp32 scratch of b200_scaled_mm_plan(M, N, K) * M * N elements for
 * the k-split partial tiles (no initialisation needed; NULL forces one CTA per tile). The reference op has no workspace
 * argument: the torch shim allocates it from the caching allocator (graph-capturable). tcgen05 kind::f8f6f4 / kind::i8.
 * Not implemented (clear error from the torch op): the asymmetric variant cutlass_scaled_mm_azp. */
enum { B200_AB_FP8_E4M3 = 0, B200_AB_INT8 = 1 };
int b200_cutlass_scaled_mm_supports_fp8(int cuda_device_capability);
int b200_scaled_mm_plan(int size_m, int size_n, int size_k);
/* channels per CTA of the following launches on this host thread (tests / A-B measurement): 0 = auto (256 — two
 * 128-channel sub-tiles sharing one activation tile — when the 128-channel tiling exceeds one wave of the SMs, else
 * 128 with a k-split), 1 = 128, 2 = 256. Returns the previous value. */
int b200_scaled_mm_set_tile(int tile);
int b200_cutlass_scaled_mm(void* out, const void* a, const void* b, const float* a_scales, const float* b_scales,
                           const void* bias, int size_m, int size_n, int size_k, int64_t lda, int64_t ldb,
                           int64_t ldc, int a_scales_numel, int b_scales_numel, int ab_dtype, int out_dtype,
                           int split_k, void* workspace, void* stream);

/* ---- sampling --------------------------------------------------------------------------------
 * replaces sampling_from_probs, top_k_ / top_p_ / min_p_ / top_k_top_p_sampling_from_probs, top_p_renorm_prob,
 *          top_k_renorm_prob, top_k_mask_logits      kernels/sampling/sampling.cu:43-390 (schemas
 *          torch_bindings.cpp:294-350, prototypes kernels/ops.h:116-145)
 * probs / logits fp32 [batch, vocab] contiguous; uniform_samples fp32 [batch] (sampling_from_probs) or
 * [max_rounds, batch]; samples int32 [batch]; success bool [batch] (1 byte each) or NULL; per-row parameter arrays may
 * be NULL (then the scalar applies). Same uniforms -> same token as the reference's rejection loop; `deterministic` is
 * accepted for the signature and ignored (every kernel here is deterministic). */
enum { B200_SAMPLE_TOP_K = 0, B200_SAMPLE_TOP_P = 1, B200_SAMPLE_MIN_P = 2, B200_SAMPLE_TOP_K_TOP_P = 3 };
int b200_sampling_from_probs(const float* probs, const float* uniform_samples, int32_t* samples, int batch_size,
                             int vocab_size, int deterministic, void* stream);
/* mode TOP_K: (top_k_arr | top_k_val); TOP_P and MIN_P: (top_p_arr | top_p_val) carry p resp. min_p; TOP_K_TOP_P: both */
int b200_rejection_sampling_from_probs(int mode, const float* probs, const float* uniform_samples, int32_t* samples,
                                       uint8_t* success, const int32_t* top_k_arr, int top_k_val,
                                       const float* top_p_arr, float top_p_val, int batch_size, int vocab_size,
                                       int max_rounds, int deterministic, void* stream);
int b200_top_p_renorm_prob(const float* probs, float* renorm_probs, const float* top_p_arr, float top_p_val,
                           int batch_size, int vocab_size, void* stream);
int b200_top_k_renorm_prob(const float* probs, float* renorm_probs, const int32_t* top_k_arr, int top_k_val,
                           int batch_size, int vocab_size, void* stream);
int b200_top_k_mask_logits(const float* logits, float* masked_logits, const int32_t* top_k_arr, int top_k_val,
                           int batch_size, int vocab_size, void* stream);

/* ---- prefix-aware prefill attention over the paged cache -------------------------------------
 * replaces context_attention_fwd   aphrodite/attention/ops/prefix_prefill.py:696-858 (Triton; reached through
 *                                  PagedAttention.forward_prefix, attention/ops/paged_attn.py:192-228)
 * q / out [num_tokens, num_heads, D], k / v [num_tokens, num_kv_heads, D] (strides in elements: *_stride_t per token,
 * *_stride_h per head, innermost contiguous); key_cache [NB, Hkv, D/x, BS, x], value_cache [NB, Hkv, D, BS] (the
 * paged-attention layouts; block / head strides in cache elements); block_tables int32 [batch, *] (row stride
 * bt_stride); start_loc int32 [batch] first query token of each sequence; seq_lens int32 [batch] = context + query
 * length; ctx_lens int32 [batch]; alibi_slopes fp32 [num_heads] or NULL; sliding_window <= 0 disables it.
 * The reference fixes scale = 1 / sqrt(D) internally (:742); it is a parameter here and the glue passes that value.
 * Head sizes 64 / 80 / 96 / 112 / 128 / 192 / 256, block sizes 8 / 16 / 32 / 64, fp16 / bf16, kv cache auto / fp8. */
int b200_context_attention_fwd(
    const void* q, const void* k, const void* v, void* out, const void* key_cache, const void* value_cache,
    const int32_t* block_tables, const int32_t* start_loc, const int32_t* seq_lens, const int32_t* ctx_lens,
    const float* alibi_slopes, int batch, int num_heads, int num_kv_heads, int head_size, int block_size, int x,
    int max_query_len, int64_t q_stride_t, int64_t q_stride_h, int64_t k_stride_t, int64_t k_stride_h,
    int64_t v_stride_t, int64_t v_stride_h, int64_t o_stride_t, int64_t o_stride_h, int64_t kc_block_stride,
    int64_t kc_head_stride, int64_t vc_block_stride, int64_t vc_head_stride, int64_t bt_stride, float scale,
    float k_scale, float v_scale, int sliding_window, int dtype, int kv_dtype, void* stream);

/* ---- device queries -----------------------------------------------------------------------------
 * replaces get_device_attribute / get_max_shared_memory_per_block_device_attribute
 *          kernels/cuda_utils_kernels.cu (schema torch_bindings.cpp:497-504) */
int64_t b200_get_device_attribute(int64_t attribute, int64_t device_id);
int64_t b200_get_max_shared_memory_per_block_device_attribute(int64_t device_id);

#ifdef __cplusplus
}
#endif
#endif /* B200_DECODE_H_ */
