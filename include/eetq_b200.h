/*
 * eetq_b200.h -- C ABI of libeetq_b200.so: the B200 (sm_100a) w8a16 weight-only GEMM path.
 *
 * This is the drop-in boundary for the ONE hot path of NetEase-FuXi/EETQ that this repository
 * replaces.  Each entry point names the reference interface it stands in for (paths relative to
 * the EETQ tree).  The reference exposes the path through a pybind11 torch extension
 * (csrc/eetpy.cpp:7-19); here the same operations are plain C functions over raw device pointers,
 * sizes and a cudaStream_t, so that any host language can bind them (ctypes stub shown in
 * INTEGRATION.md; eetq_b200/_cabi.py is the one this repository ships).
 *
 * Conventions
 *   - All pointers are DEVICE pointers on the CURRENT CUDA device unless a name ends in `_host`.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Every call only
 *     ENQUEUES work: no synchronisation, no allocation -> all calls are CUDA-graph capturable
 *     (the reference enqueues on at::cuda::getCurrentCUDAStream, fpA_intB_gemm_wrapper.cu:151).
 *   - Return value: 0 on success, a negative EETQ_B200_E* code on failure; the message is available
 *     from eetq_b200_last_error() (thread-local).  Nothing ever aborts (the reference throws
 *     std::runtime_error / asserts: csrc/utils/cuda_utils.h:29-51, weightOnlyBatchedGemv/kernelLauncher.cu:124-127).
 *   - Logical shapes follow the reference: x [M,K] row-major activations, logical weight Wq int8 [K,N]
 *     (K = in_features, N = out_features), scales [N], y [M,N] row-major.
 *   - "b200 layout" (DESIGN.md section 3): the K*N weight bytes stored output-feature-major and biased,
 *     w_b200[n*K + k] = uint8(Wq[k, n] + 128) (the reference biases by +128 too, cutlass_preprocessors.cc:337-341;
 *     the bias lets the kernels build fp16(1024 + u) with one PRMT).  It replaces the reference's sm80 interleaved
 *     layout (cutlass_preprocessors.cc:497-534).  Requires K % 64 == 0 and N % 64 == 0 exactly like the
 *     reference (cutlass_preprocessors.cc:230,455; fpA_intB_gemm_template.h:139-142).
 */
#ifndef EETQ_B200_H_
#define EETQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EETQ_B200_VERSION 100 /* 0.1.0 */

/* activation / weight element types */
enum {
    EETQ_B200_F16  = 0,
    EETQ_B200_BF16 = 1, /* extension: the reference silently reinterprets bf16 as fp16 (fpA_intB_gemm_wrapper.cu:141-144) */
    EETQ_B200_F32  = 2  /* quantiser input only */
};

/* error codes */
enum {
    EETQ_B200_OK        = 0,
    EETQ_B200_EINVAL    = -1, /* bad argument (null pointer, unsupported dtype, shape constraint) */
    EETQ_B200_ECUDA     = -2, /* a CUDA runtime/driver call failed */
    EETQ_B200_EARCH     = -3, /* not an sm_100 device */
    EETQ_B200_EWORKSPACE = -4 /* workspace too small */
};

/* flags for eetq_b200_w8a16_gemm_ex */
enum {
    EETQ_B200_FLAG_DEFAULT    = 0,
    EETQ_B200_FLAG_FORCE_GEMV = 1, /* force the SIMT streaming kernel (M <= EETQ_B200_GEMV_MAX_M) */
    EETQ_B200_FLAG_FORCE_TC   = 2, /* force the tcgen05 kernel */
    EETQ_B200_FLAG_PDL        = 4, /* launch with programmatic-dependent-launch attribute */
    EETQ_B200_FLAG_FORCE_MMA  = 8  /* force the mma.sync streaming kernel (M <= 8) */
};

#define EETQ_B200_GEMV_MAX_M 8

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* eetq_b200_last_error(void);
int         eetq_b200_version(void);

/* ---------------------------------------------------------------------------------------------
 * Q1  quant_weights  -- replaces symmetric_quantize_last_axis_of_tensor
 *     (csrc/cutlass_kernels/fpA_intB_gemm_wrapper.cu:28-107 -> ft::symmetric_quantize,
 *      cutlass_preprocessors.cc:581-678), on the GPU, bit-exact:
 *       amax[n] = max_k |float(w[k,n])| ; s32[n] = amax[n] * (1/128)
 *       q[k,n]  = int8(clamp(round_half_away(float(w[k,n]) / s32[n]), -128, 127))
 *       scales[n] = (w dtype)(s32[n])            (all-zero column -> q = 127, scale 0, like the reference)
 *   w_kn      [K,N] row-major, dtype w_dtype (F16 / BF16 / F32)
 *   q_b200    out, K*N bytes in b200 layout (the "processed" tensor of the reference API)
 *   scales    out, [N] in w_dtype
 *   s32       out, [N] fp32 scales (also used as the abs-max workspace) -- required
 *   q_kn      out or NULL, [K,N] row-major int8 (the "unprocessed" tensor of the reference API)
 * ------------------------------------------------------------------------------------------- */
int eetq_b200_quantize(const void* w_kn, int w_dtype, int64_t K, int64_t N, int8_t* q_b200, void* scales, float* s32,
                       int8_t* q_kn, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Q2  preprocess_weights  -- replaces preprocess_weights_cuda (fpA_intB_gemm_wrapper.cu:109-128 ->
 *     ft::preprocess_weights, cutlass_preprocessors.cc:536-544): row-major int8 [K,N] -> b200 layout.
 * ------------------------------------------------------------------------------------------- */
int eetq_b200_pack(const int8_t* q_kn, int64_t K, int64_t N, int8_t* q_b200, void* stream);
/* inverse (b200 layout -> row-major int8 [K,N]); no reference counterpart, used by tests/tools */
int eetq_b200_unpack(const int8_t* q_b200, int64_t K, int64_t N, int8_t* q_kn, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Q3  checkpoint compatibility: convert between the reference's sm75..sm89 interleaved bytes
 *     (what EETQ / HF `EetqLinear.weight` checkpoints store; preprocess_weights_for_mixed_gemm,
 *     cutlass_preprocessors.cc:497-534) and the b200 layout.
 * ------------------------------------------------------------------------------------------- */
int eetq_b200_from_ref_layout(const uint8_t* w_ref, int64_t K, int64_t N, int8_t* q_b200, void* stream);
int eetq_b200_to_ref_layout(const int8_t* q_b200, int64_t K, int64_t N, uint8_t* w_ref, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1  w8_a16_gemm / w8_a16_gemm_  -- replaces w8_a16_gemm_forward_cuda(_)
 *     (fpA_intB_gemm_wrapper.cu:130-202) including its M-based dispatch (SMALL_M_FAST_PATH,
 *     fpA_intB_gemm_wrapper.h:4): small M -> SIMT streaming kernel (replaces weight_only_batched_gemv,
 *     weightOnlyBatchedGemv/kernel.h:294-468), larger M -> tcgen05 kernel (replaces the CUTLASS
 *     fpA_intB GEMM, cutlass_extensions/.../fpA_intB_gemm.h:60-486).
 *       y[m,n] = dtype( sum_k x[m,k] * q[k,n] * s[n] ) (+ bias[n], fused; the reference adds bias with a
 *       separate torch op, python/eetq/modules/qlinear.py:61)
 *   x [M,K] dtype, row stride ldx elements (ldx >= K; pass K for contiguous)
 *   w_b200 K*N int8 in b200 layout;  scales [N] dtype;  bias [N] dtype or NULL
 *   y [M,N] dtype, row stride ldy elements
 *   workspace: at least eetq_b200_workspace_bytes(M,N,K) bytes, ZERO-INITIALISED ONCE by the caller; its first
 *   4 KiB hold the stream-K hand-off flags that every call leaves at zero, the rest is fp32 partial-tile scratch, so
 *   ONE buffer sized for the largest call can be reused by calls of any shape.  If it is NULL or too small the call
 *   still succeeds with whole tiles per CTA (no K splitting: slower for small M).  One workspace must not be shared
 *   by calls that can run concurrently.
 * ------------------------------------------------------------------------------------------- */
size_t eetq_b200_workspace_bytes(int64_t M, int64_t N, int64_t K);

int eetq_b200_w8a16_gemm(const void* x, const int8_t* w_b200, const void* scales, const void* bias, void* y, int64_t M,
                         int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream);

int eetq_b200_w8a16_gemm_ex(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                            void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                            size_t workspace_bytes, int flags, void* stream);

/* Same as eetq_b200_w8a16_gemm_ex with a residual epilogue: y = dtype(dtype(acc * s [+ bias]) + residual[m, n]) (row stride ldr
 * elements; `hidden = residual + o_proj(...)`).  The reference never wired FT's residual epilogues
 * (fpA_intB_gemm_template.h:492-537) and adds the residual with a separate torch op. */
int eetq_b200_w8a16_gemm_residual(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                                  const void* residual, int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype,
                                  void* workspace, size_t workspace_bytes, int flags, void* stream);

/* Diagnostics (tools/tc_trace.py): runs the INSTRUMENTED build of the tcgen05 kernel (fp16), which records per-CTA clock
 * samples of every pipeline role into `trace` ([grid][slots] uint64, sizes from eetq_b200_w8a16_gemm_trace_info). */
int eetq_b200_w8a16_gemm_trace(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, void* y, int64_t ldy, int64_t M,
                               int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* trace, size_t trace_bytes,
                               void* stream);
int eetq_b200_w8a16_gemm_trace_info(int64_t M, int64_t N, int64_t K, int* grid, int* slots);

/* Development aid: resident CTAs per SM and resident 8-CTA clusters of the decode attention kernel on the current device. */
int eetq_b200_decode_attention_occupancy(int heads, int* ctas_per_sm, int* max_clusters);

/* Development aid: point the decode kernels' in-situ timeline recorder at a device buffer of (2 + 2 * capacity) 64-bit words
 * ([0] = record count; zero the buffer first; records are {tag << 56 | block << 16 | event, %globaltimer ns}).  Only a library
 * built with EETQ_B200_BUILD_TRACE=1 records; the product build ignores the call.  Returns 1 if recording is compiled in, else 0. */
int eetq_b200_set_timeline(void* buffer, uint64_t capacity);

/* Convenience for hosts without their own device-memory plumbing: x_host / y_host are HOST buffers
 * (pinned for true async); the call stages them through the caller-provided device scratch
 * x_dev [M*K], y_dev [M*N] on `stream` (H2D copy, kernel, D2H copy; no synchronisation). */
int eetq_b200_w8a16_gemm_host(const void* x_host, void* x_dev, const int8_t* w_b200, const void* scales,
                              const void* bias, void* y_dev, void* y_host, int64_t M, int64_t N, int64_t K, int dtype,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Extensions either side of the w8a16 linears (SURVEY.md section 8 "next", ranks 2 and 4) -- NOT part of the reference's
 * w8a16 boundary, except eetq_b200_rotary_embedding_neox / eetq_b200_layernorm_forward which stand in for the two glue
 * ops the reference module also exports (csrc/eetpy.cpp:18-19).  They exist so that the headline metric (Llama-2-7B decode
 * tokens/s) is bounded by the weight stream, not by framework-op launches.  fp16.  `pdl` != 0 launches with programmatic
 * dependent launch.
 *
 * Multi-GPU exchange ("LL" buffers): a vector of E fp16 values that every rank needs is kept, on every rank, as E/2 8-byte
 * words {two values, 32-bit tag}.  A producer stores its words with single 8-byte stores straight into every rank's copy
 * (peer-mapped symmetric memory over NVLink); a consumer polls the words it is about to use until their tag matches
 * tag = *step * per_step + index + 1.  `step` is a device int32 that eetq_b200_lm_head_argmax increments once per decode
 * step; `index` numbers the exchanges inside one step (0 <= index < per_step).  This is the fused replacement of "one
 * all-gather of the activations per sharded linear" (SURVEY.md section 8e): no collective call, no fence, no flag.
 * ------------------------------------------------------------------------------------------- */
typedef struct eetq_b200_ll {        /* consumer side: which exchange a buffer currently carries */
    const void* step;                /* device int32 step counter; NULL = the buffer is a plain fp16 vector */
    int per_step;
    int index;
} eetq_b200_ll;

typedef struct eetq_b200_ll_push {   /* producer side */
    int world;                       /* number of ranks (1..8) */
    const uint64_t* peers;           /* [world] HOST array: device address of the LL buffer on every rank, index = rank */
    void* local;                     /* this rank's own copy */
    int64_t elem_off;                /* first element of this rank's slice inside the full vector (even) */
    const void* step;                /* device int32 step counter */
    int per_step;
    int index;
} eetq_b200_ll_push;

/* eetq_b200_w8a16_gemv_fused: the decode GEMV (M <= 8) with glue folded in (all fields optional / zero):
 *   Every CTA requests its first two row groups into registers BEFORE the dependency wait (programmatic dependent launch); given
 *   next_w, the grid also asks L2 (cp.async.bulk.prefetch.L2) for the head of the NEXT GEMV's weights, whose first CTAs then run out
 *   of L2, finish early and let the kernel after them start its weight stream sooner.
 *   xmode 0: plain;  1: x := RMSNorm(x; norm_weight, eps) on load (HF LlamaRMSNorm arithmetic; replaces the reference's
 *   separate generalT5LayerNorm launch);  2: x := silu(x[:, :K]) * x[:, K:2K] (ldx >= 2K);
 *   epi 0: plain;  1: weight rows are interleaved (gate_0, up_0, gate_1, up_1, ...) and the kernel emits the N/2 values
 *   fp16(silu(fp16 gate)) * fp16 up  (`act_fn(gate_proj(x)) * up_proj(x)`), no bias;
 *   residual != NULL: y = dtype(acc * s [+ bias]) + residual  (the reference never wired FT's residual epilogues,
 *   fpA_intB_gemm_template.h:492-537);
 *   x_ll / residual_ll / push (M = 1, fp16): x / residual are LL buffers of the FULL vectors (residual_off = element
 *   offset of this rank's output slice), outputs are pushed to every rank's LL buffer instead of being stored to y. */
typedef struct eetq_b200_gemv_opts {
    const void* norm_weight;
    float eps;
    int xmode;
    int epi;
    const void* residual;
    int64_t ldr;
    const eetq_b200_ll* x_ll;
    const eetq_b200_ll* residual_ll;
    int64_t residual_off;
    const eetq_b200_ll_push* push;
    const void* next_w;              /* optional L2 staging hint: the b200 weight matrix [next_n rows][next_k bytes] of the decode GEMV */
    int64_t next_n, next_k;          /* that runs next; this kernel asks L2 for the first megabytes of it (its first CTAs' slices) */
} eetq_b200_gemv_opts;
int eetq_b200_w8a16_gemv_fused(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias, void* y,
                               int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, const eetq_b200_gemv_opts* opts, int pdl,
                               void* stream);

/* x[0:H] = table[*token] (plain vector, or LL words with the tag of x_ll when x_ll != NULL) */
int eetq_b200_decode_embed(const void* table, const void* token_i64, void* x, int64_t H, const eetq_b200_ll* x_ll, int pdl, void* stream);
/* y[m] = RMSNorm(x[m]) * w over M rows of H (HF LlamaRMSNorm arithmetic), row strides ldx / ldy elements */
int eetq_b200_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int64_t M, int64_t H, float eps, int pdl, void* stream);
/* layernorm_forward of the reference (csrc/eetpy.cpp:19 -> layernorm_kernels/layernorm.cu:88-110): fp16 [m, n] rows,
 * T5-style RMS norm, out = fp16(clamp((x * rsqrt(mean(x^2) + eps)) * gamma)) with ONE rounding. */
int eetq_b200_layernorm_forward(const void* input, const void* gamma, void* out, int64_t m, int64_t n, float eps, void* stream);
/* rotary_embedding_neox of the reference (csrc/eetpy.cpp:18 -> embedding_kernels/pos_encoding_kernels.cu:55-87): in place on
 * query and key [num_tokens, num_heads, head_size] fp16; cos_sin_cache [max_position, rot_dim] = cos | sin halves. */
int eetq_b200_rotary_embedding_neox(const void* positions_i64, void* query, void* key, int64_t num_tokens, int64_t num_heads,
                                    int64_t head_size, const void* cos_sin_cache, int64_t rot_dim, void* stream);
/* Prefill glue: rotate q (in place) and k of T tokens at cache positions p0 .. p0+T-1 (HF rotate_half convention, cos/sin
 * tables [max_pos][D/2]) and write rotated k and v into the head-major KV cache [heads][max_ctx][D]; qkv rows = q | k | v. */
int eetq_b200_prefill_rope_kv(void* qkv, int64_t ld, const void* cos_t, const void* sin_t, void* kcache, void* vcache, int64_t T,
                              int64_t heads, int64_t D, int64_t max_ctx, int64_t p0, void* stream);
/* Prefill glue: act[t][i] = fp16(silu(gate)) * up; interleaved = 0: rows are gate[I] | up[I], 1: (g0, u0, g1, u1, ...) */
int eetq_b200_silu_mul(const void* gu, int64_t ldg, void* act, int64_t lda, int64_t T, int64_t I, int interleaved, void* stream);
/* Fused RoPE (rotate_half convention) + KV-cache append + attention for ONE token at position *pos, head_dim 128, ONE
 * launch: grid (local heads, 4) in clusters of 4 CTAs per head; the KV stream is moved by bulk copies (TMA) into shared memory
 * and starts before the dependency wait; the 4 partial results of a head are merged through distributed shared memory.
 *   qkv [3 * H_local] = q | k | v raw projections of this rank's heads; cos/sin [max_pos][D/2]; kcache/vcache
 *   [H_local/D][max_ctx][D] head-major (row *pos of every head is written); out [H_local] plain fp16, or NULL and `push`
 *   describes the LL exchange of the full attention vector (elem_off = first element of this rank's heads).
 *   next_w / next_n / next_k (optional): the weight matrix of the GEMV that follows (o_proj); its head is requested into L2
 *   while the attention itself leaves HBM mostly idle (see eetq_b200_gemv_opts.next_w). */
int eetq_b200_decode_attention(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                               void* vcache, void* out, int64_t H_local, int64_t D, int64_t max_ctx, const eetq_b200_ll_push* push,
                               const void* next_w, int64_t next_n, int64_t next_k,
                               int pdl, void* stream);
/* Final RMSNorm + fp16 lm_head (the reference leaves lm_head unquantised, quantizer.py:40) + greedy arg-max in ONE launch:
 *   logits[v] = RMSNorm(x; norm_w, eps) . W[v, :] over this rank's V_local vocabulary rows (global ids v_begin ..); the last CTA
 *   picks the winner (first maximum), exchanges candidates with the other ranks when `cand` is given, writes *token (int64),
 *   and advances *pos and *step by one.  logits (optional) receives the fp16 logits of this rank's rows at their global ids.
 *   scratch: eetq_b200_lm_head_scratch_bytes() bytes, zero on first use (left clean). */
size_t eetq_b200_lm_head_scratch_bytes(void);
int eetq_b200_lm_head_argmax(const void* x, const eetq_b200_ll* x_ll, const void* norm_w, float eps, const void* w, int64_t V_local,
                             int64_t H, int64_t v_begin, void* logits, void* scratch, void* token_i64, void* pos_i32, void* step_i32,
                             const eetq_b200_ll_push* cand, int rank, int pdl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Packed int4 weights (QuantType::PACKED_INT4_WEIGHT_ONLY).  In the reference int4 is reachable from Python only as
 * quant_weights(w, torch.quint4x2, ..) and preprocess_weights(w, is_int4=True) (csrc/eetpy.cpp:11-17): its w8_a16_gemm
 * hard-codes Int8b (fpA_intB_gemm_wrapper.cu:154-159) although the int4 kernels are compiled into the extension
 * (weightOnlyBatchedGemv/kernel.h:68-116, ...Bs{1..4}Int4b.cu).  Here both halves exist.
 *
 * "b200 int4 layout": K*N/2 bytes, output-feature-major rows of K/2 bytes; in every 32-bit word (8 consecutive k of
 * one output feature) nibble p < 4 holds q[8j + 2p] + 8 and nibble 4 + p holds q[8j + 2p + 1] + 8, so that
 * (word >> 4p) & 0x000f000f is the adjacent-k pair (the reference interleaves nibbles the same way for the same reason,
 * cutlass_preprocessors.cc:389-417).  Packed row-major ("unprocessed") tensors are [K, N/2] bytes, low nibble = even
 * column (cutlass_preprocessors.cc:651-669).
 * ------------------------------------------------------------------------------------------- */
#define EETQ_B200_GEMV4_MAX_M 8      /* rows served by the streaming kernels */
#define EETQ_B200_GEMV4_SIMT_MAX_M 4 /* rows the SIMT kernel takes (K % 128 != 0, or forced) */

/* quant_weights(w, quint4x2): s32[n] = amax[n] * (1/8); q = clamp(int(round_half_away(w / s32[n])), -8, 7), NaN -> -8
 * (bit-exact with ft::symmetric_quantize, cutlass_preprocessors.cc:608-669).  q4_b200: K*N/2 bytes out;
 * q4_kn: packed row-major [K, N/2] out or NULL; scales / s32 as in eetq_b200_quantize. */
int eetq_b200_quantize4(const void* w_kn, int w_dtype, int64_t K, int64_t N, uint8_t* q4_b200, void* scales, float* s32,
                        uint8_t* q4_kn, void* stream);
/* preprocess_weights(w, is_int4=True): packed row-major [K, N/2] -> b200 int4 layout, and its inverse */
int eetq_b200_pack4(const uint8_t* q4_kn, int64_t K, int64_t N, uint8_t* q4_b200, void* stream);
int eetq_b200_unpack4(const uint8_t* q4_b200, int64_t K, int64_t N, uint8_t* q4_kn, void* stream);
/* checkpoint compatibility: the reference's sm75..sm89 int4 bytes (preprocess_weights_for_mixed_gemm with
 * PACKED_INT4_WEIGHT_ONLY, cutlass_preprocessors.cc:497-534) <-> b200 int4 layout */
int eetq_b200_from_ref_layout4(const uint8_t* w4_ref, int64_t K, int64_t N, uint8_t* q4_b200, void* stream);
int eetq_b200_to_ref_layout4(const uint8_t* q4_b200, int64_t K, int64_t N, uint8_t* w4_ref, void* stream);
/* y = dtype( sum_k x[m,k] * q[k,n] * s[n] ) (+ bias) with int4 weights.  M <= EETQ_B200_GEMV4_MAX_M streams the nibbles
 * straight through the decode kernels (half the HBM bytes of w8a16): one row through the SIMT kernel, 2..8 rows through the
 * mma.sync streaming kernel.  Larger M (or 5..8 rows with K % 128 != 0) first widens the weights to the b200 int8 layout in
 * `workspace` (one extra HBM pass of 1.5 * K*N bytes, small next to the tensor-core GEMM it feeds), then runs the tcgen05
 * kernel: workspace must hold eetq_b200_w4a16_workspace_bytes(M,N,K) bytes (zero-initialised once), else
 * EETQ_B200_EWORKSPACE.  flags: EETQ_B200_FLAG_PDL, FORCE_GEMV (SIMT, M <= 4), FORCE_MMA (M <= 8). */
size_t eetq_b200_w4a16_workspace_bytes(int64_t M, int64_t N, int64_t K);
int eetq_b200_w4a16_gemm(const void* x, int64_t ldx, const uint8_t* q4_b200, const void* scales, const void* bias, void* y,
                         int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes,
                         int flags, void* stream);

/* number of kernels this library has launched on this process so far (bench.py's gpu_launches) */
uint64_t eetq_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EETQ_B200_H_ */
