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
    EETQ_B200_FLAG_PDL        = 4  /* launch with programmatic-dependent-launch attribute */
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

/* Convenience for hosts without their own device-memory plumbing: x_host / y_host are HOST buffers
 * (pinned for true async); the call stages them through the caller-provided device scratch
 * x_dev [M*K], y_dev [M*N] on `stream` (H2D copy, kernel, D2H copy; no synchronisation). */
int eetq_b200_w8a16_gemm_host(const void* x_host, void* x_dev, const int8_t* w_b200, const void* scales,
                              const void* bias, void* y_dev, void* y_host, int64_t M, int64_t N, int64_t K, int dtype,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Decode-side extensions (SURVEY.md section 8 "next", rank 2 and 4) -- NOT part of the reference's w8a16 boundary.
 * They exist so the headline metric (Llama-2-7B decode tokens/s) is bounded by the weight stream, not by
 * framework-op launches.  fp16 only, single token.  `pdl` != 0 launches with programmatic dependent launch.
 *
 * eetq_b200_w8a16_gemv_fused: the decode GEMV (M <= 8) with glue folded in:
 *   xmode 0: plain;  1: x := RMSNorm(x; norm_weight, eps) on load (HF LlamaRMSNorm arithmetic; replaces the
 *   reference's separate generalT5LayerNorm kernel, csrc/layernorm_kernels/layernorm.cu:25-51);
 *   2: x := silu(x[:, :K]) * x[:, K:2K] (fused gate|up activation; ldx >= 2K);
 *   residual != NULL: y = dtype(acc * s [+ bias]) + residual  (the reference never wired FT's residual epilogues,
 *   fpA_intB_gemm_template.h:492-537).
 * ------------------------------------------------------------------------------------------- */
int eetq_b200_w8a16_gemv_fused(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                               const void* norm_weight, float eps, int xmode, const void* residual, int64_t ldr, void* y,
                               int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, int pdl, void* stream);
/* eetq_b200_w8a16_gemv_fused that also issues an L2 prefetch of the KV-cache rows [0, *pos) ([heads][max_ctx][128] fp16)
 * the attention kernel launched right after it will read (the GEMV's own weight stream is already in flight). */
int eetq_b200_w8a16_gemv_fused_kvprefetch(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* norm_weight,
                                          float eps, int xmode, const void* residual, int64_t ldr, void* y, int64_t ldy, int64_t M,
                                          int64_t N, int64_t K, int dtype, const void* kcache, const void* vcache, const void* pos_i32,
                                          int64_t heads, int64_t max_ctx, int pdl, void* stream);

/* Up to 4 DEPENDENT decode GEMVs (M = 1, fp16; e.g. o_proj -> gate|up -> down -> next layer's q|k|v) in ONE launch:
 * the CTAs meet at a grid barrier between phases but issue the next phase's first weight loads before waiting, so the
 * HBM stream does not stop at what would otherwise be kernel boundaries.  counters: nphases-1 uint32 private to this call
 * site, zero on first use (they only increase); epoch: device int32 >= 1 that increases by exactly 1 per launch. */
typedef struct eetq_b200_gemv_phase {
    const void* x;            /* activation [K] (or [2K] for xmode 2) */
    int64_t ldx;
    const void* w;            /* b200 layout, N*K bytes */
    const void* scales;       /* [N] fp16 */
    void* y;                  /* [N] fp16 */
    int64_t N, K;
    const void* norm_weight;  /* xmode 1 */
    const void* residual;     /* [N] or NULL */
    float eps;
    int xmode;                /* 0 plain, 1 RMSNorm on load, 2 silu(x[:K]) * x[K:2K] */
} eetq_b200_gemv_phase;
int eetq_b200_w8a16_gemv_chain(const void* phases, int nphases, void* counters, const void* epoch, int pdl, void* stream);

/* Column-sharded decode GEMV with the activation all-gather FUSED into the epilogue over NVLink peer memory
 * (one all-gather per sharded linear).  Rank r owns rows [r*N/P, (r+1)*N/P) of the linear; every
 * rank stores its slice straight into all ranks' copies of the output vector (peer-mapped symmetric memory) and the last
 * CTA publishes flags[slot][rank] = *epoch on every peer (st.release.sys).  The kernel that CONSUMES the vector waits for
 * all ranks' flags of that slot (`wait_flags`, ld.acquire.sys) in its prologue.  peer_y[r]: rank r's output buffer offset to this rank's
 * first row; peer_flag[r]: address on rank r of flags[slot][this rank]; local_flags: this rank's flags[slot][0..world);
 * ticket: local uint32 (zero between calls); epoch: device int32, strictly increasing per decode step. */
int eetq_b200_w8a16_gemv_fused_p2p(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* norm_weight,
                                   float eps, int xmode, const void* residual, int64_t ldr, int64_t M, int64_t N_local, int64_t K,
                                   int dtype, int world, const uint64_t* peer_y, const uint64_t* peer_flag, const void* local_flags,
                                   const void* wait_flags, void* ticket, const void* epoch, int64_t ldy, int pdl, void* stream);
/* Consumers of a gathered buffer poll the flags of the call that produced it (wait_flags = that call's local_flags, or
 * NULL when the input is local) in their prologue -- after their own weight / KV prefetch -- instead of the producer
 * waiting at its tail.  Same kernels as eetq_b200_decode_rmsnorm / _attention with that wait added. */
int eetq_b200_decode_rmsnorm_p2p(const void* x, const void* w, void* y, int64_t M, int64_t H, float eps, const void* wait_flags,
                                 int world, const void* epoch, int pdl, void* stream);
int eetq_b200_decode_attention_p2p(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                                   void* vcache, void* partial, void* tickets, void* out, int64_t H, int64_t D, int64_t max_ctx,
                                   const void* wait_flags, int world, const void* epoch, int pdl, void* stream);
/* x[0:H] = table[*token] */
int eetq_b200_decode_embed(const void* table, const void* token_i64, void* x, int64_t H, int pdl, void* stream);
/* y = RMSNorm(x) * w over M rows of H */
int eetq_b200_decode_rmsnorm(const void* x, const void* w, void* y, int64_t M, int64_t H, float eps, int pdl, void* stream);
/* Fused RoPE (rotate_half convention; behaviour of rotary_embedding_neox, csrc/embedding_kernels/pos_encoding_kernels.cu:12-53)
 * + KV-cache append + split-KV attention + split merge for ONE token at position *pos, head_dim 128, ONE launch.
 *   qkv [3H] = q | k | v raw projections; cos/sin [max_pos][D/2]; kcache/vcache [H/D][max_ctx][D], head-major so each
 *   CTA streams one contiguous block (row *pos of every head is written);
 *   partial: (H/D) * eetq_b200_decode_attention_splits(max_ctx) * 130 floats of scratch; tickets: H/D int32, zero on first
 *   use (left zero); out [H]. */
int64_t eetq_b200_decode_attention_splits(int64_t max_ctx);
int eetq_b200_decode_attention(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                               void* vcache, void* partial, void* tickets, void* out, int64_t H, int64_t D, int64_t max_ctx,
                               int pdl, void* stream);

/* number of kernels this library has launched on this process so far (bench.py's gpu_launches) */
uint64_t eetq_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EETQ_B200_H_ */
