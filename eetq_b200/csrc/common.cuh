// common.cuh -- shared host/device helpers for libeetq_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/eetq_b200.h"

namespace eetq_b200 {

// ---- error plumbing (thread-local message, never abort) -------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define EB_CHECK_ARG(cond, ...)                                                                                        \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            ::eetq_b200::set_error(__VA_ARGS__);                                                                       \
            return EETQ_B200_EINVAL;                                                                                   \
        }                                                                                                              \
    } while (0)

#define EB_CHECK_CUDA(expr)                                                                                            \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            ::eetq_b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);        \
            return EETQ_B200_ECUDA;                                                                                    \
        }                                                                                                              \
    } while (0)

// per-device properties, fetched once (the reference re-queries these on every forward call,
// fpA_intB_gemm_template.h:390-397)
struct DeviceInfo {
    int sm_count     = 0;
    int cc_major     = 0;
    int cc_minor     = 0;
    int max_smem_optin = 0;
    bool ok          = false;
};
const DeviceInfo& device_info();  // for the current device; ok == false if the query failed

// ---- launchers implemented in the .cu files ---------------------------------------------------------
int launch_quantize(const void* w_kn, int w_dtype, int64_t K, int64_t N, int8_t* q_b200, void* scales, float* s32,
                    int8_t* q_kn, cudaStream_t stream);
int launch_transpose_bytes(const int8_t* src, int64_t rows, int64_t cols, int8_t* dst, cudaStream_t stream);
int launch_from_ref_layout(const uint8_t* w_ref, int64_t K, int64_t N, int8_t* q_b200, cudaStream_t stream);
int launch_to_ref_layout(const int8_t* q_b200, int64_t K, int64_t N, uint8_t* w_ref, cudaStream_t stream);
// packed int4 (quantize.cu)
int launch_quantize4(const void* w_kn, int w_dtype, int64_t K, int64_t N, uint8_t* q4_b200, void* scales, float* s32,
                     uint8_t* q4_kn, cudaStream_t stream);
int launch_nibble_layout(int mode, const uint8_t* src, int64_t K, int64_t N, uint8_t* dst, cudaStream_t stream);
int launch_widen4to8(const uint8_t* q4_b200, int64_t K, int64_t N, int8_t* q_b200, cudaStream_t stream);

enum { GEMV_X_PLAIN = 0, GEMV_X_RMSNORM = 1, GEMV_X_SILU_MUL = 2 };
enum { GEMV_EPI_PLAIN = 0, GEMV_EPI_SILU_PAIRS = 1 };

// ---- "LL" exchange of small vectors between ranks (column-sharded multi-GPU decode) -------------
// A vector of E fp16 values travels as E/2 8-byte words {value pair, 32-bit tag}.  The producer stores each word with ONE
// 8-byte store (single-copy atomic) straight into every rank's copy of the buffer (peer-mapped symmetric memory, NVLink);
// the consumer polls the words it needs until their tag equals the tag of the exchange.  Data and flag arrive together, so
// there is no fence, no separate flag and no collective on the critical path.  tag = step * per_step + index + 1: `step` is a
// device counter that the last kernel of a decode step increments, `index` numbers the exchanges inside a step; a buffer is
// only rewritten several exchanges later, when (by the data dependencies of the layer chain) every rank has consumed it.
struct LLTag {
    const int* tag_base = nullptr;  // device step counter; nullptr = not an LL buffer
    int per_step        = 0;
    int index           = 0;
};
struct LLPush {
    unsigned long long peer[8] = {};      // base address of the LL buffer on every rank (index = rank), peer-mapped
    unsigned long long* local  = nullptr; // this rank's own copy (for consumers inside the same kernel)
    int elem_off               = 0;       // first element of this rank's slice inside the full vector
    int world                  = 0;       // 0 = no push
    LLTag tag;
};
__device__ __forceinline__ uint32_t ll_tag(const LLTag& t)
{
    return uint32_t(*t.tag_base) * uint32_t(t.per_step) + uint32_t(t.index) + 1u;
}
__device__ __forceinline__ unsigned long long ll_pack(uint32_t data, uint32_t tag)
{
    return (static_cast<unsigned long long>(tag) << 32) | data;
}
// word index is relative to this rank's slice (elem_off / 2 is added)
__device__ __forceinline__ void ll_push_word(const LLPush& p, int local_word, uint32_t data)
{
    const unsigned long long v   = ll_pack(data, ll_tag(p.tag));
    const unsigned long long off = (static_cast<unsigned long long>(p.elem_off >> 1) + static_cast<unsigned long long>(local_word)) * 8ull;
#pragma unroll 1
    for (int r = 0; r < p.world; ++r)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.peer[r] + off), "l"(v) : "memory");
}
// N (even) consecutive LL words -> their N data halves; spins until all tags match
template <int N>
__device__ __forceinline__ void ll_load_words(const unsigned long long* src, uint32_t tag, uint32_t (&out)[N])
{
    static_assert(N % 2 == 0, "pairs of words (16-byte loads)");
    unsigned long long v[N];
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(v[i]), "=l"(v[i + 1]) : "l"(src + i) : "memory");
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
            ok = ok && (uint32_t(v[i] >> 32) == tag);
    } while (!ok);
#pragma unroll
    for (int i = 0; i < N; ++i)
        out[i] = uint32_t(v[i] & 0xffffffffull);
}

struct GemvExtras {
    const void* norm_weight = nullptr;  // [K], GEMV_X_RMSNORM
    const void* residual    = nullptr;  // [M, N] row stride ldr, or LL words (res_ll)
    int64_t ldr             = 0;
    float eps               = 0.f;
    int xmode               = GEMV_X_PLAIN;
    int epi                 = GEMV_EPI_PLAIN;
    LLTag x_ll;     // x is an LL buffer of the full K-vector
    LLTag res_ll;   // residual is an LL buffer of the full output vector; this rank's slice starts at element res_off
    int res_off = 0;
    LLPush push;    // outputs go to every rank's LL buffer instead of y
    // L2 staging hint: the weight matrix of the decode GEMV that runs NEXT.  At its start this kernel asks L2 for the first
    // megabytes of that matrix (the complete slices of its first CTAs): those CTAs then run out of L2, finish early and make room
    // for the kernel after them, which starts its own weight stream sooner (measured +3.3 % decode tokens/s).
    const void* next_w = nullptr;
    int64_t next_n = 0, next_k = 0;
    int wbits = 8;  // 8: b200 int8 layout; 4: b200 int4 layout (M <= 4)
};
int launch_gemv(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                int M, int64_t N, int64_t K, int dtype, const GemvExtras& ex, bool pdl, cudaStream_t stream);

// argument checks shared by every forward entry point (cabi.cu)
int check_arch();
int check_forward_args(const char* who, const void* x, int64_t ldx, const void* w, const void* scales, const void* y, int64_t ldy,
                       int64_t M, int64_t N, int64_t K, int dtype, int64_t n_out = -1);

// tensor-core (mma.sync) streaming kernel for up to 8 decode rows, weights in the MMA's A role, int8 or int4 weights (gemv_mma.cu)
bool gemv_mma_supported(int M, int64_t K, int wbits);
int launch_gemv_mma(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                    int64_t ldr, void* y, int64_t ldy, int M, int64_t N, int64_t K, int dtype, int wbits, bool pdl,
                    cudaStream_t stream);

size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemm_tc(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                   int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                   size_t workspace_bytes, bool pdl, unsigned long long* trace, cudaStream_t stream);
int gemm_tc_trace_slots();
int gemm_tc_grid_for(int64_t M, int64_t N, int64_t K);

// ---- small device helpers ----------------------------------------------------------------------------
template <typename T>
struct DTypeOf;
template <>
struct DTypeOf<__half> {
    static constexpr int value = EETQ_B200_F16;
};
template <>
struct DTypeOf<__nv_bfloat16> {
    static constexpr int value = EETQ_B200_BF16;
};

__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_float(float v) { return v; }

template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }

// 128-bit streaming load that does not allocate in L1 (weights are read exactly once)
__device__ __forceinline__ uint4 ldg_stream_128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- in-situ timeline (development builds only: EETQ_B200_BUILD_TRACE=1 -> -DEETQ_B200_TRACE) ---------------------------------
// Thread 0 of every CTA appends {tag << 56 | block << 16 | event, %globaltimer} records to a device buffer ([0] = record count).
// Every translation unit holds its own copy of the pointer (no relocatable device code); cabi.cu sets them all.
#ifdef EETQ_B200_TRACE
static __device__ unsigned long long* g_trace_buf = nullptr;
static __device__ unsigned int g_trace_cap        = 0;
__device__ __forceinline__ void trace_ev(int tag, int ev)
{
    if (threadIdx.x == 0 && g_trace_buf != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long i = atomicAdd(g_trace_buf, 1ull);
        if (i < g_trace_cap) {
            const unsigned long long blk = blockIdx.x + blockIdx.y * gridDim.x;
            g_trace_buf[2 + 2 * i]     = (static_cast<unsigned long long>(tag) << 56) | (blk << 16) | static_cast<unsigned long long>(ev);
            g_trace_buf[3 + 2 * i]     = t;
        }
    }
}
#define EB_TRACE_SETTER(name)                                                                                          \
    void name(void* buf, unsigned int cap)                                                                             \
    {                                                                                                                  \
        unsigned long long* b = static_cast<unsigned long long*>(buf);                                                 \
        cudaMemcpyToSymbol(g_trace_buf, &b, sizeof(b));                                                                \
        cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));                                                            \
    }
#else
__device__ __forceinline__ void trace_ev(int, int) {}
#define EB_TRACE_SETTER(name) \
    void name(void*, unsigned int) {}
#endif
void trace_set_gemv(void* buf, unsigned int cap);
void trace_set_decode(void* buf, unsigned int cap);
enum { TRACE_GEMV = 1, TRACE_ATTN = 2, TRACE_LMHEAD = 3, TRACE_EMBED = 4 };

// Fire-and-forget bulk prefetch of [p, p + bytes) into L2 (bytes a multiple of 16, p 16-byte aligned): one instruction per
// chunk, no registers, no completion tracking.  Used to keep HBM streaming while a kernel sits in its dependency wait.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// One thread walks a byte range in 32 KB pieces (the instruction takes its operands from uniform registers: issuing it from many
// lanes with different addresses would serialise lane by lane)
__device__ __forceinline__ void l2_prefetch_range(const uint8_t* base, long long bytes)
{
    constexpr long long kPiece = 32768;
    for (long long off = 0; off < bytes; off += kPiece) {
        const long long n = bytes - off < kPiece ? bytes - off : kPiece;
        l2_prefetch_bulk(base + off, static_cast<uint32_t>(n & ~15ll));
    }
}

// L2 staging hint (GemvExtras::next_w): the contiguous HEAD of the next decode GEMV's weight matrix, i.e. the complete slices of
// its first CTAs.  Every CTA of the requesting kernel asks for an equal share at its start.
struct NextHint {
    const uint8_t* w = nullptr;
    long long bytes  = 0;
};
NextHint make_next_hint(const void* w, int64_t N, int64_t K);  // gemv.cu
__device__ __forceinline__ void l2_prefetch_next(const NextHint& h, int cta, int ncta)
{
    const long long per = ((h.bytes / ncta) + 15) & ~15ll;
    const long long b0  = per * cta;
    if (b0 < h.bytes)
        l2_prefetch_range(h.w + b0, (h.bytes - b0 < per) ? h.bytes - b0 : per);
}

// programmatic dependent launch (PDL) controls; no-ops when the kernel was launched without the attribute
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace eetq_b200
