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

enum { GEMV_X_PLAIN = 0, GEMV_X_RMSNORM = 1, GEMV_X_SILU_MUL = 2 };
// Fused all-gather over NVLink peer memory (column-sharded linears): instead of writing its output
// slice locally and calling NCCL, the GEMV epilogue stores the slice into EVERY rank's activation buffer (peer-mapped
// symmetric memory), then the last CTA publishes a per-call epoch flag on every peer and waits until all peers' flags
// for the same call have arrived -- when the kernel completes, the full activation vector is present on this rank.
struct GemvP2P {
    unsigned long long peer_y[8]    = {};  // rank p's output buffer, already offset to THIS rank's first row
    unsigned long long peer_flag[8] = {};  // address, on rank p, of flags[slot][this rank]
    const unsigned* local_flags     = nullptr;  // this rank's flags[slot][0..world)
    unsigned* ticket                = nullptr;  // local CTA ticket (zero between calls)
    const int* epoch                = nullptr;  // device counter, strictly increasing per decode step
    int world                       = 1;
    // consumer side: flags of the call that produced THIS kernel's input (all ranks must have published `*epoch` there
    // before the activation is read); nullptr = the input was produced locally
    const unsigned* wait_flags      = nullptr;
};
struct GemvExtras {
    const void* norm_weight = nullptr;  // [K], GEMV_X_RMSNORM
    const void* residual    = nullptr;  // [M, N] row stride ldr
    int64_t ldr             = 0;
    float eps               = 0.f;
    int xmode               = GEMV_X_PLAIN;
    GemvP2P p2p;
    // optional L2 prefetch issued by the GEMV CTAs once their own weight stream is fully in flight: the KV cache rows
    // [0, *pf_pos) of every head (layout [heads][max_ctx][128] fp16) that the NEXT kernel (attention) will read
    const void* pf_k   = nullptr;
    const void* pf_v   = nullptr;
    const int* pf_pos  = nullptr;
    int pf_heads       = 0;
    int pf_max_ctx     = 0;
};
int launch_gemv(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                int M, int64_t N, int64_t K, int dtype, const GemvExtras& ex, bool pdl, cudaStream_t stream);

// one phase of a chained decode GEMV launch (M = 1, fp16); mirrors eetq_b200_gemv_phase in the public header
struct GemvChainPhase {
    const void* x;
    int64_t ldx;
    const void* w;
    const void* scales;
    void* y;
    int64_t N;
    int64_t K;
    const void* norm_weight;
    const void* residual;
    float eps;
    int xmode;
};
int launch_gemv_chain(const GemvChainPhase* phases, int nphases, unsigned* counters, const int* epoch, bool pdl, cudaStream_t stream);

// tensor-core (mma.sync) streaming kernel for 2 <= M <= 8 decode rows (gemv_mma.cu)
bool gemv_mma_supported(int M, int64_t K);
int launch_gemv_mma(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                    int64_t ldr, void* y, int64_t ldy, int M, int64_t N, int64_t K, int dtype, bool pdl, cudaStream_t stream);

size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemm_tc(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                   int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                   size_t workspace_bytes, bool pdl, unsigned long long* trace, cudaStream_t stream);
int gemm_tc_trace_slots();
int gemm_tc_grid_for(int64_t M, int64_t N, int64_t K);
#ifdef EETQ_B200_WITH_V1
// round-1 kernel, A/B baseline only (EETQ_B200_TC_IMPL=v1)
size_t gemm_tc_v1_workspace_bytes(int64_t M, int64_t N, int64_t K);
int launch_gemm_tc_v1(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y,
                      int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes,
                      bool pdl, cudaStream_t stream);
#endif

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

// wait until every rank has published `epoch` in flags[0..world) (written by peers with st.release.sys); all threads of
// the CTA must call it (contains a __syncthreads)
__device__ __forceinline__ void p2p_wait_flags(const unsigned* flags, int world, const int* epoch_ptr)
{
    if (flags != nullptr) {
        if (int(threadIdx.x) < world) {
            const unsigned epoch = unsigned(*epoch_ptr);
            unsigned v;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            } while (v < epoch);
        }
        __syncthreads();
    }
}

// programmatic dependent launch (PDL) controls; no-ops when the kernel was launched without the attribute
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace eetq_b200
