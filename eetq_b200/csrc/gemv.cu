// gemv.cu -- decode-time w8a16 streaming kernel for M <= 8 (sm_100a, SIMT, no tensor cores).
//
// Replaces the reference decode path
//   weight_only_batched_gemv<half, Int8b, PerChannel, ...>   /root/reference/csrc/weightOnlyBatchedGemv/kernel.h:294-468
//   (dispatch: weightOnlyBatchedGemv/kernelLauncher.cu:165-199, called from fpA_intB_gemm_wrapper.cu:149-160)
// with a design sized for B200's HBM3e rather than a translation of it:
//
//   * weights are in the b200 layout (row n = the K biased bytes u = q + 128 of output feature n, contiguous),
//     so a CTA streams whole rows with perfectly coalesced 128-bit loads; thread t of the CTA owns the 16-byte
//     K-chunks {t, t+256, ...} of EVERY row and keeps the matching activation slice in REGISTERS (packed fp16:
//     8 registers per 16 values) -- weights flow HBM -> registers -> FMA with no shared-memory hop and zero
//     activation re-reads;
//   * rows are processed R at a time, register double-buffered: 2 x R x KITERS 16-byte loads in flight per
//     thread (256 B/thread, ~128 KB/SM at 2 CTAs/SM), which is what Little's law needs to cover ~6.5 TB/s;
//   * fp16 inner loop = 1 PRMT + 2 FHFMA per 2 weights and NO int->float conversion: PRMT drops each byte under
//     the fp16 exponent byte 0x64, giving fp16(1024 + u) for free; FHFMA (fma.rn.f32.f16, new on sm_100)
//     multiplies it with the fp16 activation and accumulates in fp32 -- products are exact (11 x 11 bits) and the
//     constant (1024 + 128) * sum(x) is subtracted once per row.  The per-channel scale is applied ONCE per
//     output in the epilogue with the optional bias.  The reference multiplies every weight by the scale in fp16
//     and accumulates in fp16 per thread (kernel.h:355-377, :425-435), so this kernel is strictly more
//     accurate -- parity is checked against the fp32-accumulation oracle at 1e-3 norm-relative
//     (tests/test_gemm_gpu.py).  bf16 activations (extension) use an fp32 mantissa-trick + FFMA2 path;
//   * per-row partial sums are reduced with a transposing warp butterfly (9 shuffles per 8 rows instead of
//     40) and one shared-memory pass at the end of the CTA;
//   * the grid is a multiple of the SM count and rows are split evenly (+-1) over CTAs;
//   * the weight loads of the first TWO row groups (16 x 16 bytes per thread, ~19 MB chip-wide) are issued BEFORE
//     griddepcontrol.wait, so under programmatic dependent launch the HBM stream of kernel i+1 starts while
//     kernel i drains.
//
// Also in this file, both opt-in and measured slower than the kernel above: the TMA-bulk shared-
// memory-ring variant (w8a16_gemv_stream_kernel) and the chained multi-GEMV launch (w8a16_gemv_chain_kernel).
//
// Algorithmic bytes per call: K*N + 2*N + 2*M*K + 2*M*N; each weight byte is read exactly once.
#include <cstdlib>

#include "common.cuh"

namespace eetq_b200 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps   = kThreads / 32;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// acc(fp32) += a(fp16) * b(fp16): one FHFMA on sm_100a (exact 22-bit product, single fp32 rounding)
__device__ __forceinline__ float fhfma(uint16_t a, uint16_t b, float acc)
{
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
    return acc;
}

// ---- per-dtype inner product of one 16-byte weight chunk (16 biased bytes u = q + 128) with 16 activations ----
// fp16:  the bytes are PRMT-ed under the exponent byte 0x64 -> fp16(1024 + u) (no arithmetic), then FHFMA with
//        the packed fp16 activations accumulates x * (1024 + u) exactly in fp32; the constant (1024 + 128) * sum(x)
//        is removed once per row (XOffset below).  3 instructions per 2 weights, no int->float conversion at all.
// bf16:  bf16 cannot hold 1024 + u, so bytes go through the fp32 mantissa trick (2^23 + u) - (2^23 + 128) = q and
//        packed FFMA2 against fp32 activations.
template <typename T>
struct XSlice;

template <>
struct XSlice<__half> {
    uint32_t h2[8];  // 16 activations, packed fp16 pairs
    float sum;       // their fp32 sum
    __device__ __forceinline__ void load(const __half* p)
    {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
        h2[0] = a.x; h2[1] = a.y; h2[2] = a.z; h2[3] = a.w;
        h2[4] = b.x; h2[5] = b.y; h2[6] = b.z; h2[7] = b.w;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
            s += f.x + f.y;
        }
        sum = s;
    }
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            h2[j] = 0u;
        sum = 0.f;
    }
    __device__ __forceinline__ void resum()
    {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
            s += f.x + f.y;
        }
        sum = s;
    }
    __device__ __forceinline__ float sumsq() const
    {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
            s = fmaf(f.x, f.x, s);
            s = fmaf(f.y, f.y, s);
        }
        return s;
    }
    // RMSNorm in place, HF Llama arithmetic: fp16( fp16(x_f32 * r) * w )
    __device__ __forceinline__ void apply_norm(float r, const __half* nw)
    {
        const uint4 a = *reinterpret_cast<const uint4*>(nw);
        const uint4 b = *reinterpret_cast<const uint4*>(nw + 8);
        const uint32_t wr[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f  = __half22float2(*reinterpret_cast<const __half2*>(&h2[j]));
            const __half2 n = __floats2half2_rn(f.x * r, f.y * r);
            const __half2 o = __hmul2(n, *reinterpret_cast<const __half2*>(&wr[j]));
            h2[j]           = *reinterpret_cast<const uint32_t*>(&o);
        }
        resum();
    }
    // x = fp16( fp16(silu(g)) * u )  (HF: act_fn(gate_proj(x)) * up_proj(x) in fp16)
    __device__ __forceinline__ void load_silu_mul(const __half* g, const __half* u)
    {
        const uint4 ga = *reinterpret_cast<const uint4*>(g), gb = *reinterpret_cast<const uint4*>(g + 8);
        const uint4 ua = *reinterpret_cast<const uint4*>(u), ub = *reinterpret_cast<const uint4*>(u + 8);
        const uint32_t gr[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const uint32_t ur[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 f  = __half22float2(*reinterpret_cast<const __half2*>(&gr[j]));
            const __half2 a = __floats2half2_rn(f.x / (1.f + __expf(-f.x)), f.y / (1.f + __expf(-f.y)));
            const __half2 o = __hmul2(a, *reinterpret_cast<const __half2*>(&ur[j]));
            h2[j]           = *reinterpret_cast<const uint32_t*>(&o);
        }
        resum();
    }
    // acc += sum_j x_j * (1024 + u_j)
    __device__ __forceinline__ void dot(const uint4& wv, float& acc) const
    {
        const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t p01 = __byte_perm(words[q], 0x64646464u, 0x5140);
            const uint32_t p23 = __byte_perm(words[q], 0x64646464u, 0x5342);
            acc = fhfma(uint16_t(p01 & 0xffffu), uint16_t(h2[2 * q] & 0xffffu), acc);
            acc = fhfma(uint16_t(p01 >> 16), uint16_t(h2[2 * q] >> 16), acc);
            acc = fhfma(uint16_t(p23 & 0xffffu), uint16_t(h2[2 * q + 1] & 0xffffu), acc);
            acc = fhfma(uint16_t(p23 >> 16), uint16_t(h2[2 * q + 1] >> 16), acc);
        }
    }
    static constexpr float kOffset = 1152.f;  // 1024 (exponent trick) + 128 (storage bias)
};

template <>
struct XSlice<__nv_bfloat16> {
    float2 f2[8];
    float sum;  // unused (offset already removed per element)
    __device__ __forceinline__ void load(const __nv_bfloat16* p)
    {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
        const uint32_t raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
            f2[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[j]));
        sum = 0.f;
    }
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            f2[j] = make_float2(0.f, 0.f);
        sum = 0.f;
    }
    __device__ __forceinline__ float sumsq() const
    {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s = fmaf(f2[j].x, f2[j].x, s);
            s = fmaf(f2[j].y, f2[j].y, s);
        }
        return s;
    }
    static __device__ __forceinline__ float rb(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
    __device__ __forceinline__ void apply_norm(float r, const __nv_bfloat16* nw)
    {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            f2[j].x = rb(rb(f2[j].x * r) * __bfloat162float(nw[2 * j]));
            f2[j].y = rb(rb(f2[j].y * r) * __bfloat162float(nw[2 * j + 1]));
        }
    }
    __device__ __forceinline__ void load_silu_mul(const __nv_bfloat16* g, const __nv_bfloat16* u)
    {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float g0 = __bfloat162float(g[2 * j]), g1 = __bfloat162float(g[2 * j + 1]);
            f2[j].x = rb(rb(g0 / (1.f + __expf(-g0))) * __bfloat162float(u[2 * j]));
            f2[j].y = rb(rb(g1 / (1.f + __expf(-g1))) * __bfloat162float(u[2 * j + 1]));
        }
        sum = 0.f;
    }
    __device__ __forceinline__ void dot(const uint4& wv, float& acc) const
    {
        const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 lo, hi;
            lo.x = __uint_as_float(__byte_perm(words[q], 0x4B000000u, 0x7650)) - 8388736.f;
            lo.y = __uint_as_float(__byte_perm(words[q], 0x4B000000u, 0x7651)) - 8388736.f;
            hi.x = __uint_as_float(__byte_perm(words[q], 0x4B000000u, 0x7652)) - 8388736.f;
            hi.y = __uint_as_float(__byte_perm(words[q], 0x4B000000u, 0x7653)) - 8388736.f;
            a2   = ffma2(lo, f2[2 * q], a2);
            a2   = ffma2(hi, f2[2 * q + 1], a2);
        }
        acc += a2.x + a2.y;
    }
    static constexpr float kOffset = 0.f;
};

// Sum v[r] over the 32 lanes for all R rows with a transposing butterfly.  On return every lane holds,
// in v[0], the warp total of row (lane >> (5 - log2 R)).
template <int R>
__device__ __forceinline__ float warp_reduce_rows(float (&v)[R], int lane)
{
    int mask = 16;
#pragma unroll
    for (int width = R; width > 1; width >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
        const int h      = width >> 1;
#pragma unroll
        for (int j = 0; j < h; ++j) {
            const float keep = upper ? v[j + h] : v[j];
            const float send = upper ? v[j] : v[j + h];
            v[j]             = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
#pragma unroll
    for (; mask >= 1; mask >>= 1)
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
    return v[0];
}

template <int R>
struct Log2;
template <>
struct Log2<1> {
    static constexpr int v = 0;
};
template <>
struct Log2<2> {
    static constexpr int v = 1;
};
template <>
struct Log2<4> {
    static constexpr int v = 2;
};
template <>
struct Log2<8> {
    static constexpr int v = 3;
};

// dynamic smem: partial[row][m][warp] fp32
extern __shared__ float gemv_partial[];

// Kernel selection (development knobs, read once):
//   EETQ_B200_GEMV_IMPL = ldg (default: register double-buffered LDG kernel) | tma (cp.async.bulk ring kernel)
//   EETQ_B200_GEMV_PREFETCH = 1: issue only ONE row group of weight loads before the dependency wait (default 2)
int gemv_impl()
{
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("EETQ_B200_GEMV_IMPL");
        impl          = (e != nullptr && e[0] == 't') ? 0 : 1;
    }
    return impl;
}
int gemv_prefetch_groups()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("EETQ_B200_GEMV_PREFETCH");
        v             = (e != nullptr && e[0] == '1') ? 1 : 2;
    }
    return v;
}
constexpr __host__ __device__ int min_ctas(int M, int KITERS, bool XREG) { return (XREG && M * KITERS <= 4) ? 2 : 1; }

// Optional fusions around the GEMV (decode-side glue folded into the hot kernel; all pointers may be null):
//   xmode GEMV_X_RMSNORM : the activation is RMS-normalised on load (x is the residual stream, norm_weight [K])
//   xmode GEMV_X_SILU_MUL: the activation is silu(x[:, :K]) * x[:, K:2K]   (x is the fused gate|up output, ldx >= 2K)
//   residual             : y = fp16(acc * s [+ bias]) + residual   (fp16 add, like `hidden = residual + o_proj(..)`)
template <typename T>
struct GemvFuse {
    const T* norm_weight;
    const T* residual;
    int64_t ldr;
    float eps;
    int xmode;
    GemvP2P p2p;
    int prefetch_groups;
    const char* pf_k;
    const char* pf_v;
    const int* pf_pos;
    int pf_heads;
    int pf_max_ctx;
};

// L2 prefetch of the KV rows the following attention kernel will stream (128-byte lines, spread over the whole grid)
template <typename T>
__device__ __forceinline__ void prefetch_kv_l2(const GemvFuse<T>& f)
{
    if (f.pf_k == nullptr)
        return;
    const int pos            = *f.pf_pos;
    const int lines_per_head = pos * 2;  // 256 bytes per cached position
    const int total          = 2 * f.pf_heads * lines_per_head;
    const int64_t head_bytes = int64_t(f.pf_max_ctx) * 256;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tensor = i / (f.pf_heads * lines_per_head);
        const int rem    = i - tensor * (f.pf_heads * lines_per_head);
        const int head   = rem / lines_per_head;
        const int line   = rem - head * lines_per_head;
        const char* addr = (tensor ? f.pf_v : f.pf_k) + head * head_bytes + int64_t(line) * 128;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
    }
}

__device__ __forceinline__ void st_release_sys(unsigned long long addr, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// store one output element locally or, for a column-sharded linear, into every rank's buffer
template <typename T>
__device__ __forceinline__ void store_out(T* y, int64_t idx, T o, const GemvP2P& pp)
{
    if (pp.world > 1) {
#pragma unroll 1
        for (int r = 0; r < pp.world; ++r)
            reinterpret_cast<T*>(pp.peer_y[r])[idx] = o;
    }
    else {
        y[idx] = o;
    }
}

// after all stores of this CTA: the last CTA publishes this rank's flag on every peer.  Called by every thread of the CTA.
__device__ __forceinline__ void p2p_signal_and_wait(const GemvP2P& pp, int tid)
{
    if (pp.world <= 1)
        return;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned old = atomicAdd(pp.ticket, 1u);
        if (old == gridDim.x - 1) {
            *pp.ticket = 0;  // self-cleaning
            __threadfence_system();
            const unsigned epoch = unsigned(*pp.epoch);
            for (int r = 0; r < pp.world; ++r)
                st_release_sys(pp.peer_flag[r], epoch);
            // no wait here: the CONSUMER of this buffer polls the flags in its prologue (p2p_wait_flags), after it has
            // issued its own weight prefetch, so the NVLink latency overlaps the next kernel's launch and prefetch
        }
    }
}

template <typename T, int M, int KITERS, int R, bool XREG>
__global__ void __launch_bounds__(kThreads, min_ctas(M, KITERS, XREG))
    w8a16_gemv_kernel(const T* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ w, const T* __restrict__ scales,
                      const T* __restrict__ bias, T* __restrict__ y, int64_t ldy, int N, int K, const GemvFuse<T> fuse)
{
    __shared__ float red_smem[M][kWarps];
    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int warp    = tid >> 5;
    const int nchunks = K >> 4;

    // even split of the N rows over the grid
    const int row_begin = int((int64_t(blockIdx.x) * N) / gridDim.x);
    const int row_end   = int((int64_t(blockIdx.x + 1) * N) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + R - 1) / R;

    // let the next kernel in the stream start its own prologue (no-op without PDL)
    pdl_launch_dependents();

    // warp butterfly + per-warp partial to smem for the R rows of group g
    auto reduce_store = [&](float (&acc)[M][R], int g) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float tot = warp_reduce_rows<R>(acc[m], lane);
            const int rid   = lane >> (5 - Log2<R>::v);
            if ((lane & ((32 >> Log2<R>::v) - 1)) == 0)
                gemv_partial[((g * R + rid) * M + m) * kWarps + warp] = tot;
        }
    };

    if constexpr (XREG) {
        // ------------------------------------------------------------------ register-resident activations
        uint4 wb[2][R][KITERS];
        auto load_group = [&](uint4 (&buf)[R][KITERS], int g) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = row_begin + g * R + r;
#pragma unroll
                for (int i = 0; i < KITERS; ++i) {
                    const int c = tid + i * kThreads;
                    if (row < row_end && c < nchunks)
                        buf[r][i] = ldg_stream_128(w + int64_t(row) * K + int64_t(c) * 16);
                    else
                        buf[r][i] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        };

        // weights do not depend on the previous kernel: start streaming before the dependency wait
        if (ngroups > 0)
            load_group(wb[0], 0);
        if (ngroups > 1 && fuse.prefetch_groups >= 2)
            load_group(wb[1], 1);
        pdl_wait_prior_grids();
        p2p_wait_flags(fuse.p2p.wait_flags, fuse.p2p.world, fuse.p2p.epoch);  // gathered input: all ranks' slices present?

        XSlice<T> xs[M][KITERS];
        float xoff[M];
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kThreads;
                if (c >= nchunks)
                    xs[m][i].zero();
                else if (fuse.xmode == GEMV_X_SILU_MUL)
                    xs[m][i].load_silu_mul(x + int64_t(m) * ldx + int64_t(c) * 16, x + int64_t(m) * ldx + K + int64_t(c) * 16);
                else
                    xs[m][i].load(x + int64_t(m) * ldx + int64_t(c) * 16);
            }
        if (ngroups > 1 && fuse.prefetch_groups < 2)
            load_group(wb[1], 1);  // experiment: activation loads go out ahead of the second weight group
        if (fuse.xmode == GEMV_X_RMSNORM) {
            // every CTA holds the whole activation row across its threads: block-reduce sum(x^2), normalise in registers
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < KITERS; ++i)
                    ss += xs[m][i].sumsq();
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1)
                    ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0)
                    red_smem[m][warp] = ss;
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float tot = 0.f;
#pragma unroll
                for (int wi = 0; wi < kWarps; ++wi)
                    tot += red_smem[m][wi];
                const float r = rsqrtf(tot / float(K) + fuse.eps);
#pragma unroll
                for (int i = 0; i < KITERS; ++i) {
                    const int c = tid + i * kThreads;
                    if (c < nchunks)
                        xs[m][i].apply_norm(r, fuse.norm_weight + int64_t(c) * 16);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float so = 0.f;
#pragma unroll
            for (int i = 0; i < KITERS; ++i)
                so += xs[m][i].sum;
            xoff[m] = -XSlice<T>::kOffset * so;
        }

        prefetch_kv_l2<T>(fuse);  // hints only: no registers, no waiting

        auto compute_group = [&](uint4 (&buf)[R][KITERS], int g) {
            float acc[M][R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float a = xoff[m];
#pragma unroll
                    for (int i = 0; i < KITERS; ++i)
                        xs[m][i].dot(buf[r][i], a);
                    acc[m][r] = a;
                }
            }
            reduce_store(acc, g);
        };

        for (int g = 0; g < ngroups; g += 2) {
            compute_group(wb[0], g);
            if (g + 2 < ngroups)
                load_group(wb[0], g + 2);
            if (g + 1 < ngroups) {
                compute_group(wb[1], g + 1);
                if (g + 3 < ngroups)
                    load_group(wb[1], g + 3);
            }
        }
    }
    else {
        // ------------------------------------------------------------------ activations re-read through L1 (any M, any K)
        pdl_wait_prior_grids();
        p2p_wait_flags(fuse.p2p.wait_flags, fuse.p2p.world, fuse.p2p.epoch);
        const int kiters = (nchunks + kThreads - 1) / kThreads;
        for (int g = 0; g < ngroups; ++g) {
            float acc[M][R];
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int r = 0; r < R; ++r)
                    acc[m][r] = 0.f;
            for (int i = 0; i < kiters; ++i) {
                const int c = tid + i * kThreads;
                if (c >= nchunks)
                    break;
                uint4 buf[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int row = row_begin + g * R + r;
                    buf[r]        = (row < row_end) ? ldg_stream_128(w + int64_t(row) * K + int64_t(c) * 16)
                                                    : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    XSlice<T> xv;
                    xv.load(x + int64_t(m) * ldx + int64_t(c) * 16);
                    const float off = -XSlice<T>::kOffset * xv.sum;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float a = off;
                        xv.dot(buf[r], a);
                        acc[m][r] += a;
                    }
                }
            }
            reduce_store(acc, g);
        }
    }

    __syncthreads();
    // epilogue: cross-warp sum, per-channel scale (+bias), store
    for (int idx = tid; idx < nrows * M; idx += kThreads) {
        const int r = idx / M;
        const int m = idx - r * M;
        float s     = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi)
            s += gemv_partial[(r * M + m) * kWarps + wi];
        const int n = row_begin + r;
        float out   = s * to_float(scales[n]);
        if (bias != nullptr)
            out += to_float(bias[n]);
        T o = from_float<T>(out);
        if (fuse.residual != nullptr)
            o = from_float<T>(to_float(o) + to_float(fuse.residual[int64_t(m) * fuse.ldr + n]));
        store_out<T>(y, int64_t(m) * ldy + n, o, fuse.p2p);
    }
    p2p_signal_and_wait(fuse.p2p, tid);
}

// =====================================================================================================================
// Chained variant: up to 4 dependent decode GEMVs (o_proj -> gate|up -> down -> next layer's q|k|v) in ONE launch.
// Between phases the CTAs meet at a grid barrier (monotonic counter, target = epoch * gridDim.x), but BEFORE waiting each
// CTA has already issued the first two row groups of the next phase's weights (16 x 16-byte loads per thread, ~19 MB
// chip-wide), so HBM keeps streaming across what used to be three kernel boundaries per layer.  M = 1, fp16.
// =====================================================================================================================
struct GemvPhase {
    const __half* x;
    int64_t ldx;
    const uint8_t* w;
    const __half* scales;
    __half* y;
    int N;
    int K;
    const __half* norm_weight;
    const __half* residual;
    float eps;
    int xmode;
};
constexpr int kMaxChain = 4;
struct GemvChain {
    GemvPhase ph[kMaxChain];
    int nphases;
    unsigned* counters;   // one per phase boundary, never reset
    const int* epoch;     // device step counter (>= 1), strictly increasing per launch of THIS chain
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int KITERS, int R>
__device__ __forceinline__ void chain_phase(const GemvPhase& ph, bool first, const unsigned* wait_counter, unsigned target,
                                            float (*red_smem)[kWarps])
{
    using T = __half;
    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int warp    = tid >> 5;
    const int K       = ph.K;
    const int nchunks = K >> 4;
    const int row_begin = int((int64_t(blockIdx.x) * ph.N) / gridDim.x);
    const int row_end   = int((int64_t(blockIdx.x + 1) * ph.N) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + R - 1) / R;

    uint4 wb[2][R][KITERS];
    auto load_group = [&](uint4 (&buf)[R][KITERS], int g) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row_begin + g * R + r;
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kThreads;
                if (row < row_end && c < nchunks)
                    buf[r][i] = ldg_stream_128(ph.w + int64_t(row) * K + int64_t(c) * 16);
                else
                    buf[r][i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    // weights first (independent of the previous phase), then the dependency
    if (ngroups > 0)
        load_group(wb[0], 0);
    if (ngroups > 1)
        load_group(wb[1], 1);
    if (first) {
        pdl_wait_prior_grids();
    }
    else {
        if (tid == 0) {
            while (ld_acquire_gpu(wait_counter) < target) {
            }
        }
        __syncthreads();
    }

    XSlice<T> xs[KITERS];
#pragma unroll
    for (int i = 0; i < KITERS; ++i) {
        const int c = tid + i * kThreads;
        if (c >= nchunks)
            xs[i].zero();
        else if (ph.xmode == GEMV_X_SILU_MUL)
            xs[i].load_silu_mul(ph.x + int64_t(c) * 16, ph.x + K + int64_t(c) * 16);
        else
            xs[i].load(ph.x + int64_t(c) * 16);
    }
    if (ph.xmode == GEMV_X_RMSNORM) {
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < KITERS; ++i)
            ss += xs[i].sumsq();
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0)
            red_smem[0][warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi)
            tot += red_smem[0][wi];
        const float r = rsqrtf(tot / float(K) + ph.eps);
#pragma unroll
        for (int i = 0; i < KITERS; ++i) {
            const int c = tid + i * kThreads;
            if (c < nchunks)
                xs[i].apply_norm(r, ph.norm_weight + int64_t(c) * 16);
        }
    }
    float so = 0.f;
#pragma unroll
    for (int i = 0; i < KITERS; ++i)
        so += xs[i].sum;
    const float xoff = -XSlice<T>::kOffset * so;

    auto compute_group = [&](uint4 (&buf)[R][KITERS], int g) {
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float a = xoff;
#pragma unroll
            for (int i = 0; i < KITERS; ++i)
                xs[i].dot(buf[r][i], a);
            acc[r] = a;
        }
        const float tot = warp_reduce_rows<R>(acc, lane);
        const int rid   = lane >> (5 - Log2<R>::v);
        if ((lane & ((32 >> Log2<R>::v) - 1)) == 0)
            gemv_partial[(g * R + rid) * kWarps + warp] = tot;
    };
    for (int g = 0; g < ngroups; g += 2) {
        compute_group(wb[0], g);
        if (g + 2 < ngroups)
            load_group(wb[0], g + 2);
        if (g + 1 < ngroups) {
            compute_group(wb[1], g + 1);
            if (g + 3 < ngroups)
                load_group(wb[1], g + 3);
        }
    }
    __syncthreads();
    for (int r = tid; r < nrows; r += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi)
            s += gemv_partial[r * kWarps + wi];
        const int n = row_begin + r;
        __half o    = __float2half_rn(s * __half2float(ph.scales[n]));
        if (ph.residual != nullptr)
            o = __float2half_rn(__half2float(o) + __half2float(ph.residual[n]));
        ph.y[n] = o;
    }
}

template <int KI>
struct ChainR {
    static constexpr int v = KI == 1 ? 8 : (KI == 2 ? 4 : 2);
};

// KI0..KI3: K-chunk iterations (ceil(K/4096)) of each phase, 0 = phase absent.  Fully unrolled so every phase indexes the
// kernel-parameter struct statically (no local-memory copy) and gets its own register allocation.
template <int KI0, int KI1, int KI2, int KI3>
__global__ void __launch_bounds__(kThreads, 2) w8a16_gemv_chain_kernel(const __grid_constant__ GemvChain c)
{
    __shared__ float red_smem[1][kWarps];
    pdl_launch_dependents();
    const unsigned target = unsigned(*c.epoch) * gridDim.x;
    auto arrive = [&](int p) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(c.counters + p, 1u);
        }
    };
    chain_phase<KI0, ChainR<KI0>::v>(c.ph[0], true, c.counters, target, red_smem);
    if constexpr (KI1 > 0) {
        arrive(0);
        chain_phase<KI1, ChainR<KI1>::v>(c.ph[1], false, c.counters + 0, target, red_smem);
    }
    if constexpr (KI2 > 0) {
        arrive(1);
        chain_phase<KI2, ChainR<KI2>::v>(c.ph[2], false, c.counters + 1, target, red_smem);
    }
    if constexpr (KI3 > 0) {
        arrive(2);
        chain_phase<KI3, ChainR<KI3>::v>(c.ph[3], false, c.counters + 2, target, red_smem);
    }
}

// =====================================================================================================================
// TMA-streamed variant (default): the CTA's rows are ONE contiguous byte range of the b200 layout, so a single
// producer thread streams it with cp.async.bulk (TMA 1-D) into a shared-memory ring (~96 KB/CTA, one CTA per SM) and
// runs AHEAD of the consumers -- across stages and, under programmatic dependent launch, across kernels: the next
// GEMV's CTA co-resides (low register / half the smem) and fills its ring while this kernel is still streaming,
// so HBM never idles at kernel boundaries.  The 8 consumer warps read the ring with conflict-free LDS.128 (thread t
// owns 16-byte K-chunk t of every row, activation slice in registers) and do the same PRMT + FHFMA arithmetic.
// =====================================================================================================================
constexpr int kConsumers      = 256;              // 8 consumer warps
constexpr int kStreamThreads  = kConsumers + 32;  // + 1 producer warp
constexpr int kMaxStages      = 16;
constexpr int kRingBytes      = 96 * 1024;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GEMV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GEMV_DONE;\n"
        "bra GEMV_WAIT;\n"
        "GEMV_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion on an mbarrier, L2 evict-first (weights are read once per token)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

// dynamic smem layout: [ring: stages * stage_bytes][partial: padded_rows * M * 8 floats]
template <typename T, int M, int KITERS, int R>
__global__ void __launch_bounds__(kStreamThreads, 1)
    w8a16_gemv_stream_kernel(const T* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ w, const T* __restrict__ scales,
                             const T* __restrict__ bias, T* __restrict__ y, int64_t ldy, int N, int K, int stages, int stage_bytes,
                             const GemvFuse<T> fuse)
{
    extern __shared__ __align__(128) uint8_t stream_smem[];
    __shared__ __align__(8) uint64_t bars[2 * kMaxStages];
    __shared__ float red_smem[M][kConsumers / 32];

    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int warp    = tid >> 5;
    const int nchunks = K >> 4;

    const int row_begin = int((int64_t(blockIdx.x) * N) / gridDim.x);
    const int row_end   = int((int64_t(blockIdx.x + 1) * N) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + R - 1) / R;  // one ring stage per group of R rows

    const uint32_t full0  = smem_addr(&bars[0]);
    const uint32_t empty0 = smem_addr(&bars[kMaxStages]);
    float* partial        = reinterpret_cast<float*>(stream_smem + size_t(stages) * stage_bytes);

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, kConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();  // the next kernel may start (and prefetch its own weights) as soon as it fits

    if (warp == kConsumers / 32) {
        // ================================================================= producer: runs ahead, independent of x
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            const uint8_t* src   = w + int64_t(row_begin) * K;
            const uint32_t ring0 = smem_addr(stream_smem);
            for (int g = 0; g < ngroups; ++g) {
                const int s       = g % stages;
                const uint32_t ph = (g / stages) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                const int rows       = min(R, nrows - g * R);
                const uint32_t bytes = uint32_t(rows) * uint32_t(K);
                mbar_expect_tx(full0 + 8 * s, bytes);
                // rows are adjacent in memory: the whole group is one contiguous range; issue it in <= 16 KB pieces
                uint32_t done = 0;
                while (done < bytes) {
                    const uint32_t piece = min(bytes - done, 16384u);
                    bulk_g2s(ring0 + uint32_t(s) * uint32_t(stage_bytes) + done, src + int64_t(g) * R * K + done, piece,
                             full0 + 8 * s, policy);
                    done += piece;
                }
            }
        }
        return;
    }

    // ===================================================================== consumers
    pdl_wait_prior_grids();  // activations / residual come from the previous kernel

    XSlice<T> xs[M][KITERS];
    float xoff[M];
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int i = 0; i < KITERS; ++i) {
            const int c = tid + i * kConsumers;
            if (c >= nchunks)
                xs[m][i].zero();
            else if (fuse.xmode == GEMV_X_SILU_MUL)
                xs[m][i].load_silu_mul(x + int64_t(m) * ldx + int64_t(c) * 16, x + int64_t(m) * ldx + K + int64_t(c) * 16);
            else
                xs[m][i].load(x + int64_t(m) * ldx + int64_t(c) * 16);
        }
    if (fuse.xmode == GEMV_X_RMSNORM) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < KITERS; ++i)
                ss += xs[m][i].sumsq();
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1)
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0)
                red_smem[m][warp] = ss;
        }
        consumer_sync();
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float tot = 0.f;
#pragma unroll
            for (int wi = 0; wi < kConsumers / 32; ++wi)
                tot += red_smem[m][wi];
            const float r = rsqrtf(tot / float(K) + fuse.eps);
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kConsumers;
                if (c < nchunks)
                    xs[m][i].apply_norm(r, fuse.norm_weight + int64_t(c) * 16);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float so = 0.f;
#pragma unroll
        for (int i = 0; i < KITERS; ++i)
            so += xs[m][i].sum;
        xoff[m] = -XSlice<T>::kOffset * so;
    }

    for (int g = 0; g < ngroups; ++g) {
        const int s       = g % stages;
        const uint32_t ph = (g / stages) & 1;
        mbar_wait(full0 + 8 * s, ph);
        const uint8_t* stage = stream_smem + size_t(s) * stage_bytes;
        uint4 wv[R][KITERS];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kConsumers;
                // rows beyond the group (ragged last group) hold stale bytes: their results are never stored
                wv[r][i] = (c < nchunks) ? *reinterpret_cast<const uint4*>(stage + size_t(r) * K + size_t(c) * 16)
                                         : make_uint4(0u, 0u, 0u, 0u);
            }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(empty0 + 8 * s);  // data is in registers: hand the stage back to the producer
        float acc[M][R];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float a = xoff[m];
#pragma unroll
                for (int i = 0; i < KITERS; ++i)
                    xs[m][i].dot(wv[r][i], a);
                acc[m][r] = a;
            }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const float tot = warp_reduce_rows<R>(acc[m], lane);
            const int rid   = lane >> (5 - Log2<R>::v);
            if ((lane & ((32 >> Log2<R>::v) - 1)) == 0)
                partial[((g * R + rid) * M + m) * (kConsumers / 32) + warp] = tot;
        }
    }

    consumer_sync();
    for (int idx = tid; idx < nrows * M; idx += kConsumers) {
        const int r = idx / M;
        const int m = idx - r * M;
        float sum   = 0.f;
#pragma unroll
        for (int wi = 0; wi < kConsumers / 32; ++wi)
            sum += partial[(r * M + m) * (kConsumers / 32) + wi];
        const int n = row_begin + r;
        float out   = sum * to_float(scales[n]);
        if (bias != nullptr)
            out += to_float(bias[n]);
        T o = from_float<T>(out);
        if (fuse.residual != nullptr)
            o = from_float<T>(to_float(o) + to_float(fuse.residual[int64_t(m) * fuse.ldr + n]));
        y[int64_t(m) * ldy + n] = o;
    }
}

template <typename T, int M, int KITERS, int R>
int launch_stream(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N, int K,
                  const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv: device query failed");
        return EETQ_B200_ECUDA;
    }
    auto kernel           = w8a16_gemv_stream_kernel<T, M, KITERS, R>;
    const int stage_bytes = R * K;
    const int ctas        = 1;
    int stages            = (kRingBytes / ctas) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    int grid = di.sm_count * ctas;
    if (grid > N) grid = N;
    const int max_rows = (N + grid - 1) / grid;
    if (stages > (max_rows + R - 1) / R) stages = (max_rows + R - 1) / R;  // never more stages than row groups
    if (stages < 1) stages = 1;
    const int padded  = ((max_rows + R - 1) / R) * R;
    const size_t smem = size_t(stages) * stage_bytes + size_t(padded) * M * (kConsumers / 32) * sizeof(float);
    if (smem > size_t(di.max_smem_optin) - 2048) {
        set_error("gemv(stream): %zu bytes of shared memory needed (N=%d K=%d)", smem, N, K);
        return EETQ_B200_EINVAL;
    }
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        EB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin - 2048));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3(unsigned(grid));
    cfg.blockDim         = dim3(kStreamThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, x, ldx, w, scales, bias, y, ldy, N, K, stages, stage_bytes, fuse);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv(stream) launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T, int M, int KITERS, int R, bool XREG>
int launch_variant(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N,
                   int K, const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv: device query failed");
        return EETQ_B200_ECUDA;
    }
    constexpr int kCtasPerSm = min_ctas(M, KITERS, XREG);
    constexpr int kMaxRows   = 96;  // rows per CTA bound (sizes the partial-sum buffer)
    int grid                 = di.sm_count * kCtasPerSm;
    // keep rows/CTA <= kMaxRows, and grid a multiple of the SM count
    while ((N + grid - 1) / grid > kMaxRows)
        grid += di.sm_count;
    if (grid > N)
        grid = N;
    const int max_rows = (N + grid - 1) / grid;
    const int padded   = ((max_rows + R - 1) / R) * R;  // the partial buffer is indexed by padded group rows
    const size_t smem  = size_t(padded) * M * kWarps * sizeof(float);

    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3(unsigned(grid));
    cfg.blockDim         = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = pdl ? 1 : 0;

    const cudaError_t e =
        cudaLaunchKernelEx(&cfg, w8a16_gemv_kernel<T, M, KITERS, R, XREG>, x, ldx, w, scales, bias, y, ldy, N, K, fuse);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T, int M>
int dispatch_k(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N, int K,
               const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    const int nchunks = K / 16;
    const int kiters  = (nchunks + kThreads - 1) / kThreads;
    // register-resident activations while the slice stays small (fp16: 8 regs, bf16: 16 regs per 16 values)
    constexpr int kMaxXregIters = (DTypeOf<T>::value == EETQ_B200_F16) ? 8 / M : 4 / M;
    // default: TMA-streamed kernel (rows per ring stage: 4 / 2 / 1 / 1 for K <= 4096 / 8192 / 12288 / 16384)
    if (gemv_impl() == 0 && fuse.p2p.world <= 1) {
#define EB_STREAM_CASE(KI, RR)                                                                                          \
    if (kiters == KI) {                                                                                                 \
        if constexpr (M * KI <= 4 && KI <= kMaxXregIters)                                                               \
            return launch_stream<T, M, KI, RR>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);               \
    }
        EB_STREAM_CASE(1, 4)
        EB_STREAM_CASE(2, 2)
        EB_STREAM_CASE(3, 1)
        EB_STREAM_CASE(4, 1)
#undef EB_STREAM_CASE
    }
#define EB_GEMV_CASE(KI, RS, RB)                                                                                        \
    if (kiters == KI) {                                                                                                 \
        if constexpr (KI <= kMaxXregIters)                                                                             \
            return launch_variant<T, M, KI, (M <= 2 ? RS : RB), true>(x, ldx, w, scales, bias, y, ldy, N, K, fuse,     \
                                                                      pdl, stream);                                     \
    }
    EB_GEMV_CASE(1, 8, 4)
    EB_GEMV_CASE(2, 4, 2)
    EB_GEMV_CASE(3, 2, 1)
    EB_GEMV_CASE(4, 2, 1)
#undef EB_GEMV_CASE
    // general path: activations re-read through L1 (no fused prologue there)
    if (fuse.xmode != GEMV_X_PLAIN) {
        set_error("gemv: fused RMSNorm / SiLU-mul prologue needs M * ceil(K/4096) <= %d (got M=%d, K=%d)", M * kMaxXregIters, M, K);
        return EETQ_B200_EINVAL;
    }
    constexpr int R = (M <= 2) ? 8 : 4;
    return launch_variant<T, M, 1, R, false>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
}

template <typename T>
int dispatch_m(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int M, int N,
               int K, const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    switch (M) {
        case 1: return dispatch_k<T, 1>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 2: return dispatch_k<T, 2>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 3: return dispatch_k<T, 3>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 4: return dispatch_k<T, 4>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 5: return dispatch_k<T, 5>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 6: return dispatch_k<T, 6>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 7: return dispatch_k<T, 7>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 8: return dispatch_k<T, 8>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        default: set_error("gemv: M=%d out of range [1,%d]", M, EETQ_B200_GEMV_MAX_M); return EETQ_B200_EINVAL;
    }
}

int launch_gemv_chain_impl(const GemvChain& chain, int max_rows_hint, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv_chain: device query failed");
        return EETQ_B200_ECUDA;
    }
    const int grid = di.sm_count * 2;
    int max_rows   = 0;
    for (int p = 0; p < chain.nphases; ++p) {
        const int rows = (chain.ph[p].N + grid - 1) / grid;
        if (rows > max_rows) max_rows = rows;
        if (chain.ph[p].N < grid) {
            set_error("gemv_chain: phase %d (N=%d) smaller than the grid", p, chain.ph[p].N);
            return EETQ_B200_EINVAL;
        }
    }
    int ki[kMaxChain] = {0, 0, 0, 0};
    for (int p = 0; p < chain.nphases; ++p)
        ki[p] = ((chain.ph[p].K >> 4) + kThreads - 1) / kThreads;
    // instantiated shapes: Llama-7B-like (hidden <= 4096, 8192 < inter <= 12288): o -> gate|up -> down [-> q|k|v]
    void (*kernel)(const GemvChain) = nullptr;
    if (ki[0] == 1 && ki[1] == 1 && ki[2] == 3 && ki[3] == 1)
        kernel = w8a16_gemv_chain_kernel<1, 1, 3, 1>;
    else if (ki[0] == 1 && ki[1] == 1 && ki[2] == 3 && ki[3] == 0)
        kernel = w8a16_gemv_chain_kernel<1, 1, 3, 0>;
    else if (ki[0] == 1 && ki[1] == 1 && ki[2] == 1 && ki[3] == 1)
        kernel = w8a16_gemv_chain_kernel<1, 1, 1, 1>;
    else if (ki[0] == 1 && ki[1] == 1 && ki[2] == 1 && ki[3] == 0)
        kernel = w8a16_gemv_chain_kernel<1, 1, 1, 0>;
    if (kernel == nullptr) {
        set_error("gemv_chain: K-iteration pattern (%d,%d,%d,%d) is not instantiated", ki[0], ki[1], ki[2], ki[3]);
        return EETQ_B200_EINVAL;
    }
    (void)max_rows_hint;
    const size_t smem = size_t(((max_rows + 7) / 8) * 8) * kWarps * sizeof(float);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim          = dim3(unsigned(grid));
    cfg.blockDim         = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, chain);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv_chain launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

}  // namespace

int launch_gemv(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                int M, int64_t N, int64_t K, int dtype, const GemvExtras& ex, bool pdl, cudaStream_t stream)
{
    const uint8_t* wu = reinterpret_cast<const uint8_t*>(w);
    if (dtype == EETQ_B200_F16) {
        using T = __half;
        GemvFuse<T> f{static_cast<const T*>(ex.norm_weight), static_cast<const T*>(ex.residual), ex.ldr, ex.eps, ex.xmode, ex.p2p, gemv_prefetch_groups(), static_cast<const char*>(ex.pf_k), static_cast<const char*>(ex.pf_v), ex.pf_pos, ex.pf_heads, ex.pf_max_ctx};
        return dispatch_m<T>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                             static_cast<T*>(y), ldy, M, int(N), int(K), f, pdl, stream);
    }
    if (dtype == EETQ_B200_BF16) {
        using T = __nv_bfloat16;
        GemvFuse<T> f{static_cast<const T*>(ex.norm_weight), static_cast<const T*>(ex.residual), ex.ldr, ex.eps, ex.xmode, ex.p2p, gemv_prefetch_groups(), static_cast<const char*>(ex.pf_k), static_cast<const char*>(ex.pf_v), ex.pf_pos, ex.pf_heads, ex.pf_max_ctx};
        return dispatch_m<T>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                             static_cast<T*>(y), ldy, M, int(N), int(K), f, pdl, stream);
    }
    set_error("gemv: unsupported activation dtype %d", dtype);
    return EETQ_B200_EINVAL;
}

int launch_gemv_chain(const GemvChainPhase* phases, int nphases, unsigned* counters, const int* epoch, bool pdl, cudaStream_t stream)
{
    if (nphases < 1 || nphases > kMaxChain) {
        set_error("gemv_chain: nphases must be in [1, %d]", kMaxChain);
        return EETQ_B200_EINVAL;
    }
    GemvChain c{};
    for (int p = 0; p < nphases; ++p) {
        const GemvChainPhase& s = phases[p];
        c.ph[p] = GemvPhase{static_cast<const __half*>(s.x), s.ldx, reinterpret_cast<const uint8_t*>(s.w),
                            static_cast<const __half*>(s.scales), static_cast<__half*>(s.y), int(s.N), int(s.K),
                            static_cast<const __half*>(s.norm_weight), static_cast<const __half*>(s.residual), s.eps, s.xmode};
    }
    c.nphases  = nphases;
    c.counters = counters;
    c.epoch    = epoch;
    return launch_gemv_chain_impl(c, 0, pdl, stream);
}

}  // namespace eetq_b200
