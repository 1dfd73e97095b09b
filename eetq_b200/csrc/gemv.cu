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
// Fusions (all optional): RMSNorm or nothing on the activation load, residual add, SiLU(gate) * up over interleaved
// (gate, up) row pairs in the epilogue, and -- for column-sharded multi-GPU decode -- activations read from / outputs pushed to
// "LL" exchange buffers ({2 x fp16, 32-bit tag} words written straight into every peer's memory over NVLink, see common.cuh).
// The TMA-ring and chained-launch variants of round 1 (measured slower) were removed.
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
    // 16 activations from 8 LL words ({2 x fp16, tag}); spins until all eight carry `tag`
    __device__ __forceinline__ void load_ll(const unsigned long long* words, uint32_t tag)
    {
        ll_load_words<8>(words, tag, h2);
        resum();
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
    __device__ __forceinline__ void apply_norm(float r, const uint4& a, const uint4& b)
    {
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
    // int4 (b200 int4 layout): a 32-bit word holds 8 consecutive k; nibble p < 4 is k = 2p, nibble 4 + p is k = 2p + 1, so
    // (word >> 4p) & 0x000f000f is the adjacent-k pair and OR-ing the exponent pattern gives fp16x2(1024 + u) with u = q + 8:
    // one shift + one LOP3 per 2 weights, then the same FHFMA pair
    __device__ __forceinline__ void dot(const uint2& wv, float& acc) const
    {
        const uint32_t words[2] = {wv.x, wv.y};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const uint32_t pr = ((words[q] >> (4 * p)) & 0x000f000fu) | 0x64006400u;
                acc = fhfma(uint16_t(pr & 0xffffu), uint16_t(h2[4 * q + p] & 0xffffu), acc);
                acc = fhfma(uint16_t(pr >> 16), uint16_t(h2[4 * q + p] >> 16), acc);
            }
        }
    }
    // 1024 (exponent trick) + storage bias (128 for int8 bytes, 8 for int4 nibbles)
    static constexpr __host__ __device__ float offset(int wbits) { return wbits == 8 ? 1152.f : 1032.f; }
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
    __device__ __forceinline__ void apply_norm(float r, const uint4& a, const uint4& b)
    {
        const uint32_t wr[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const __nv_bfloat16* nw = reinterpret_cast<const __nv_bfloat16*>(wr);
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
    // int4: nibbles through the same fp32 mantissa trick, (2^23 + u) - (2^23 + 8) = q
    __device__ __forceinline__ void dot(const uint2& wv, float& acc) const
    {
        const uint32_t words[2] = {wv.x, wv.y};
        float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float2 v;
                v.x = __uint_as_float(((words[q] >> (4 * p)) & 0xfu) | 0x4B000000u) - 8388616.f;
                v.y = __uint_as_float(((words[q] >> (4 * p + 16)) & 0xfu) | 0x4B000000u) - 8388616.f;
                a2  = ffma2(v, f2[4 * q + p], a2);
            }
        }
        acc += a2.x + a2.y;
    }
    static constexpr __host__ __device__ float offset(int) { return 0.f; }
};

// what a thread loads per (row, k-iteration): 16 consecutive k of one output feature
template <int WB>
struct WChunk;
template <>
struct WChunk<8> {
    using type = uint4;
    // byte address of 16-k chunk c of row `row`
    static __device__ __forceinline__ const uint8_t* at(const uint8_t* w, int row, int K, int c) { return w + int64_t(row) * K + int64_t(c) * 16; }
    static __device__ __forceinline__ uint4 load(const uint8_t* p) { return ldg_stream_128(p); }
    static __device__ __forceinline__ uint4 zero() { return make_uint4(0u, 0u, 0u, 0u); }
};
template <>
struct WChunk<4> {
    using type = uint2;
    static __device__ __forceinline__ const uint8_t* at(const uint8_t* w, int row, int K, int c) { return w + int64_t(row) * (K >> 1) + int64_t(c) * 8; }
    static __device__ __forceinline__ uint2 load(const uint8_t* p)
    {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return r;
    }
    static __device__ __forceinline__ uint2 zero() { return make_uint2(0u, 0u); }
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

constexpr __host__ __device__ int min_ctas(int M, int KITERS, bool XREG) { return (XREG && M * KITERS <= 4) ? 2 : 1; }

// Optional fusions around the GEMV (decode-side glue folded into the hot kernel; all pointers may be null):
//   xmode GEMV_X_RMSNORM : the activation is RMS-normalised on load (x is the residual stream, norm_weight [K])
//   xmode GEMV_X_SILU_MUL: the activation is silu(x[:, :K]) * x[:, K:2K]   (x is a gate|up vector, ldx >= 2K)
//   residual             : y = fp16(acc * s [+ bias]) + residual   (fp16 add, like `hidden = residual + o_proj(..)`)
//   epi GEMV_EPI_SILU_PAIRS: weight rows are interleaved (gate_0, up_0, gate_1, up_1, ...); the kernel emits N/2 values
//                          fp16(silu(fp16 gate)) * fp16 up  (HF `act_fn(gate_proj(x)) * up_proj(x)` arithmetic)
//   x_ll / res_ll / push : LL exchange buffers instead of plain vectors (M = 1, fp16)
template <typename T>
struct GemvFuse {
    const T* norm_weight;
    const T* residual;
    int64_t ldr;
    float eps;
    int xmode;
    int epi;
    LLTag x_ll;     // tag_base != nullptr: x points at LL words of the full K-vector
    LLTag res_ll;   // tag_base != nullptr: residual points at LL words of the full output vector; element offset res_off
    int res_off;
    LLPush push;    // world > 0: outputs are pushed as LL words to every rank instead of being stored to y
    NextHint next;  // w != nullptr: ask L2 for the head of the next GEMV's weights
    int nowait;     // != 0: every input is an LL buffer (data carries its own tag): do not wait for the previous grid to complete
};

template <typename T, int M, int KITERS, int R, bool XREG, int WB>
__global__ void __launch_bounds__(kThreads, min_ctas(M, KITERS, XREG))
    w8a16_gemv_kernel(const T* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ w, const T* __restrict__ scales,
                      const T* __restrict__ bias, T* __restrict__ y, int64_t ldy, int N, int K, const GemvFuse<T> fuse)
{
    using WV = typename WChunk<WB>::type;  // 16 k-values of one row: 16 bytes (int8) or 8 bytes (int4)
    __shared__ float red_smem[M][kWarps];
    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int warp    = tid >> 5;
    const int nchunks = K >> 4;

    // even split of the N rows over the grid, in units of `align` rows (row pairs for the SiLU*up epilogue, output pairs
    // for LL words)
    const bool pairs    = fuse.epi == GEMV_EPI_SILU_PAIRS;
    const bool ll_out   = fuse.push.world > 0;
    const int align     = (pairs ? 2 : 1) * (ll_out ? 2 : 1);
    const int units     = N / align;
    const int row_begin = align * int((int64_t(blockIdx.x) * units) / gridDim.x);
    const int row_end   = align * int((int64_t(blockIdx.x + 1) * units) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + R - 1) / R;

    const int nelem      = pairs ? nrows / 2 : nrows;        // output elements produced by this CTA
    const int elem_begin = pairs ? row_begin / 2 : row_begin;
    float pre_s0 = 0.f, pre_s1 = 0.f, pre_bias = 0.f, pre_res = 0.f;  // M == 1, register-resident path: fetched ahead of the main loop
    (void)pre_s1; (void)pre_bias; (void)pre_res;

    trace_ev(TRACE_GEMV, 0);
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
        WV wb[2][R][KITERS];
        auto load_group = [&](WV (&buf)[R][KITERS], int g) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = row_begin + g * R + r;
#pragma unroll
                for (int i = 0; i < KITERS; ++i) {
                    const int c = tid + i * kThreads;
                    if (row < row_end && c < nchunks)
                        buf[r][i] = WChunk<WB>::load(WChunk<WB>::at(w, row, K, c));
                    else
                        buf[r][i] = WChunk<WB>::zero();
                }
            }
        };

        // weights do not depend on the previous kernel: start streaming before the dependency wait
        if (ngroups > 0)
            load_group(wb[0], 0);
        if (ngroups > 1)
            load_group(wb[1], 1);
        // L2 staging for the NEXT GEMV (see GemvExtras::next_w).  Measured alternatives, all slower than no hint: staging the rest of
        // this CTA's OWN slice (10.9 vs 9.8 us per launch), staging a few rows for EVERY CTA of the next kernel (571 vs 589 tok/s),
        // requesting when the CTA leaves its main loop instead of here (548 vs 583 tok/s).
        if (fuse.next.w != nullptr && tid == kThreads - 32)
            l2_prefetch_next(fuse.next, blockIdx.x, gridDim.x);
        // so do all other parameters: the scale(s) / bias of the output element this thread will finish (M == 1: element `tid`),
        // and the RMSNorm weights of its activation chunks -- nothing static is left to fetch on the dependent path
        if constexpr (M == 1) {
            if (tid < nelem) {
                if (pairs) {
                    pre_s0 = to_float(scales[row_begin + 2 * tid]);
                    pre_s1 = to_float(scales[row_begin + 2 * tid + 1]);
                }
                else {
                    pre_s0 = to_float(scales[row_begin + tid]);
                    if (bias != nullptr)
                        pre_bias = to_float(bias[row_begin + tid]);
                }
            }
        }
        uint4 nw[KITERS][2];
        if (fuse.xmode == GEMV_X_RMSNORM) {
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kThreads;
                if (c < nchunks) {
                    nw[i][0] = *reinterpret_cast<const uint4*>(fuse.norm_weight + int64_t(c) * 16);
                    nw[i][1] = *reinterpret_cast<const uint4*>(fuse.norm_weight + int64_t(c) * 16 + 8);
                }
            }
        }
        trace_ev(TRACE_GEMV, 1);
        // Plain inputs become valid when the previous grid has COMPLETED (griddepcontrol.wait returns ~1.4 us after its last CTA);
        // LL inputs carry their own tags, polled word by word below, so the chain of decode kernels need not drain between stages.
        if (fuse.nowait == 0)
            pdl_wait_prior_grids();
        trace_ev(TRACE_GEMV, 2);
        // the residual of this thread's output element travels together with the activations
        if constexpr (M == 1) {
            if (fuse.residual != nullptr && fuse.res_ll.tag_base == nullptr && tid < nelem)
                pre_res = to_float(fuse.residual[elem_begin + tid]);
        }

        XSlice<T> xs[M][KITERS];
        float xoff[M];
        if constexpr (M == 1 && DTypeOf<T>::value == EETQ_B200_F16) {
            if (fuse.x_ll.tag_base != nullptr) {
                // gathered input: poll the LL words of this thread's chunks until every rank's slice has landed
                const uint32_t tag = ll_tag(fuse.x_ll);
#pragma unroll
                for (int i = 0; i < KITERS; ++i) {
                    const int c = tid + i * kThreads;
                    if (c >= nchunks)
                        xs[0][i].zero();
                    else
                        xs[0][i].load_ll(reinterpret_cast<const unsigned long long*>(x) + int64_t(c) * 8, tag);
                }
            }
        }
        if (!(M == 1 && DTypeOf<T>::value == EETQ_B200_F16 && fuse.x_ll.tag_base != nullptr)) {
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
        }
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
                        xs[m][i].apply_norm(r, nw[i][0], nw[i][1]);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float so = 0.f;
#pragma unroll
            for (int i = 0; i < KITERS; ++i)
                so += xs[m][i].sum;
            xoff[m] = -XSlice<T>::offset(WB) * so;
        }

        trace_ev(TRACE_GEMV, 3);
        auto compute_group = [&](WV (&buf)[R][KITERS], int g) {
            float acc[M][R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                // int4 rows cost 30 instructions per 16 weights and the kernel is issue-bound there (ncu: 60 % issue-active with 4 warps
                // per scheduler): skip the zero-filled rows of the last group (warp-uniform).  The int8 code is left exactly as measured.
                const bool live = (WB == 8) || (row_begin + g * R + r < row_end);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float a = xoff[m];
                    if (live) {
#pragma unroll
                        for (int i = 0; i < KITERS; ++i)
                            xs[m][i].dot(buf[r][i], a);
                    }
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
                WV buf[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int row = row_begin + g * R + r;
                    buf[r]        = (row < row_end) ? WChunk<WB>::load(WChunk<WB>::at(w, row, K, c)) : WChunk<WB>::zero();
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    XSlice<T> xv;
                    xv.load(x + int64_t(m) * ldx + int64_t(c) * 16);
                    const float off = -XSlice<T>::offset(WB) * xv.sum;
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

    trace_ev(TRACE_GEMV, 4);
    __syncthreads();
    // epilogue: cross-warp sum, per-channel scale (+bias), optional SiLU(gate)*up over row pairs, optional residual, then a
    // plain store or an LL push to every rank
    auto row_value = [&](int r, int m) -> float {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi)
            s += gemv_partial[(r * M + m) * kWarps + wi];
        return s * to_float(scales[row_begin + r]);
    };
    auto out_value = [&](int e, int m) -> T {
        T o;
        if (pairs) {
            const T g16 = from_float<T>(row_value(2 * e, m));
            const T u16 = from_float<T>(row_value(2 * e + 1, m));
            const float gf = to_float(g16);
            o = from_float<T>(to_float(from_float<T>(gf / (1.f + __expf(-gf)))) * to_float(u16));
        }
        else {
            float out = row_value(e, m);
            if (bias != nullptr)
                out += to_float(bias[elem_begin + e]);
            o = from_float<T>(out);
        }
        if (fuse.residual != nullptr) {
            float rv;
            if (fuse.res_ll.tag_base != nullptr) {
                // residual vector lives in an LL buffer that an earlier kernel of this step already validated
                const int gi = fuse.res_off + elem_begin + e;
                const unsigned long long wv = reinterpret_cast<const unsigned long long*>(fuse.residual)[gi >> 1];
                const uint32_t d = uint32_t(wv & 0xffffffffull);
                rv = __half2float(__ushort_as_half(uint16_t((gi & 1) ? (d >> 16) : (d & 0xffffu))));
            }
            else {
                rv = to_float(fuse.residual[int64_t(m) * fuse.ldr + elem_begin + e]);
            }
            o = from_float<T>(to_float(o) + rv);
        }
        return o;
    };
    if constexpr (M == 1 && XREG) {
        // every operand of the tail is already in registers: partial sums -> scale (+bias) [-> SiLU * up] (+residual) -> store
        auto raw = [&](int r) -> float {
            float a = 0.f;
#pragma unroll
            for (int wi = 0; wi < kWarps; ++wi)
                a += gemv_partial[r * kWarps + wi];
            return a;
        };
        const bool plain_res = fuse.residual != nullptr && fuse.res_ll.tag_base == nullptr;
        if (!ll_out && (fuse.residual == nullptr || plain_res)) {
            if (tid < nelem) {
                T o;
                if (pairs) {
                    const T g16    = from_float<T>(raw(2 * tid) * pre_s0);
                    const T u16    = from_float<T>(raw(2 * tid + 1) * pre_s1);
                    const float gf = to_float(g16);
                    o = from_float<T>(to_float(from_float<T>(gf / (1.f + __expf(-gf)))) * to_float(u16));
                }
                else {
                    float out = raw(tid) * pre_s0;
                    if (bias != nullptr)
                        out += pre_bias;
                    o = from_float<T>(out);
                }
                if (plain_res)
                    o = from_float<T>(to_float(o) + pre_res);
                y[elem_begin + tid] = o;
            }
            trace_ev(TRACE_GEMV, 5);
            return;
        }
    }
    if constexpr (M == 1 && DTypeOf<T>::value == EETQ_B200_F16) {
        if (ll_out) {
            for (int wd = tid; wd < nelem / 2; wd += kThreads) {
                const uint32_t lo = uint32_t(__half_as_ushort(out_value(2 * wd, 0)));
                const uint32_t hi = uint32_t(__half_as_ushort(out_value(2 * wd + 1, 0)));
                ll_push_word(fuse.push, (elem_begin >> 1) + wd, lo | (hi << 16));
            }
            return;
        }
    }
    for (int idx = tid; idx < nelem * M; idx += kThreads) {
        const int e = idx / M;
        const int m = idx - e * M;
        y[int64_t(m) * ldy + elem_begin + e] = out_value(e, m);
    }
}

// launch geometry shared by launch_variant and make_next_hint
int gemv_grid(int sm_count, int ctas_per_sm, int N, int align)
{
    constexpr int kMaxRows = 96;  // rows per CTA bound (sizes the partial-sum buffer)
    int grid = sm_count * ctas_per_sm;
    while ((N + grid - 1) / grid + align > kMaxRows)  // keep rows/CTA <= kMaxRows, and grid a multiple of the SM count
        grid += sm_count;
    if (grid > N / align)
        grid = N / align;
    return grid;
}

template <typename T, int M, int KITERS, int R, bool XREG, int WB>
int launch_variant(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N,
                   int K, const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv: device query failed");
        return EETQ_B200_ECUDA;
    }
    constexpr int kCtasPerSm = min_ctas(M, KITERS, XREG);
    const int align          = (fuse.epi == GEMV_EPI_SILU_PAIRS ? 2 : 1) * (fuse.push.world > 0 ? 2 : 1);
    const int grid           = gemv_grid(di.sm_count, kCtasPerSm, N, align);
    const int max_rows = (N + grid - 1) / grid + align;
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
        cudaLaunchKernelEx(&cfg, w8a16_gemv_kernel<T, M, KITERS, R, XREG, WB>, x, ldx, w, scales, bias, y, ldy, N, K, fuse);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T, int M, int WB>
int dispatch_k(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N, int K,
               const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    const int nchunks = K / 16;
    const int kiters  = (nchunks + kThreads - 1) / kThreads;
    // register-resident activations while the slice stays small (fp16: 8 regs, bf16: 16 regs per 16 values)
    constexpr int kMaxXregIters = (DTypeOf<T>::value == EETQ_B200_F16) ? 8 / M : 4 / M;
    // rows per group are the same for int8 and int4: doubling them for the half-as-long int4 rows (same bytes in flight) measured
    // 2-8 % slower (more zero-padded rows; the int4 kernel is issue-bound, profiles/r02_kbench_int4_simt_rows8.json)
#define EB_GEMV_CASE(KI, RS, RB)                                                                                        \
    if (kiters == KI) {                                                                                                 \
        if constexpr (KI <= kMaxXregIters)                                                                             \
            return launch_variant<T, M, KI, (M <= 2 ? RS : RB), true, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse,  \
                                                                          pdl, stream);                                  \
    }
    EB_GEMV_CASE(1, 8, 4)
    EB_GEMV_CASE(2, 4, 2)
    EB_GEMV_CASE(3, (M == 1 ? 2 : 4), 1)  // M = 2 runs one CTA per SM there: four rows per group keep enough bytes in flight
    EB_GEMV_CASE(4, 2, 1)
#undef EB_GEMV_CASE
    // general path: activations re-read through L1 (no fused prologue / LL input there)
    if (fuse.xmode != GEMV_X_PLAIN || fuse.x_ll.tag_base != nullptr) {
        set_error("gemv: fused RMSNorm / SiLU-mul prologue and LL input need M * ceil(K/4096) <= %d (got M=%d, K=%d)", M * kMaxXregIters, M, K);
        return EETQ_B200_EINVAL;
    }
    constexpr int R = (M <= 2) ? 8 : 4;
    return launch_variant<T, M, 1, R, false, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
}

template <typename T, int WB>
int dispatch_m(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int M, int N,
               int K, const GemvFuse<T>& fuse, bool pdl, cudaStream_t stream)
{
    switch (M) {
        case 1: return dispatch_k<T, 1, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 2: return dispatch_k<T, 2, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 3: return dispatch_k<T, 3, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        case 4: return dispatch_k<T, 4, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
        default: break;
    }
    if constexpr (WB == 8) {
        switch (M) {
            case 5: return dispatch_k<T, 5, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
            case 6: return dispatch_k<T, 6, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
            case 7: return dispatch_k<T, 7, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
            case 8: return dispatch_k<T, 8, WB>(x, ldx, w, scales, bias, y, ldy, N, K, fuse, pdl, stream);
            default: break;
        }
    }
    set_error("gemv: M=%d out of range [1,%d]", M, WB == 8 ? EETQ_B200_GEMV_MAX_M : EETQ_B200_GEMV4_SIMT_MAX_M);
    return EETQ_B200_EINVAL;
}

template <typename T>
GemvFuse<T> make_fuse(const GemvExtras& ex)
{
    GemvFuse<T> f{};
    f.norm_weight = static_cast<const T*>(ex.norm_weight);
    f.residual    = static_cast<const T*>(ex.residual);
    f.ldr         = ex.ldr;
    f.eps         = ex.eps;
    f.xmode       = ex.xmode;
    f.epi         = ex.epi;
    f.x_ll        = ex.x_ll;
    f.res_ll      = ex.res_ll;
    f.res_off     = ex.res_off;
    f.push        = ex.push;
    // measured on one GPU with local LL buffers: skipping the wait is slower (509 vs 537 tok/s, the early pollers compete with the
    // producer's weight stream), so it stays opt-in
    static const bool ll_nowait = [] {
        const char* e = getenv("EETQ_B200_LL_NOWAIT");
        return e != nullptr && e[0] == '1';
    }();
    f.nowait = (ll_nowait && ex.x_ll.tag_base != nullptr && (ex.residual == nullptr || ex.res_ll.tag_base != nullptr)) ? 1 : 0;
    static const bool l2_next = [] {
        const char* e = getenv("EETQ_B200_L2_NEXT");
        return !(e != nullptr && e[0] == '0');
    }();
    if (l2_next && ex.next_w != nullptr)
        f.next = make_next_hint(ex.next_w, ex.next_n, ex.next_k);
    return f;
}

}  // namespace

EB_TRACE_SETTER(trace_set_gemv)

NextHint make_next_hint(const void* w, int64_t N, int64_t K)
{
    NextHint h;
    if (w == nullptr || N <= 0 || K <= 0)
        return h;
    // measured flat between 6 and 24 MB; >= 48 MB evicts the current kernel's own stream from L2
    static const long long budget = [] {
        const char* e = getenv("EETQ_B200_L2_NEXT_MB");
        return static_cast<long long>((e != nullptr && e[0] != '\0') ? atoi(e) : 12) << 20;
    }();
    h.w     = static_cast<const uint8_t*>(w);
    h.bytes = (N * K < budget ? N * K : budget) & ~15ll;
    return h;
}

int launch_gemv(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                int M, int64_t N, int64_t K, int dtype, const GemvExtras& ex, bool pdl, cudaStream_t stream)
{
    const uint8_t* wu = reinterpret_cast<const uint8_t*>(w);
    const bool ll     = ex.x_ll.tag_base != nullptr || ex.res_ll.tag_base != nullptr || ex.push.world > 0;
    if (ll && !(M == 1 && dtype == EETQ_B200_F16)) {
        set_error("gemv: LL exchange buffers need M = 1 and fp16");
        return EETQ_B200_EINVAL;
    }
    if (ex.epi == GEMV_EPI_SILU_PAIRS && (bias != nullptr || (N % 4) != 0)) {
        set_error("gemv: the SiLU*up pair epilogue takes no bias and needs N %% 4 == 0");
        return EETQ_B200_EINVAL;
    }
    if (ex.wbits != 8 && ex.wbits != 4) {
        set_error("gemv: weights must be 8 or 4 bits wide (got %d)", ex.wbits);
        return EETQ_B200_EINVAL;
    }
    if (dtype == EETQ_B200_F16) {
        using T = __half;
        if (ex.wbits == 4)
            return dispatch_m<T, 4>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                    static_cast<T*>(y), ldy, M, int(N), int(K), make_fuse<T>(ex), pdl, stream);
        return dispatch_m<T, 8>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                static_cast<T*>(y), ldy, M, int(N), int(K), make_fuse<T>(ex), pdl, stream);
    }
    if (dtype == EETQ_B200_BF16) {
        using T = __nv_bfloat16;
        if (ex.wbits == 4)
            return dispatch_m<T, 4>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                    static_cast<T*>(y), ldy, M, int(N), int(K), make_fuse<T>(ex), pdl, stream);
        return dispatch_m<T, 8>(static_cast<const T*>(x), ldx, wu, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                static_cast<T*>(y), ldy, M, int(N), int(K), make_fuse<T>(ex), pdl, stream);
    }
    set_error("gemv: unsupported activation dtype %d", dtype);
    return EETQ_B200_EINVAL;
}

}  // namespace eetq_b200
