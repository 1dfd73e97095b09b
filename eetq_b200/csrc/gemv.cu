// gemv.cu -- decode-time w8a16 streaming kernel for M <= 8 (sm_100a, SIMT, no tensor cores).
//
// Replaces the reference decode path
//   weight_only_batched_gemv<half, Int8b, PerChannel, ...>   /root/reference/csrc/weightOnlyBatchedGemv/kernel.h:294-468
//   (dispatch: weightOnlyBatchedGemv/kernelLauncher.cu:165-199, called from fpA_intB_gemm_wrapper.cu:149-160)
// with a design sized for B200's HBM3e rather than a translation of it:
//
//   * weights are in the b200 layout (row n = the K int8 of output feature n, contiguous), so a CTA streams
//     whole rows with perfectly coalesced 128-bit loads; thread t of the CTA owns the 16-byte K-chunks
//     {t, t+THREADS, ...} of EVERY row and keeps the matching activation slice in fp32 REGISTERS (XREG) --
//     the weights flow HBM -> registers -> FMA with no shared-memory hop and zero activation re-reads;
//   * rows are processed R at a time, register double-buffered: 2 x R x KITERS 16-byte loads in flight per
//     thread (256 B/thread, ~128 KB/SM at 2 CTAs/SM) which is what Little's law needs to cover ~6.5 TB/s;
//   * int8 -> fp32 conversion is exact (PRMT into the mantissa of 2^23, one FADD), products/accumulation are
//     fp32 (packed FFMA2), the per-channel scale is applied ONCE per output in the epilogue together with the
//     optional bias; the reference multiplies every weight by the scale in fp16 and accumulates in fp16
//     (kernel.h:355-377, :425-435), so this kernel is strictly more accurate -- parity is checked against the
//     fp32-accumulation oracle at 1e-3 norm-relative (tests/test_gemm_gpu.py);
//   * per-row partial sums are reduced with a transposing warp butterfly (9 shuffles per 8 rows instead of
//     40) and one shared-memory pass at the end of the CTA;
//   * the grid is a multiple of the SM count and rows are split evenly (+-1) over CTAs;
//   * weight loads for the first row group are issued BEFORE griddepcontrol.wait, so under programmatic
//     dependent launch the HBM stream of layer i+1 starts while layer i drains.
//
// Algorithmic bytes per call: K*N + 2*N + 2*M*K + 2*M*N; each weight byte is read exactly once.
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

// 4 signed int8 packed in a word -> 4 exact fp32 values.
// CVT == 0: byte-select I2F.   CVT == 1: bias to unsigned, PRMT the byte into the mantissa of 2^23, subtract.
template <int CVT>
__device__ __forceinline__ void cvt4(uint32_t w, float2& lo, float2& hi)
{
    if constexpr (CVT == 0) {
        lo.x = float(int(int8_t(w & 0xffu)));
        lo.y = float(int(int8_t((w >> 8) & 0xffu)));
        hi.x = float(int(int8_t((w >> 16) & 0xffu)));
        hi.y = float(int(int8_t(w >> 24)));
    }
    else {
        const uint32_t u = w ^ 0x80808080u;
        lo.x             = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7650)) - 8388736.f;
        lo.y             = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7651)) - 8388736.f;
        hi.x             = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7652)) - 8388736.f;
        hi.y             = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7653)) - 8388736.f;
    }
}

// 16 activations of type T (32 bytes) -> 8 float2
template <typename T>
__device__ __forceinline__ void load_x16(const T* p, float2 (&out)[8])
{
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
    const uint32_t raw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if constexpr (sizeof(T) == 2 && DTypeOf<T>::value == EETQ_B200_F16) {
            out[j] = __half22float2(*reinterpret_cast<const __half2*>(&raw[j]));
        }
        else {
            out[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[j]));
        }
    }
}

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

template <typename T, int M, int KITERS, int R, bool XREG, int CVT>
__global__ void __launch_bounds__(kThreads, (XREG && M * KITERS <= 2) ? 2 : 1)
    w8a16_gemv_kernel(const T* __restrict__ x, int64_t ldx, const int8_t* __restrict__ w, const T* __restrict__ scales,
                      const T* __restrict__ bias, T* __restrict__ y, int64_t ldy, int N, int K, int max_rows)
{
    const int tid     = threadIdx.x;
    const int lane    = tid & 31;
    const int warp    = tid >> 5;
    const int nchunks = K >> 4;

    // even split of the N rows over the grid
    const int row_begin = int((int64_t(blockIdx.x) * N) / gridDim.x);
    const int row_end   = int((int64_t(blockIdx.x + 1) * N) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + R - 1) / R;

    // number of K-chunk iterations: compile-time for XREG, run-time otherwise
    const int kiters = XREG ? KITERS : (nchunks + kThreads - 1) / kThreads;

    // let the next kernel in the stream start its own prologue (no-op without PDL)
    pdl_launch_dependents();

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
        pdl_wait_prior_grids();

        float2 xs[M][KITERS][8];
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * kThreads;
                if (c < nchunks) {
                    load_x16<T>(x + int64_t(m) * ldx + int64_t(c) * 16, xs[m][i]);
                }
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        xs[m][i][j] = make_float2(0.f, 0.f);
                }
            }

        auto compute_group = [&](uint4 (&buf)[R][KITERS], int g) {
            float acc[M][R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 a2[M];
#pragma unroll
                for (int m = 0; m < M; ++m)
                    a2[m] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < KITERS; ++i) {
                    const uint32_t words[4] = {buf[r][i].x, buf[r][i].y, buf[r][i].z, buf[r][i].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float2 lo, hi;
                        cvt4<CVT>(words[q], lo, hi);
#pragma unroll
                        for (int m = 0; m < M; ++m) {
                            a2[m] = ffma2(lo, xs[m][i][2 * q], a2[m]);
                            a2[m] = ffma2(hi, xs[m][i][2 * q + 1], a2[m]);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < M; ++m)
                    acc[m][r] = a2[m].x + a2[m].y;
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float tot = warp_reduce_rows<R>(acc[m], lane);
                const int rid   = lane >> (5 - Log2<R>::v);
                if ((lane & ((32 >> Log2<R>::v) - 1)) == 0)
                    gemv_partial[((g * R + rid) * M + m) * kWarps + warp] = tot;
            }
        };

        for (int g = 0; g < ngroups; g += 2) {
            if (g + 1 < ngroups)
                load_group(wb[1], g + 1);
            compute_group(wb[0], g);
            if (g + 1 < ngroups) {
                if (g + 2 < ngroups)
                    load_group(wb[0], g + 2);
                compute_group(wb[1], g + 1);
            }
        }
    }
    else {
        // ------------------------------------------------------------------ activations re-read through L1
        pdl_wait_prior_grids();
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
                float2 a2[M][R];
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        a2[m][r] = make_float2(0.f, 0.f);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float2 xv[8];
                    load_x16<T>(x + int64_t(m) * ldx + int64_t(c) * 16, xv);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const uint32_t words[4] = {buf[r].x, buf[r].y, buf[r].z, buf[r].w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float2 lo, hi;
                            cvt4<CVT>(words[q], lo, hi);
                            a2[m][r] = ffma2(lo, xv[2 * q], a2[m][r]);
                            a2[m][r] = ffma2(hi, xv[2 * q + 1], a2[m][r]);
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        acc[m][r] += a2[m][r].x + a2[m][r].y;
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float tot = warp_reduce_rows<R>(acc[m], lane);
                const int rid   = lane >> (5 - Log2<R>::v);
                if ((lane & ((32 >> Log2<R>::v) - 1)) == 0)
                    gemv_partial[((g * R + rid) * M + m) * kWarps + warp] = tot;
            }
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
        y[int64_t(m) * ldy + n] = from_float<T>(out);
    }
    (void)max_rows;
}

int g_gemv_cvt_mode = 1;  // 0 = I2F byte-select, 1 = PRMT magic (default); switchable for experiments

template <typename T, int M, int KITERS, int R, bool XREG>
int launch_variant(const T* x, int64_t ldx, const int8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N,
                   int K, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv: device query failed");
        return EETQ_B200_ECUDA;
    }
    constexpr int kCtasPerSm = (XREG && M * KITERS <= 2) ? 2 : 1;
    constexpr int kMaxRows   = 96;  // rows per CTA bound (sizes the partial-sum buffer)
    int grid                 = di.sm_count * kCtasPerSm;
    // keep rows/CTA <= kMaxRows, and grid a multiple of the SM count
    while ((N + grid - 1) / grid > kMaxRows)
        grid += di.sm_count;
    if (grid > N)
        grid = N;
    const int max_rows = (N + grid - 1) / grid;
    // partial buffer is indexed by padded group rows: round up to a multiple of R
    const int padded  = ((max_rows + R - 1) / R) * R;
    const size_t smem = size_t(padded) * M * kWarps * sizeof(float);

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

    cudaError_t e;
    if (g_gemv_cvt_mode == 0)
        e = cudaLaunchKernelEx(&cfg, w8a16_gemv_kernel<T, M, KITERS, R, XREG, 0>, x, ldx, w, scales, bias, y, ldy, N, K,
                               max_rows);
    else
        e = cudaLaunchKernelEx(&cfg, w8a16_gemv_kernel<T, M, KITERS, R, XREG, 1>, x, ldx, w, scales, bias, y, ldy, N, K,
                               max_rows);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T, int M>
int dispatch_k(const T* x, int64_t ldx, const int8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int N,
               int K, bool pdl, cudaStream_t stream)
{
    const int nchunks = K / 16;
    const int kiters  = (nchunks + kThreads - 1) / kThreads;
    // register-resident activations when M * kiters * 16 fp32 fit comfortably (<= 64 registers)
    if constexpr (M == 1) {
        if (kiters == 1)
            return launch_variant<T, 1, 1, 8, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        if (kiters == 2)
            return launch_variant<T, 1, 2, 4, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        if (kiters == 3)
            return launch_variant<T, 1, 3, 2, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        if (kiters == 4)
            return launch_variant<T, 1, 4, 2, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
    }
    if constexpr (M == 2) {
        if (kiters == 1)
            return launch_variant<T, 2, 1, 8, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        if (kiters == 2)
            return launch_variant<T, 2, 2, 4, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
    }
    if constexpr (M == 3 || M == 4) {
        if (kiters == 1)
            return launch_variant<T, M, 1, 4, true>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
    }
    // general path: activations re-read through L1
    constexpr int R = (M <= 2) ? 8 : 4;
    return launch_variant<T, M, 1, R, false>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
}

template <typename T>
int dispatch_m(const T* x, int64_t ldx, const int8_t* w, const T* scales, const T* bias, T* y, int64_t ldy, int M, int N,
               int K, bool pdl, cudaStream_t stream)
{
    switch (M) {
        case 1: return dispatch_k<T, 1>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 2: return dispatch_k<T, 2>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 3: return dispatch_k<T, 3>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 4: return dispatch_k<T, 4>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 5: return dispatch_k<T, 5>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 6: return dispatch_k<T, 6>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 7: return dispatch_k<T, 7>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        case 8: return dispatch_k<T, 8>(x, ldx, w, scales, bias, y, ldy, N, K, pdl, stream);
        default: set_error("gemv: M=%d out of range [1,%d]", M, EETQ_B200_GEMV_MAX_M); return EETQ_B200_EINVAL;
    }
}

}  // namespace

extern "C" void eetq_b200_debug_set_gemv_cvt_mode(int mode) { g_gemv_cvt_mode = mode ? 1 : 0; }

int launch_gemv(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                int M, int64_t N, int64_t K, int dtype, bool pdl, cudaStream_t stream)
{
    if (dtype == EETQ_B200_F16)
        return dispatch_m<__half>(static_cast<const __half*>(x), ldx, w, static_cast<const __half*>(scales),
                                  static_cast<const __half*>(bias), static_cast<__half*>(y), ldy, M, int(N), int(K),
                                  pdl, stream);
    if (dtype == EETQ_B200_BF16)
        return dispatch_m<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), ldx, w,
                                         static_cast<const __nv_bfloat16*>(scales),
                                         static_cast<const __nv_bfloat16*>(bias), static_cast<__nv_bfloat16*>(y), ldy, M,
                                         int(N), int(K), pdl, stream);
    set_error("gemv: unsupported activation dtype %d", dtype);
    return EETQ_B200_EINVAL;
}

}  // namespace eetq_b200
