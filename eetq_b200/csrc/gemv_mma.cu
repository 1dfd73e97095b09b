// gemv_mma.cu -- decode-time streaming kernel for 2..8 token rows, int8 or int4 weights (sm_100a).
//
// Replaces, for the batched-decode rows, the reference's weight_only_batched_gemv<..., BatchSize 2..4>
//   /root/reference/csrc/weightOnlyBatchedGemv/kernel.h:294-468 (dispatch kernelLauncher.cu:165-199; Int4b details :68-116)
// whose cost grows with the batch (one HFMA2 chain per row per weight pair, fp16 accumulation).  Here the weight stream is
// the same register-fed LDG.128 stream as the M = 1 kernel (gemv.cu), but the arithmetic goes through warp-level
// mma.sync.m16n8k16 with the WEIGHTS in the A role -- D[feature (16), token (8)] += W[feature, k] * X[k, token] -- so the
// instruction count per weight byte is independent of M and accumulation is fp32 (design notes at the kernel).
// A first version with the tokens in the 16-row role (4 MMAs per 16 weights per lane, 8-row tiles) was measured slower on every
// cell it served (profiles/r02_kbench_mma2.json) and removed.
//
// Algorithmic bytes per call: K*N*bits/8 + 2*N + 2*M*K + 2*M*N; each weight byte is read exactly once.
#include "common.cuh"

namespace eetq_b200 {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps   = 8;
constexpr int kBuf     = 8;  // 16-byte weight loads per register buffer

template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    if constexpr (DTypeOf<T>::value == EETQ_B200_F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// one 32-bit word of 4 biased bytes -> two operand registers (k pairs {0,1} and {2,3})
template <typename T>
__device__ __forceinline__ void cvt_word(uint32_t w, uint32_t& b0, uint32_t& b1)
{
    if constexpr (DTypeOf<T>::value == EETQ_B200_F16) {
        b0 = __byte_perm(w, 0x64646464u, 0x5140);  // fp16(1024 + u0), fp16(1024 + u1)
        b1 = __byte_perm(w, 0x64646464u, 0x5342);
    }
    else {
        const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)) - 8388736.f;
        const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)) - 8388736.f;
        const float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652)) - 8388736.f;
        const float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653)) - 8388736.f;
        b0 = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);  // exact: |q| <= 128 fits bf16
        b1 = __byte_perm(__float_as_uint(f2), __float_as_uint(f3), 0x7632);
    }
}

// =====================================================================================================================
// D[feature (16), token (8)] += W[feature, k] * X[k, token]: one m16n8k16 consumes 16 features x 16 k = 8 weights per lane; int8
// bytes or int4 nibbles (WB):
//   * a work item is (16-row tile, k-block); lane (g = lane/4, t = lane%4) loads 16 bytes of row g and 16 bytes of row g + 8 at
//     byte offset 16 t of the block's 64 bytes (int8: 64 k per block, 16 k per lane; int4: 128 k per block, 32 k per lane);
//   * every group of 4 consecutive k of a lane (one int8 word, half an int4 word) is one MMA: A = {row g pair 0, row g+8 pair 0,
//     row g pair 1, row g+8 pair 1}, B = the same two k pairs of token g from the staged activations (one LDS.64) -- the MMA's
//     k slots are a fixed permutation of the real k, identical on both operands;
//   * int8: PRMT under the exponent byte -> fp16(1024 + u), the constant (1024 + 128) * sum_k x is removed in the epilogue; int4
//     (b200 int4 layout): (word >> 4p) & 0x000f000f | 0x64006400 -> fp16x2(1024 + u) of the adjacent-k pair p, minus 1032 -> exact q;
//   * instructions per 16 weights per lane: int8 8 PRMT + 2 LDS.64 + 2 MMA (the SIMT kernel: 24 per token row),
//     int4 14 shift/LOP3 + 8 HADD2 + 2 LDS.64 + 2 MMA (SIMT: 30 per token row).
// Accumulation is fp32 inside the tensor core.  Rows >= M of the token tile are zero (lanes g >= MP feed zero B fragments).
// =====================================================================================================================
template <typename T, int WB>
struct MmaOffset {
    // int8 fp16: the MMA consumes fp16(1024 + u) as is and (1024 + 128) * sum_k x is removed in the epilogue.  int4 nibbles are
    // converted to exact q first: their signal is 16x smaller against the same constant, and the tensor core's truncating fp32
    // accumulation of the constant's products showed up at K = 32768 (1.2e-3 norm-relative, over the 1e-3 bar)
    static constexpr float value = (DTypeOf<T>::value == EETQ_B200_F16 && WB == 8) ? 1152.f : 0.f;
};

// adjacent-k pair p (0..3) of one b200 int4 word -> an fp16x2 / bf16x2 operand register
template <typename T>
__device__ __forceinline__ uint32_t cvt_nib_pair(uint32_t w, int p)
{
    if constexpr (DTypeOf<T>::value == EETQ_B200_F16) {
        const uint32_t pr = ((w >> (4 * p)) & 0x000f000fu) | 0x64006400u;  // fp16(1024 + u_2p), fp16(1024 + u_2p+1)
        const uint32_t c  = 0x64086408u;                                     // fp16x2(1032): 1024 + the storage bias 8
        const __half2 q   = __hsub2(*reinterpret_cast<const __half2*>(&pr), *reinterpret_cast<const __half2*>(&c));  // exact
        return *reinterpret_cast<const uint32_t*>(&q);
    }
    else {
        const float f0 = __uint_as_float(((w >> (4 * p)) & 0xfu) | 0x4B000000u) - 8388616.f;
        const float f1 = __uint_as_float(((w >> (4 * p + 16)) & 0xfu) | 0x4B000000u) - 8388616.f;
        return __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);  // exact: |q| <= 8 fits bf16
    }
}

// dynamic smem: [MP rows of (K + 4) T] | [kWarps][tiles][16][MP] fp32 partials | [MP] fp32 row sums
template <typename T, int MP, int WB>
__global__ void __launch_bounds__(kThreads, 2)
    w8a16_gemv_mma_kernel(const T* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ w, const T* __restrict__ scales,
                           const T* __restrict__ bias, const T* __restrict__ residual, int64_t ldr, T* __restrict__ y, int64_t ldy,
                           int M, int N, int K, int max_tiles)
{
    constexpr int KB  = 512 / WB;  // k per block: 64 bytes of one row
    constexpr int KL  = KB / 4;    // k per lane: 16 bytes
    constexpr int IPB = kBuf / 2;  // work items per register buffer (two 16-byte loads each)
    extern __shared__ __align__(16) uint8_t mma_smem[];
    const int xstride = 2 * K + 8;  // bytes per staged activation row (8-byte pad spreads the rows over the banks)
    uint8_t* xs       = mma_smem;
    float* partial    = reinterpret_cast<float*>(mma_smem + ((MP * xstride + 15) & ~15));
    float* xsum       = partial + kWarps * max_tiles * 16 * MP;

    const int tid  = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int g    = lane >> 2;  // A fragment: feature rows g and g + 8 of the tile; B fragment: token
    const int t    = lane & 3;

    const int row_begin   = int((int64_t(blockIdx.x) * N) / gridDim.x);
    const int row_end     = int((int64_t(blockIdx.x + 1) * N) / gridDim.x);
    const int nrows       = row_end - row_begin;
    const int ntiles      = (nrows + 15) >> 4;
    const int64_t rowbytes = (int64_t(K) * WB) >> 3;

    // this warp's contiguous range of k-blocks
    const int nkb_total = K / KB;
    const int kb0       = (warp * nkb_total) / kWarps;
    const int nkb       = ((warp + 1) * nkb_total) / kWarps - kb0;
    const int total     = ntiles * nkb;  // work items (tile, k-block) of this warp, tile-major

    pdl_launch_dependents();

    uint4 wb[2][kBuf];
    int l_tile = 0, l_j = 0;  // loader state: next item to fetch
    auto load_buf = [&](uint4 (&buf)[kBuf]) {
#pragma unroll
        for (int i = 0; i < IPB; ++i) {
            const int row = row_begin + l_tile * 16 + g;
            const uint8_t* p = w + int64_t(row) * rowbytes + int64_t(kb0 + l_j) * 64 + t * 16;
            buf[2 * i]     = (l_tile < ntiles && row < row_end) ? ldg_stream_128(p) : make_uint4(0u, 0u, 0u, 0u);
            buf[2 * i + 1] = (l_tile < ntiles && row + 8 < row_end) ? ldg_stream_128(p + 8 * rowbytes) : make_uint4(0u, 0u, 0u, 0u);
            if (++l_j == nkb) {
                l_j = 0;
                ++l_tile;
            }
        }
    };
    // weights do not depend on the previous kernel: start streaming before the dependency wait
    if (nkb > 0) {
        load_buf(wb[0]);
        load_buf(wb[1]);
    }
    pdl_wait_prior_grids();

    // stage the activations: every 8-byte piece of the M real rows travels by cp.async (all pieces in flight at once, no
    // registers, ONE global-memory latency -- a load/store loop here cost 4 x M dependent latencies, ~2 us); rows >= M are zero
    {
        const int pieces = K >> 2;
        for (int m = 0; m < MP; ++m) {
            uint8_t* dst = xs + m * xstride;
            if (m < M) {
                const T* src = x + int64_t(m) * ldx;
                for (int c = tid; c < pieces; c += kThreads)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(dst + c * 8))),
                                 "l"(src + int64_t(c) * 4)
                                 : "memory");
            }
            else {
                for (int c = tid; c < pieces; c += kThreads)
                    *reinterpret_cast<uint2*>(dst + c * 8) = make_uint2(0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if constexpr (MmaOffset<T, WB>::value != 0.f) {
            // per-row sums of the staged activations (for the constant removed in the epilogue)
            float* red = partial;  // reused before any partial is written
#pragma unroll
            for (int m = 0; m < MP; ++m) {
                float v = 0.f;
                for (int c = tid; c < pieces; c += kThreads) {
                    const uint2 p  = *reinterpret_cast<const uint2*>(xs + m * xstride + c * 8);
                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&p.x));
                    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&p.y));
                    v += (a.x + a.y) + (b.x + b.y);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1)
                    v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0)
                    red[m * kWarps + warp] = v;
            }
            __syncthreads();
            if (tid < MP) {
                float v = 0.f;
#pragma unroll
                for (int wi = 0; wi < kWarps; ++wi)
                    v += red[tid * kWarps + wi];
                xsum[tid] = v;
            }
            __syncthreads();
        }
    }

    int c_tile = 0, c_j = 0;  // consumer state
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // D fragment: acc[0..1] = (feature g, tokens 2t, 2t+1), acc[2..3] = (feature g + 8, tokens 2t, 2t+1)
    auto flush = [&](int tile) {
        if (2 * t < MP) {
            float* base = partial + (warp * max_tiles + tile) * 16 * MP;
            *reinterpret_cast<float2*>(base + g * MP + 2 * t)       = make_float2(acc[0], acc[1]);
            *reinterpret_cast<float2*>(base + (g + 8) * MP + 2 * t) = make_float2(acc[2], acc[3]);
        }
        acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    };
    auto compute_buf = [&](const uint4 (&buf)[kBuf], int count) {
#pragma unroll
        for (int i = 0; i < IPB; ++i) {
            if (i < count) {
                const uint32_t lo[4] = {buf[2 * i].x, buf[2 * i].y, buf[2 * i].z, buf[2 * i].w};                  // row g
                const uint32_t hi[4] = {buf[2 * i + 1].x, buf[2 * i + 1].y, buf[2 * i + 1].z, buf[2 * i + 1].w};  // row g + 8
                const uint8_t* xrow  = xs + g * xstride + ((kb0 + c_j) * KB + t * KL) * 2;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    if constexpr (WB == 8) {
                        uint32_t a[4];
                        cvt_word<T>(lo[jj], a[0], a[2]);
                        cvt_word<T>(hi[jj], a[1], a[3]);
                        uint2 xv = make_uint2(0u, 0u);
                        if (g < MP)
                            xv = *reinterpret_cast<const uint2*>(xrow + jj * 8);
                        mma_16816<T>(acc, a, xv.x, xv.y);
                    }
                    else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t a[4];
                            a[0] = cvt_nib_pair<T>(lo[jj], 2 * h);
                            a[1] = cvt_nib_pair<T>(hi[jj], 2 * h);
                            a[2] = cvt_nib_pair<T>(lo[jj], 2 * h + 1);
                            a[3] = cvt_nib_pair<T>(hi[jj], 2 * h + 1);
                            uint2 xv = make_uint2(0u, 0u);
                            if (g < MP)
                                xv = *reinterpret_cast<const uint2*>(xrow + jj * 16 + h * 8);
                            mma_16816<T>(acc, a, xv.x, xv.y);
                        }
                    }
                }
                if (++c_j == nkb) {
                    flush(c_tile);
                    c_j = 0;
                    ++c_tile;
                }
            }
        }
    };

    for (int done = 0; done < total; done += 2 * IPB) {
        const int n0 = min(IPB, total - done);
        compute_buf(wb[0], n0);
        if (done + 2 * IPB < total)
            load_buf(wb[0]);
        const int n1 = min(IPB, total - done - IPB);
        if (n1 > 0) {
            compute_buf(wb[1], n1);
            if (done + 3 * IPB < total)
                load_buf(wb[1]);
        }
    }
    __syncthreads();

    // epilogue: cross-warp sum, remove the constant * sum(x), per-channel scale (+bias, +residual), store
    for (int idx = tid; idx < nrows * M; idx += kThreads) {
        const int r    = idx / M;
        const int m    = idx - r * M;
        const int tile = r >> 4, f = r & 15;
        float s        = 0.f;
        for (int wi = 0; wi < kWarps; ++wi)  // warps without any k-block never wrote their slot
            if (((wi + 1) * nkb_total) / kWarps > (wi * nkb_total) / kWarps)
                s += partial[((wi * max_tiles + tile) * 16 + f) * MP + m];
        if constexpr (MmaOffset<T, WB>::value != 0.f)
            s -= MmaOffset<T, WB>::value * xsum[m];
        const int n = row_begin + r;
        float out   = s * to_float(scales[n]);
        if (bias != nullptr)
            out += to_float(bias[n]);
        T o = from_float<T>(out);
        if (residual != nullptr)
            o = from_float<T>(to_float(o) + to_float(residual[int64_t(m) * ldr + n]));
        y[int64_t(m) * ldy + n] = o;
    }
}

template <typename T, int MP, int WB>
int launch_mma(const T* x, int64_t ldx, const uint8_t* w, const T* scales, const T* bias, const T* residual, int64_t ldr, T* y,
                int64_t ldy, int M, int N, int K, bool pdl, cudaStream_t stream)
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("gemv_mma: device query failed");
        return EETQ_B200_ECUDA;
    }
    constexpr int kMaxRows = 96;
    const size_t x_bytes   = (size_t(MP) * (2 * size_t(K) + 8) + 15) & ~size_t(15);
    // geometry for a given number of CTAs per SM; two per SM whenever two footprints (+1 KB reserved each) fit the SM's shared memory
    int grid = 0, max_tiles = 0;
    size_t smem = 0;
    auto plan = [&](int ctas_per_sm) {
        grid = di.sm_count * ctas_per_sm;
        while ((N + grid - 1) / grid > kMaxRows)
            grid += di.sm_count;
        if (grid > N / 16)
            grid = N / 16 > 0 ? N / 16 : 1;
        const int max_rows = (N + grid - 1) / grid;
        max_tiles          = (max_rows + 15) / 16 + 1;
        smem               = x_bytes + size_t(kWarps) * max_tiles * 16 * MP * sizeof(float) + MP * sizeof(float) + 64;
    };
    plan(2);
    if (2 * (smem + 1024) > size_t(di.max_smem_optin) + 1024)
        plan(1);
    if (smem > size_t(di.max_smem_optin)) {
        set_error("gemv_mma: %zu bytes of shared memory needed (M=%d, K=%d)", smem, M, K);
        return EETQ_B200_EINVAL;
    }
    auto kernel = w8a16_gemv_mma_kernel<T, MP, WB>;
    static size_t attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && attr_set[dev] < smem) {
        EB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(di.max_smem_optin)));
        attr_set[dev] = size_t(di.max_smem_optin);
    }
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
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, x, ldx, w, scales, bias, residual, ldr, y, ldy, M, N, K, max_tiles);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemv_mma launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T, int WB>
int launch_mma_mp(const void* x, int64_t ldx, const uint8_t* w, const void* scales, const void* bias, const void* residual, int64_t ldr,
                   void* y, int64_t ldy, int M, int N, int K, bool pdl, cudaStream_t stream)
{
    if (M <= 2)  // two staged rows: the activations of K = 11008 still leave room for two CTAs per SM
        return launch_mma<T, 2, WB>(static_cast<const T*>(x), ldx, w, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                     static_cast<const T*>(residual), ldr, static_cast<T*>(y), ldy, M, N, K, pdl, stream);
    if (M <= 4)
        return launch_mma<T, 4, WB>(static_cast<const T*>(x), ldx, w, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                     static_cast<const T*>(residual), ldr, static_cast<T*>(y), ldy, M, N, K, pdl, stream);
    return launch_mma<T, 8, WB>(static_cast<const T*>(x), ldx, w, static_cast<const T*>(scales), static_cast<const T*>(bias),
                                 static_cast<const T*>(residual), ldr, static_cast<T*>(y), ldy, M, N, K, pdl, stream);
}

}  // namespace

// 1 <= M <= 8, int8 or int4 weights (int4 needs K % 128 == 0); the staged activations must fit in shared memory
bool gemv_mma_supported(int M, int64_t K, int wbits)
{
    const DeviceInfo& di = device_info();
    if (!di.ok || M < 1 || M > 8 || (wbits != 8 && wbits != 4) || (wbits == 4 && (K % 128) != 0))
        return false;
    const size_t mp = M <= 2 ? 2 : M <= 4 ? 4 : 8;
    return mp * (2 * size_t(K) + 8) + 49152 <= size_t(di.max_smem_optin);
}

int launch_gemv_mma(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                     int64_t ldr, void* y, int64_t ldy, int M, int64_t N, int64_t K, int dtype, int wbits, bool pdl,
                     cudaStream_t stream)
{
    const uint8_t* wu = reinterpret_cast<const uint8_t*>(w);
    if (!gemv_mma_supported(M, K, wbits)) {
        set_error("gemv_mma: unsupported call (M=%d, K=%lld, %d-bit weights)", M, (long long)K, wbits);
        return EETQ_B200_EINVAL;
    }
    if (dtype == EETQ_B200_F16)
        return wbits == 8 ? launch_mma_mp<__half, 8>(x, ldx, wu, scales, bias, residual, ldr, y, ldy, M, int(N), int(K), pdl, stream)
                          : launch_mma_mp<__half, 4>(x, ldx, wu, scales, bias, residual, ldr, y, ldy, M, int(N), int(K), pdl, stream);
    if (dtype == EETQ_B200_BF16)
        return wbits == 8
                   ? launch_mma_mp<__nv_bfloat16, 8>(x, ldx, wu, scales, bias, residual, ldr, y, ldy, M, int(N), int(K), pdl, stream)
                   : launch_mma_mp<__nv_bfloat16, 4>(x, ldx, wu, scales, bias, residual, ldr, y, ldy, M, int(N), int(K), pdl, stream);
    set_error("gemv_mma: unsupported activation dtype %d", dtype);
    return EETQ_B200_EINVAL;
}

}  // namespace eetq_b200
