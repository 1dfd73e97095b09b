// quantize.cu -- GPU per-output-channel symmetric INT8 quantiser and weight-layout kernels (sm_100a).
//
// Replaces, bit-exactly, the reference's single-threaded host quantiser
//   ft::symmetric_quantize            /root/reference/csrc/cutlass_kernels/cutlass_preprocessors.cc:581-678
// and its 4-pass host layout transform
//   preprocess_weights_for_mixed_gemm cutlass_preprocessors.cc:497-534
// with HBM-bound kernels.  The target layout is NOT the reference's interleaved one but the
// "b200 layout": biased bytes, output-feature-major  w_b200[n*K + k] = uint8(q[k, n] + 128)  (DESIGN.md section 3),
// which both the streaming GEMV (rows are contiguous K-vectors) and the tcgen05 GEMM (K-major UMMA
// operand via a 2-D TMA box) consume directly, and whose column shards are contiguous byte ranges.
//
// Algorithmic bytes per weight element: abs-max pass reads sizeof(T); quantise pass reads
// sizeof(T) and writes 1 (+1 if the row-major copy is requested).
#include "common.cuh"
#include "int4_layout.cuh"

namespace eetq_b200 {

namespace {

// ---------------------------------------------------------------------------------------------------
// pass 1: amax[n] = max_k |w[k,n]|   (cutlass_preprocessors.cc:623-628)
// Each thread owns VEC adjacent columns (one 16-byte load per row) and a slab of rows; slabs are merged
// with atomicMax on the (non-negative) float bit pattern.  NaN inputs are ignored exactly like
// std::max(acc, NaN) does in the reference.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colabsmax_kernel(const T* __restrict__ w, int64_t K, int64_t N, int rows_per_slab,
                                                        float* __restrict__ amax)
{
    constexpr int VEC = 16 / sizeof(T);
    const int64_t col0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * VEC;
    if (col0 >= N)
        return;
    const int64_t k_begin = int64_t(blockIdx.y) * rows_per_slab;
    const int64_t k_end   = min(K, k_begin + rows_per_slab);

    float m[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        m[j] = 0.f;

    const T* p = w + k_begin * N + col0;
#pragma unroll 4
    for (int64_t k = k_begin; k < k_end; ++k, p += N) {
        const uint4 raw = *reinterpret_cast<const uint4*>(p);
        const T* v      = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            m[j] = fmaxf(m[j], fabsf(to_float(v[j])));  // fmaxf(a, NaN) == a
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        atomicMax(reinterpret_cast<unsigned int*>(amax + col0 + j), __float_as_uint(m[j]));
}

// s32[n] = amax[n] * (1/128) in fp32 (1/8 for int4); stored scale = T(s32[n])   (cutlass_preprocessors.cc:610, :631-635)
template <typename T>
__global__ void finalize_scales_kernel(float* __restrict__ s32, T* __restrict__ scales, int64_t N, float range_scale)
{
    const int64_t n = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (n < N) {
        const float s = s32[n] * range_scale;
        s32[n]        = s;
        scales[n]     = from_float<T>(s);
    }
}

// q = int8(max(-128, min(127, round(w / s))))  with the reference's comparison order (NaN -> 127),
// IEEE division by the FP32 scale and round-half-away-from-zero   (cutlass_preprocessors.cc:644-648)
__device__ __forceinline__ int quant_one(float w, float s)
{
    const float r  = roundf(__fdiv_rn(w, s));
    const float lo = (r < 127.f) ? r : 127.f;
    const float hi = (-128.f < lo) ? lo : -128.f;
    return static_cast<int>(hi);
}

// ---------------------------------------------------------------------------------------------------
// pass 2: quantise a 64(k) x 64(n) tile and emit it transposed (b200 layout), optionally also row-major.
// ---------------------------------------------------------------------------------------------------
constexpr int QT       = 64;
constexpr int QT_PITCH = QT + 16;  // keeps 16-byte alignment of each smem row

template <typename T>
__global__ void __launch_bounds__(256) quantize_tile_kernel(const T* __restrict__ w, const float* __restrict__ s32,
                                                            int64_t K, int64_t N, int8_t* __restrict__ q_b200,
                                                            int8_t* __restrict__ q_kn)
{
    __shared__ __align__(16) int8_t tile[QT][QT_PITCH];  // tile[n][k]

    const int64_t n0 = int64_t(blockIdx.x) * QT;
    const int64_t k0 = int64_t(blockIdx.y) * QT;
    const int t      = threadIdx.x;
    const int nl     = (t & 15) * 4;  // 4 adjacent columns
    const int kl     = t >> 4;        // 0..15

    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        s[j] = s32[n0 + nl + j];

#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int k  = kl + 16 * p;
        const T* src = w + (k0 + k) * N + n0 + nl;
        T v[4];
        if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(src);
        }
        else {
            *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(src);
        }
        int q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[j]            = quant_one(to_float(v[j]), s[j]);
            tile[nl + j][k] = static_cast<int8_t>(q[j] ^ 0x80);  // stored biased: u = q + 128
        }
        if (q_kn != nullptr) {
            const uint32_t packed = (uint32_t(q[0]) & 0xffu) | ((uint32_t(q[1]) & 0xffu) << 8)
                                    | ((uint32_t(q[2]) & 0xffu) << 16) | ((uint32_t(q[3]) & 0xffu) << 24);
            *reinterpret_cast<uint32_t*>(q_kn + (k0 + k) * N + n0 + nl) = packed;
        }
    }
    __syncthreads();

    const int n  = t >> 2;
    const int kc = (t & 3) * 16;
    const uint4 out                                                 = *reinterpret_cast<const uint4*>(&tile[n][kc]);
    *reinterpret_cast<uint4*>(q_b200 + (n0 + n) * K + k0 + kc) = out;
}

// ---------------------------------------------------------------------------------------------------
// byte-matrix transpose src[rows][cols] -> dst[cols][rows]   (pack: rows=K, cols=N; unpack: rows=N, cols=K)
// ---------------------------------------------------------------------------------------------------
// Every byte is XOR-ed with 0x80 on the way (signed int8 <-> biased uint8).
__global__ void __launch_bounds__(256) transpose_bytes_kernel(const int8_t* __restrict__ src, int64_t rows, int64_t cols,
                                                              int8_t* __restrict__ dst)
{
    __shared__ __align__(16) int8_t tile[QT][QT_PITCH];  // tile[r][c]
    const int64_t c0 = int64_t(blockIdx.x) * QT;
    const int64_t r0 = int64_t(blockIdx.y) * QT;
    const int t      = threadIdx.x;
    {
        const int r  = t >> 2;
        const int cc = (t & 3) * 16;
        *reinterpret_cast<uint4*>(&tile[r][cc]) = *reinterpret_cast<const uint4*>(src + (r0 + r) * cols + c0 + cc);
    }
    __syncthreads();
    {
        const int c  = t >> 2;
        const int rc = (t & 3) * 16;
        __align__(16) int8_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            v[j] = int8_t(tile[rc + j][c] ^ 0x80);
        *reinterpret_cast<uint4*>(dst + (c0 + c) * rows + r0 + rc) = *reinterpret_cast<const uint4*>(v);
    }
}

// ---------------------------------------------------------------------------------------------------
// reference sm75..sm89 interleaved layout <-> b200 layout.
// Reference bytes viewed as [N/2][K/64][2][4][16] = (pair, ktile, c, g, p) hold uint8(q[k,n]+128) with
// n = 2*pair + c and k = 64*ktile + 16*g + (p>>1) + 8*(p&1)   (closed form of
// cutlass_preprocessors.cc:497-534).  One thread moves one 16-byte group.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) from_ref_layout_kernel(const uint4* __restrict__ w_ref, int64_t K, int64_t N,
                                                              int8_t* __restrict__ q_b200)
{
    const int64_t gi = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // 16-byte group index in reference order
    if (gi >= K * N / 16)
        return;
    const int64_t ktiles = K / 64;
    const int g          = int(gi & 3);
    const int c          = int((gi >> 2) & 1);
    const int64_t kt     = (gi >> 3) % ktiles;
    const int64_t pair   = (gi >> 3) / ktiles;
    const uint4 in       = w_ref[gi];
    uint4 out;
    // out byte j = in byte p(j),  p = 2*(j&7) + (j>>3);  both layouts store the biased byte u = q + 128
    out.x = __byte_perm(in.x, in.y, 0x6420);
    out.y = __byte_perm(in.z, in.w, 0x6420);
    out.z = __byte_perm(in.x, in.y, 0x7531);
    out.w = __byte_perm(in.z, in.w, 0x7531);
    const int64_t n = 2 * pair + c;
    *reinterpret_cast<uint4*>(q_b200 + n * K + 64 * kt + 16 * g) = out;
}

__global__ void __launch_bounds__(256) to_ref_layout_kernel(const int8_t* __restrict__ q_b200, int64_t K, int64_t N,
                                                            uint4* __restrict__ w_ref)
{
    const int64_t gi = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gi >= K * N / 16)
        return;
    const int64_t ktiles = K / 64;
    const int g          = int(gi & 3);
    const int c          = int((gi >> 2) & 1);
    const int64_t kt     = (gi >> 3) % ktiles;
    const int64_t pair   = (gi >> 3) / ktiles;
    const int64_t n      = 2 * pair + c;
    const uint4 in       = *reinterpret_cast<const uint4*>(q_b200 + n * K + 64 * kt + 16 * g);
    uint4 out;
    // out byte p = in byte j(p),  j = (p>>1) + 8*(p&1)
    out.x = __byte_perm(in.x, in.z, 0x5140);
    out.y = __byte_perm(in.x, in.z, 0x7362);
    out.z = __byte_perm(in.y, in.w, 0x5140);
    out.w = __byte_perm(in.y, in.w, 0x7362);
    w_ref[gi] = out;
}

// ---------------------------------------------------------------------------------------------------
// packed int4 (QuantType::PACKED_INT4_WEIGHT_ONLY)
// ---------------------------------------------------------------------------------------------------
// q = max(-8, min(7, int(round(w / s))))   (cutlass_preprocessors.cc:655-660).  The reference converts to int BEFORE clamping;
// a NaN (0/0 of an all-zero column, NaN weight) becomes INT_MIN there (x86-64 cvttss2si) and clamps to -8.
__device__ __forceinline__ int quant_one4(float w, float s)
{
    const float r = roundf(__fdiv_rn(w, s));
    if (r != r)
        return -8;
    const int v = __float2int_rz(fminf(fmaxf(r, -100.f), 100.f));
    return max(-8, min(7, v));
}

// quantise a 64(k) x 64(n) tile; emit it in the b200 int4 layout, optionally also packed row-major [K][N/2]
template <typename T>
__global__ void __launch_bounds__(256) quantize4_tile_kernel(const T* __restrict__ w, const float* __restrict__ s32,
                                                             int64_t K, int64_t N, uint8_t* __restrict__ q4_b200,
                                                             uint8_t* __restrict__ q4_kn)
{
    __shared__ __align__(16) uint8_t tile[QT][QT_PITCH];  // tile[n][k] = biased nibble u = q + 8

    const int64_t n0 = int64_t(blockIdx.x) * QT;
    const int64_t k0 = int64_t(blockIdx.y) * QT;
    const int t      = threadIdx.x;
    const int nl     = (t & 15) * 4;  // 4 adjacent columns
    const int kl     = t >> 4;        // 0..15

    float s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        s[j] = s32[n0 + nl + j];

#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int k  = kl + 16 * p;
        const T* src = w + (k0 + k) * N + n0 + nl;
        T v[4];
        if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(src);
        }
        else {
            *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(src);
        }
        int q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            q[j]            = quant_one4(to_float(v[j]), s[j]);
            tile[nl + j][k] = static_cast<uint8_t>(q[j] + 8);
        }
        if (q4_kn != nullptr) {
            // low nibble = even column, two's-complement nibbles (cutlass_preprocessors.cc:664-666)
            const uint16_t packed = uint16_t((q[0] & 15) | ((q[1] & 15) << 4) | ((q[2] & 15) << 8) | ((q[3] & 15) << 12));
            *reinterpret_cast<uint16_t*>(q4_kn + (((k0 + k) * N + n0 + nl) >> 1)) = packed;
        }
    }
    __syncthreads();

    // thread -> (row n, 16 consecutive k) = two 32-bit words of the b200 int4 layout
    const int n  = t >> 2;
    const int kc = (t & 3) * 16;
    uint32_t wd[2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
        wd[h] = b200_pack_word(&tile[n][kc + 8 * h]);
    *reinterpret_cast<uint2*>(q4_b200 + (((n0 + n) * K + k0 + kc) >> 1)) = make_uint2(wd[0], wd[1]);
}

// ---------------------------------------------------------------------------------------------------
// int4 layout conversions (run once per weight at load time) and the int4 -> int8 widening pass; the index arithmetic lives in
// int4_layout.cuh (host + device, so tests/test_int4_host.py can run the very same functions on the CPU against the oracle).
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) nibble_layout_kernel(const uint8_t* __restrict__ src, int64_t K, int64_t N,
                                                            uint32_t* __restrict__ dst)
{
    const int64_t wi = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // output word
    if (wi < K * N / 8)
        dst[wi] = nibble_layout_word<MODE>(src, K, N, wi);
}

// b200 int4 -> b200 int8 (u8 = q + 128 = u4 + 120): feeds the tcgen05 GEMM for M > 4.  One thread widens 16 k (8 -> 16 bytes).
__global__ void __launch_bounds__(256) widen4to8_kernel(const uint2* __restrict__ src, int64_t chunks, uint4* __restrict__ dst)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= chunks)
        return;
    const uint2 in = src[i];
    uint4 o;
    widen4to8_word(in.x, o.x, o.y);
    widen4to8_word(in.y, o.z, o.w);
    dst[i] = o;
}

template <typename T>
int quantize4_impl(const T* w, int64_t K, int64_t N, uint8_t* q4_b200, T* scales, float* s32, uint8_t* q4_kn, cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    EB_CHECK_CUDA(cudaMemsetAsync(s32, 0, sizeof(float) * N, stream));
    {
        const int rows_per_slab = 64;
        dim3 grid(unsigned((N / VEC + 255) / 256), unsigned((K + rows_per_slab - 1) / rows_per_slab));
        colabsmax_kernel<T><<<grid, 256, 0, stream>>>(w, K, N, rows_per_slab, s32);
    }
    finalize_scales_kernel<T><<<unsigned((N + 255) / 256), 256, 0, stream>>>(s32, scales, N, 1.f / 8.f);
    {
        dim3 grid(unsigned(N / QT), unsigned(K / QT));
        quantize4_tile_kernel<T><<<grid, 256, 0, stream>>>(w, s32, K, N, q4_b200, q4_kn);
    }
    count_launch(3);
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

template <typename T>
int quantize_impl(const T* w, int64_t K, int64_t N, int8_t* q_b200, T* scales, float* s32, int8_t* q_kn,
                  cudaStream_t stream)
{
    constexpr int VEC = 16 / sizeof(T);
    EB_CHECK_CUDA(cudaMemsetAsync(s32, 0, sizeof(float) * N, stream));
    {
        const int rows_per_slab = 64;
        dim3 grid(unsigned((N / VEC + 255) / 256), unsigned((K + rows_per_slab - 1) / rows_per_slab));
        colabsmax_kernel<T><<<grid, 256, 0, stream>>>(w, K, N, rows_per_slab, s32);
    }
    finalize_scales_kernel<T><<<unsigned((N + 255) / 256), 256, 0, stream>>>(s32, scales, N, 1.f / 128.f);
    {
        dim3 grid(unsigned(N / QT), unsigned(K / QT));
        quantize_tile_kernel<T><<<grid, 256, 0, stream>>>(w, s32, K, N, q_b200, q_kn);
    }
    count_launch(3);
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

}  // namespace

int launch_quantize(const void* w_kn, int w_dtype, int64_t K, int64_t N, int8_t* q_b200, void* scales, float* s32,
                    int8_t* q_kn, cudaStream_t stream)
{
    switch (w_dtype) {
        case EETQ_B200_F16:
            return quantize_impl<__half>(static_cast<const __half*>(w_kn), K, N, q_b200, static_cast<__half*>(scales),
                                         s32, q_kn, stream);
        case EETQ_B200_BF16:
            return quantize_impl<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(w_kn), K, N, q_b200,
                                                static_cast<__nv_bfloat16*>(scales), s32, q_kn, stream);
        case EETQ_B200_F32:
            return quantize_impl<float>(static_cast<const float*>(w_kn), K, N, q_b200, static_cast<float*>(scales), s32,
                                        q_kn, stream);
        default:
            set_error("quantize: unsupported weight dtype %d", w_dtype);
            return EETQ_B200_EINVAL;
    }
}

int launch_transpose_bytes(const int8_t* src, int64_t rows, int64_t cols, int8_t* dst, cudaStream_t stream)
{
    dim3 grid(unsigned(cols / QT), unsigned(rows / QT));
    transpose_bytes_kernel<<<grid, 256, 0, stream>>>(src, rows, cols, dst);
    count_launch();
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

int launch_from_ref_layout(const uint8_t* w_ref, int64_t K, int64_t N, int8_t* q_b200, cudaStream_t stream)
{
    const int64_t groups = K * N / 16;
    from_ref_layout_kernel<<<unsigned((groups + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(w_ref), K,
                                                                               N, q_b200);
    count_launch();
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

int launch_to_ref_layout(const int8_t* q_b200, int64_t K, int64_t N, uint8_t* w_ref, cudaStream_t stream)
{
    const int64_t groups = K * N / 16;
    to_ref_layout_kernel<<<unsigned((groups + 255) / 256), 256, 0, stream>>>(q_b200, K, N,
                                                                             reinterpret_cast<uint4*>(w_ref));
    count_launch();
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

int launch_quantize4(const void* w_kn, int w_dtype, int64_t K, int64_t N, uint8_t* q4_b200, void* scales, float* s32,
                     uint8_t* q4_kn, cudaStream_t stream)
{
    switch (w_dtype) {
        case EETQ_B200_F16:
            return quantize4_impl<__half>(static_cast<const __half*>(w_kn), K, N, q4_b200, static_cast<__half*>(scales), s32,
                                          q4_kn, stream);
        case EETQ_B200_BF16:
            return quantize4_impl<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(w_kn), K, N, q4_b200,
                                                 static_cast<__nv_bfloat16*>(scales), s32, q4_kn, stream);
        case EETQ_B200_F32:
            return quantize4_impl<float>(static_cast<const float*>(w_kn), K, N, q4_b200, static_cast<float*>(scales), s32, q4_kn,
                                         stream);
        default:
            set_error("quantize4: unsupported weight dtype %d", w_dtype);
            return EETQ_B200_EINVAL;
    }
}

int launch_nibble_layout(int mode, const uint8_t* src, int64_t K, int64_t N, uint8_t* dst, cudaStream_t stream)
{
    const int64_t words = K * N / 8;
    const unsigned grid = unsigned((words + 255) / 256);
    uint32_t* d         = reinterpret_cast<uint32_t*>(dst);
    switch (mode) {
        case NIB_PACK4: nibble_layout_kernel<NIB_PACK4><<<grid, 256, 0, stream>>>(src, K, N, d); break;
        case NIB_UNPACK4: nibble_layout_kernel<NIB_UNPACK4><<<grid, 256, 0, stream>>>(src, K, N, d); break;
        case NIB_FROM_REF4: nibble_layout_kernel<NIB_FROM_REF4><<<grid, 256, 0, stream>>>(src, K, N, d); break;
        case NIB_TO_REF4: nibble_layout_kernel<NIB_TO_REF4><<<grid, 256, 0, stream>>>(src, K, N, d); break;
        default: set_error("nibble layout: bad mode %d", mode); return EETQ_B200_EINVAL;
    }
    count_launch();
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

int launch_widen4to8(const uint8_t* q4_b200, int64_t K, int64_t N, int8_t* q_b200, cudaStream_t stream)
{
    const int64_t chunks = K * N / 16;
    widen4to8_kernel<<<unsigned((chunks + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint2*>(q4_b200), chunks,
                                                                       reinterpret_cast<uint4*>(q_b200));
    count_launch();
    EB_CHECK_CUDA(cudaGetLastError());
    return EETQ_B200_OK;
}

}  // namespace eetq_b200
