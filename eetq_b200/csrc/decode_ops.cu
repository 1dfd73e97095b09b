// decode_ops.cu -- decode-side glue kernels around the w8a16 GEMV: embedding gather,
// RMSNorm, RoPE + KV-cache append, split-KV decode attention.  They exist so that the headline metric (Llama-2-7B
// decode tokens/s) is not dominated by framework-op launch overhead; they are NOT part of the reference's hot-path
// boundary.  Reference counterparts, for behaviour only:
//   rotary_embedding_neox_kernel  /root/reference/csrc/embedding_kernels/pos_encoding_kernels.cu:12-53
//   generalT5LayerNorm (RMSNorm)  /root/reference/csrc/layernorm_kernels/layernorm.cu:25-51
//   EETLlamaAttention (SDPA path) /root/reference/python/eetq/modules/llama_modules.py:68-149
// All kernels are single-token (M = 1) decode kernels, fp16, launched with optional programmatic dependent launch.
#include "common.cuh"

namespace eetq_b200 {

namespace {

cudaError_t launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, dim3 grid, dim3 block, size_t smem, bool pdl,
                       cudaStream_t stream)
{
    cfg                  = cudaLaunchConfig_t{};
    cfg.gridDim          = grid;
    cfg.blockDim         = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = stream;
    attr[0].id           = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs            = attr;
    cfg.numAttrs         = pdl ? 1 : 0;
    return cudaSuccess;
}

// x[h] = table[token][h]
__global__ void __launch_bounds__(256) embed_kernel(const __half* __restrict__ table, const int64_t* __restrict__ token,
                                                     __half* __restrict__ x, int H)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int64_t t = *token;
    const uint4* src = reinterpret_cast<const uint4*>(table + t * H);
    uint4* dst       = reinterpret_cast<uint4*>(x);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H / 8; i += gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// y = fp16( fp16(x_f32 * rsqrt(mean(x^2) + eps)) * w )   (HF LlamaRMSNorm arithmetic), one CTA per row
__global__ void __launch_bounds__(512) rmsnorm_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                       __half* __restrict__ y, int H, float eps)
{
    __shared__ float red[16];
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const __half* xr = x + int64_t(blockIdx.x) * H;
    __half* yr       = y + int64_t(blockIdx.x) * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        const float v = __half2float(xr[i]);
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0)
        red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i)
        tot += red[i];
    const float r = rsqrtf(tot / float(H) + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x)
        yr[i] = __hmul(__float2half_rn(__half2float(xr[i]) * r), w[i]);
}

// RoPE (HF "rotate_half" convention: pairs (i, i + D/2)) on q and k of one token, then append k, v to the cache.
//   qkv   [3*H] = q | k | v   (q is rotated in place)
//   cos/sin tables [max_pos][D/2] fp16 (HF computes them in fp32 and casts to the model dtype)
//   kcache/vcache [max_ctx][H]
__global__ void __launch_bounds__(256) rope_append_kernel(__half* __restrict__ qkv, const __half* __restrict__ cos_t,
                                                           const __half* __restrict__ sin_t, const int* __restrict__ pos_p,
                                                           __half* __restrict__ kcache, __half* __restrict__ vcache, int H, int D)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int pos  = *pos_p;
    const int half = D / 2;
    const int idx  = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (head, i < D/2)
    if (idx >= H / 2)
        return;
    const int head = idx / half;
    const int i    = idx - head * half;
    const float c  = __half2float(cos_t[int64_t(pos) * half + i]);
    const float s  = __half2float(sin_t[int64_t(pos) * half + i]);
    const int a    = head * D + i;
    const int b    = a + half;
    {
        const float q0 = __half2float(qkv[a]), q1 = __half2float(qkv[b]);
        // HF: q*cos + rotate_half(q)*sin evaluated in fp16: each product and the sum are rounded
        qkv[a] = __hadd(__float2half_rn(q0 * c), __float2half_rn(-q1 * s));
        qkv[b] = __hadd(__float2half_rn(q1 * c), __float2half_rn(q0 * s));
    }
    {
        const float k0 = __half2float(qkv[H + a]), k1 = __half2float(qkv[H + b]);
        kcache[int64_t(pos) * H + a] = __hadd(__float2half_rn(k0 * c), __float2half_rn(-k1 * s));
        kcache[int64_t(pos) * H + b] = __hadd(__float2half_rn(k1 * c), __float2half_rn(k0 * s));
    }
    vcache[int64_t(pos) * H + a] = qkv[2 * H + a];
    vcache[int64_t(pos) * H + b] = qkv[2 * H + b];
}

// Split-KV decode attention, one query token.  grid = (heads, splits), 128 threads, head_dim 128.
//   partial[(head*splits + split)] = { o[128] (unnormalised fp32), m, l }
constexpr int ATT_D       = 128;
constexpr int ATT_THREADS = 128;
constexpr int ATT_MAXCHUNK = 512;  // positions per split held in smem

__global__ void __launch_bounds__(ATT_THREADS) attn_split_kernel(const __half* __restrict__ q, const __half* __restrict__ kcache,
                                                                  const __half* __restrict__ vcache, const int* __restrict__ pos_p,
                                                                  float* __restrict__ partial, int H, float scale)
{
    __shared__ float sc[ATT_MAXCHUNK];
    __shared__ float red[ATT_THREADS / 32];
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int L      = *pos_p + 1;  // attend to positions [0, pos]
    const int head   = blockIdx.x;
    const int splits = gridDim.y;
    const int split  = blockIdx.y;
    const int p0     = int((int64_t(split) * L) / splits);
    const int p1     = int((int64_t(split + 1) * L) / splits);
    const int n      = p1 - p0;
    const int lane   = threadIdx.x & 31;
    const int warp   = threadIdx.x >> 5;
    float* out       = partial + (int64_t(head) * splits + split) * (ATT_D + 2);

    // q slice of this lane: 4 consecutive dims
    float qf[4];
    {
        const uint2 raw = *reinterpret_cast<const uint2*>(q + head * ATT_D + lane * 4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        qf[0] = a.x * scale; qf[1] = a.y * scale; qf[2] = b.x * scale; qf[3] = b.y * scale;
    }
    // scores: each warp takes positions warp, warp+4, ... (4 in flight)
    for (int j0 = warp * 4; j0 < n; j0 += 16) {
        float d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            d[u] = 0.f;
            if (j < n) {
                const uint2 raw = *reinterpret_cast<const uint2*>(kcache + int64_t(p0 + j) * H + head * ATT_D + lane * 4);
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                d[u] = qf[0] * a.x + qf[1] * a.y + qf[2] * b.x + qf[3] * b.y;
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
            for (int u = 0; u < 4; ++u)
                d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
        const float mine = (lane == 0) ? d[0] : (lane == 1) ? d[1] : (lane == 2) ? d[2] : d[3];
        if (lane < 4 && j0 + lane < n)
            sc[j0 + lane] = mine;
    }
    __syncthreads();
    // chunk max
    float m = -INFINITY;
    for (int j = threadIdx.x; j < n; j += ATT_THREADS)
        m = fmaxf(m, sc[j]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0)
        red[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float l = 0.f;
    for (int j = threadIdx.x; j < n; j += ATT_THREADS) {
        const float e = __expf(sc[j] - m);
        sc[j] = e;
        l += e;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0)
        red[warp] = l;
    __syncthreads();
    l = red[0] + red[1] + red[2] + red[3];
    // o[d] = sum_j p_j * V[j][d]; thread = d
    const int dcol = threadIdx.x;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const __half* vp = vcache + int64_t(p0) * H + head * ATT_D + dcol;
    int j = 0;
    for (; j + 4 <= n; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            acc[u] = fmaf(sc[j + u], __half2float(vp[int64_t(j + u) * H]), acc[u]);
    }
    for (; j < n; ++j)
        acc[0] = fmaf(sc[j], __half2float(vp[int64_t(j) * H]), acc[0]);
    out[dcol] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    if (threadIdx.x == 0) {
        out[ATT_D]     = (n > 0) ? m : -INFINITY;
        out[ATT_D + 1] = (n > 0) ? l : 0.f;
    }
}

// merge the split partials: out[head*128 + d] = sum_s w_s o_s[d] / sum_s w_s l_s,  w_s = exp(m_s - m)
__global__ void __launch_bounds__(ATT_THREADS) attn_combine_kernel(const float* __restrict__ partial, __half* __restrict__ out,
                                                                    int splits)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int head = blockIdx.x;
    const float* p = partial + int64_t(head) * splits * (ATT_D + 2);
    float m = -INFINITY;
    for (int s = 0; s < splits; ++s)
        m = fmaxf(m, p[s * (ATT_D + 2) + ATT_D]);
    float num = 0.f, den = 0.f;
    for (int s = 0; s < splits; ++s) {
        const float ms = p[s * (ATT_D + 2) + ATT_D];
        const float w  = (ms == -INFINITY) ? 0.f : __expf(ms - m);
        num = fmaf(w, p[s * (ATT_D + 2) + threadIdx.x], num);
        den = fmaf(w, p[s * (ATT_D + 2) + ATT_D + 1], den);
    }
    out[head * ATT_D + threadIdx.x] = __float2half_rn(num / den);
}

}  // namespace
}  // namespace eetq_b200

using namespace eetq_b200;

extern "C" {

int eetq_b200_decode_embed(const void* table, const void* token_i64, void* x, int64_t H, int pdl, void* stream)
{
    EB_CHECK_ARG(table && token_i64 && x && H % 8 == 0, "decode_embed: bad argument");
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned((H / 8 + 255) / 256)), dim3(256), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, embed_kernel, static_cast<const __half*>(table), static_cast<const int64_t*>(token_i64),
                                     static_cast<__half*>(x), int(H)));
    count_launch();
    return EETQ_B200_OK;
}

int eetq_b200_decode_rmsnorm(const void* x, const void* w, void* y, int64_t M, int64_t H, float eps, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w && y && M > 0 && H > 0, "decode_rmsnorm: bad argument");
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned(M)), dim3(512), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_kernel, static_cast<const __half*>(x), static_cast<const __half*>(w),
                                     static_cast<__half*>(y), int(H), eps));
    count_launch();
    return EETQ_B200_OK;
}

int eetq_b200_decode_rope_append(void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache, void* vcache,
                                 int64_t H, int64_t D, int pdl, void* stream)
{
    EB_CHECK_ARG(qkv && cos_t && sin_t && pos_i32 && kcache && vcache && H % D == 0 && D % 2 == 0, "decode_rope_append: bad argument");
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned((H / 2 + 255) / 256)), dim3(256), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rope_append_kernel, static_cast<__half*>(qkv), static_cast<const __half*>(cos_t),
                                     static_cast<const __half*>(sin_t), static_cast<const int*>(pos_i32),
                                     static_cast<__half*>(kcache), static_cast<__half*>(vcache), int(H), int(D)));
    count_launch();
    return EETQ_B200_OK;
}

// attention for one token over cache positions [0, *pos]; partial: heads*splits*(128+2) floats of scratch.
int eetq_b200_decode_attention(const void* q, const void* kcache, const void* vcache, const void* pos_i32, void* partial,
                               void* out, int64_t H, int64_t D, int64_t splits, int64_t max_ctx, int pdl, void* stream)
{
    EB_CHECK_ARG(q && kcache && vcache && pos_i32 && partial && out, "decode_attention: null pointer argument");
    EB_CHECK_ARG(D == ATT_D && H % D == 0, "decode_attention: head_dim must be 128");
    EB_CHECK_ARG(splits >= 1 && (max_ctx + splits - 1) / splits + 1 <= ATT_MAXCHUNK,
                 "decode_attention: max_ctx/splits must be < %d positions", ATT_MAXCHUNK);
    const int heads = int(H / D);
    cudaStream_t s  = static_cast<cudaStream_t>(stream);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned(heads), unsigned(splits)), dim3(ATT_THREADS), 0, pdl != 0, s);
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_split_kernel, static_cast<const __half*>(q), static_cast<const __half*>(kcache),
                                     static_cast<const __half*>(vcache), static_cast<const int*>(pos_i32),
                                     static_cast<float*>(partial), int(H), 1.0f / sqrtf(float(D))));
    launch_cfg(cfg, attr, dim3(unsigned(heads)), dim3(ATT_THREADS), 0, pdl != 0, s);
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_combine_kernel, static_cast<const float*>(partial), static_cast<__half*>(out),
                                     int(splits)));
    count_launch(2);
    return EETQ_B200_OK;
}

// The decode GEMV with its fusions exposed: optional RMSNorm / SiLU*up on the activation load, optional residual add.
//   xmode: 0 plain, 1 RMSNorm(x; norm_weight, eps), 2 silu(x[:, :K]) * x[:, K:2K]
int eetq_b200_w8a16_gemv_fused(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                               const void* norm_weight, float eps, int xmode, const void* residual, int64_t ldr, void* y,
                               int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w_b200 && scales && y, "w8a16_gemv_fused: null pointer argument");
    EB_CHECK_ARG(M >= 1 && M <= EETQ_B200_GEMV_MAX_M, "w8a16_gemv_fused: M must be in [1, %d]", EETQ_B200_GEMV_MAX_M);
    EB_CHECK_ARG(K > 0 && N > 0 && K % 64 == 0 && N % 64 == 0, "w8a16_gemv_fused: K and N must be positive multiples of 64");
    EB_CHECK_ARG(xmode >= 0 && xmode <= 2 && (xmode != GEMV_X_RMSNORM || norm_weight != nullptr), "w8a16_gemv_fused: bad xmode");
    EB_CHECK_ARG(ldx >= (xmode == GEMV_X_SILU_MUL ? 2 * K : K) && ldy >= N, "w8a16_gemv_fused: bad leading dimension");
    GemvExtras ex;
    ex.norm_weight = norm_weight;
    ex.residual    = residual;
    ex.ldr         = ldr;
    ex.eps         = eps;
    ex.xmode       = xmode;
    return launch_gemv(x, ldx, w_b200, scales, bias, y, ldy, int(M), N, K, dtype, ex, pdl != 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
