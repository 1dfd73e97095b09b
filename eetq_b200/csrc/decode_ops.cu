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
                                                       __half* __restrict__ y, int H, float eps, const unsigned* wait_flags,
                                                       int world, const int* epoch)
{
    __shared__ float red[16];
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    p2p_wait_flags(wait_flags, world, epoch);
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

// ---------------------------------------------------------------------------------------------------------------
// Fused RoPE + KV-append + split-KV decode attention + split merge, one query token, ONE launch per layer.
//   grid = (heads, splits), 128 threads, head_dim 128.  CTA (h, s) owns the fixed cache rows [64 s, 64 s + 64): thread (rowlane = t/16,
//   sub = t%16) holds 16-byte slices of 8 K rows and 8 V rows IN REGISTERS -- all 16 loads are issued up front (256 B in
//   flight per thread), so the whole KV read of the layer is in flight at once and the kernel is a pure HBM stream.
//   The CTA that owns the newest position rotates k, appends k/v to the cache and uses them from shared memory.
//   Partials {o[128], m, l} go to scratch; the last CTA of a head (atomic ticket, self-resetting) merges them.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATT_D       = 128;
constexpr int ATT_THREADS = 128;
constexpr int ATT_ROWS    = 64;  // cache positions per CTA (8 per rowlane)

__device__ __forceinline__ float dot8(const uint4& kv, const float (&q)[8])
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&kv.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&kv.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&kv.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&kv.w));
    return q[0] * a.x + q[1] * a.y + q[2] * b.x + q[3] * b.y + q[4] * c.x + q[5] * c.y + q[6] * d.x + q[7] * d.y;
}
__device__ __forceinline__ void axpy8(float p, const uint4& vv, float (&acc)[8])
{
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&vv.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&vv.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&vv.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&vv.w));
    acc[0] = fmaf(p, a.x, acc[0]); acc[1] = fmaf(p, a.y, acc[1]); acc[2] = fmaf(p, b.x, acc[2]); acc[3] = fmaf(p, b.y, acc[3]);
    acc[4] = fmaf(p, c.x, acc[4]); acc[5] = fmaf(p, c.y, acc[5]); acc[6] = fmaf(p, d.x, acc[6]); acc[7] = fmaf(p, d.y, acc[7]);
}

__global__ void __launch_bounds__(ATT_THREADS) attn_fused_kernel(const __half* __restrict__ qkv, const __half* __restrict__ cos_t,
                                                                  const __half* __restrict__ sin_t, const int* __restrict__ pos_p,
                                                                  __half* __restrict__ kcache, __half* __restrict__ vcache,
                                                                  float* __restrict__ partial, int* __restrict__ tickets,
                                                                  __half* __restrict__ out, int H, int max_ctx, float scale,
                                                                  const unsigned* wait_flags, int world, const int* epoch)
{
    __shared__ float q_s[ATT_D];
    __shared__ __align__(16) __half knew[ATT_D];
    __shared__ __align__(16) __half vnew[ATT_D];
    __shared__ float sc[ATT_ROWS];
    __shared__ float red[ATT_THREADS / 32];
    __shared__ float osum[8][ATT_D];
    __shared__ int is_last_s;

    const int t       = threadIdx.x;
    const int lane    = t & 31;
    const int warp    = t >> 5;
    const int sub     = t & 15;   // which 8-dim slice of the head
    const int rowlane = t >> 4;   // 0..7
    const int head    = blockIdx.x;
    const int splits  = gridDim.y;
    const int split   = blockIdx.y;

    pdl_launch_dependents();
    // Split s owns the FIXED cache rows [64 s, 64 s + 64): the loads below need neither `pos` nor anything the preceding
    // kernel produces (rows < pos were written by earlier STEPS; rows >= pos are masked out later), so the whole KV
    // stream of the layer is in flight before the dependency wait and overlaps the tail of the q|k|v GEMV.
    const int p0 = split * ATT_ROWS;
    uint4 kreg[8], vreg[8];
    const __half* kbase = kcache + (int64_t(head) * max_ctx + p0) * ATT_D + sub * 8;
    const __half* vbase = vcache + (int64_t(head) * max_ctx + p0) * ATT_D + sub * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = i * 8 + rowlane;
        kreg[i]     = (p0 + j < max_ctx) ? ldg_stream_128(kbase + int64_t(j) * ATT_D) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = i * 8 + rowlane;
        vreg[i]     = (p0 + j < max_ctx) ? ldg_stream_128(vbase + int64_t(j) * ATT_D) : make_uint4(0u, 0u, 0u, 0u);
    }
    const int pos = *pos_p;                            // written at the end of the previous step
    const int L   = pos + 1;                           // attend to positions [0, pos]
    const int n   = max(0, min(ATT_ROWS, L - p0));     // valid rows of this split
    const bool owns_new = (pos >= p0) && (pos < p0 + ATT_ROWS);
    float rope_c = 0.f, rope_s = 0.f;
    if (t < ATT_D / 2) {
        rope_c = __half2float(cos_t[int64_t(pos) * (ATT_D / 2) + t]);
        rope_s = __half2float(sin_t[int64_t(pos) * (ATT_D / 2) + t]);
    }
    pdl_wait_prior_grids();  // qkv of the current token comes from the preceding GEMV
    p2p_wait_flags(wait_flags, world, epoch);  // column-sharded q|k|v: every rank's slice must have landed

    // RoPE of q (every CTA) and of the new k (owner CTA); HF rotate_half convention evaluated in fp16
    if (t < ATT_D / 2) {
        const float c = rope_c, s = rope_s;
        const int a = head * ATT_D + t, b = a + ATT_D / 2;
        const float q0 = __half2float(qkv[a]), q1 = __half2float(qkv[b]);
        q_s[t]             = __half2float(__hadd(__float2half_rn(q0 * c), __float2half_rn(-q1 * s))) * scale;
        q_s[t + ATT_D / 2] = __half2float(__hadd(__float2half_rn(q1 * c), __float2half_rn(q0 * s))) * scale;
        if (owns_new) {
            const float k0 = __half2float(qkv[H + a]), k1 = __half2float(qkv[H + b]);
            const __half r0 = __hadd(__float2half_rn(k0 * c), __float2half_rn(-k1 * s));
            const __half r1 = __hadd(__float2half_rn(k1 * c), __float2half_rn(k0 * s));
            knew[t] = r0; knew[t + ATT_D / 2] = r1;
            const int64_t row = (int64_t(head) * max_ctx + pos) * ATT_D;
            kcache[row + t] = r0; kcache[row + t + ATT_D / 2] = r1;
            const __half v0 = qkv[2 * H + a], v1 = qkv[2 * H + b];
            vnew[t] = v0; vnew[t + ATT_D / 2] = v1;
            vcache[row + t] = v0; vcache[row + t + ATT_D / 2] = v1;
        }
    }
    __syncthreads();

    float qf[8];
#pragma unroll
    for (int d = 0; d < 8; ++d)
        qf[d] = q_s[sub * 8 + d];

    // scores: 16 lanes share a row
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = i * 8 + rowlane;
        uint4 kv    = kreg[i];
        if (owns_new && j == n - 1)
            kv = *reinterpret_cast<const uint4*>(&knew[sub * 8]);
        float d = dot8(kv, qf);
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1)
            d += __shfl_xor_sync(0xffffffffu, d, o);
        if (sub == 0 && j < n)
            sc[j] = d;
    }
    __syncthreads();

    // softmax statistics of this chunk
    float m = (t < n) ? sc[t] : -INFINITY;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0)
        red[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float e = 0.f;
    if (t < n) {
        e     = __expf(sc[t] - m);
        sc[t] = e;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0)
        red[warp] = e;
    __syncthreads();
    const float l = red[0] + red[1] + red[2] + red[3];

    // o = sum_j p_j V_j : each thread accumulates its 8 dims over its rows, then the 8 rowlanes are summed in smem
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = i * 8 + rowlane;
        if (j < n) {
            uint4 vv = vreg[i];
            if (owns_new && j == n - 1)
                vv = *reinterpret_cast<const uint4*>(&vnew[sub * 8]);
            axpy8(sc[j], vv, acc);
        }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d)
        osum[rowlane][sub * 8 + d] = acc[d];
    __syncthreads();
    float o = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
        o += osum[r][t];

    float* mine = partial + (int64_t(head) * splits + split) * (ATT_D + 2);
    mine[t] = o;
    if (t == 0) {
        mine[ATT_D]     = (n > 0) ? m : -INFINITY;
        mine[ATT_D + 1] = (n > 0) ? l : 0.f;
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        const int old = atomicAdd(tickets + head, 1);
        const int last = (old == splits - 1) ? 1 : 0;
        if (last)
            tickets[head] = 0;  // self-cleaning for the next launch
        is_last_s = last;
    }
    __syncthreads();
    if (is_last_s) {
        __threadfence();
        const float* p = partial + int64_t(head) * splits * (ATT_D + 2);
        // merge: split statistics first (parallel over splits, staged in smem), then each thread merges its own dim with
        // independent loads kept in flight (a serial chain of L2 round trips here used to cost more than the KV stream)
        float mm = -INFINITY, den = 0.f, num = 0.f;
        for (int base = 0; base < splits; base += ATT_ROWS) {
            const int cnt = min(ATT_ROWS, splits - base);
            __syncthreads();
            if (t < cnt) {
                sc[t]         = __ldcg(p + (base + t) * (ATT_D + 2) + ATT_D);       // m_s
                osum[0][t]    = __ldcg(p + (base + t) * (ATT_D + 2) + ATT_D + 1);   // l_s
            }
            __syncthreads();
            float cm = mm;
            for (int s2 = 0; s2 < cnt; ++s2)
                cm = fmaxf(cm, sc[s2]);
            const float rescale = (mm == -INFINITY) ? 0.f : __expf(mm - cm);
            num *= rescale;
            den *= rescale;
            mm = cm;
            int s2 = 0;
            for (; s2 + 4 <= cnt; s2 += 4) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = __ldcg(p + (base + s2 + u) * (ATT_D + 2) + t);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float ms = sc[s2 + u];
                    const float w  = (ms == -INFINITY) ? 0.f : __expf(ms - mm);
                    num = fmaf(w, v[u], num);
                    den = fmaf(w, osum[0][s2 + u], den);
                }
            }
            for (; s2 < cnt; ++s2) {
                const float ms = sc[s2];
                const float w  = (ms == -INFINITY) ? 0.f : __expf(ms - mm);
                num = fmaf(w, __ldcg(p + (base + s2) * (ATT_D + 2) + t), num);
                den = fmaf(w, osum[0][s2], den);
            }
        }
        out[head * ATT_D + t] = __float2half_rn(num / den);
    }
}

}  // namespace
}  // namespace eetq_b200

using namespace eetq_b200;

extern "C" {

int eetq_b200_decode_rmsnorm_p2p(const void* x, const void* w, void* y, int64_t M, int64_t H, float eps, const void* wait_flags,
                                 int world, const void* epoch, int pdl, void* stream);
int eetq_b200_decode_attention_p2p(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                                   void* vcache, void* partial, void* tickets, void* out, int64_t H, int64_t D, int64_t max_ctx,
                                   const void* wait_flags, int world, const void* epoch, int pdl, void* stream);

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
    return eetq_b200_decode_rmsnorm_p2p(x, w, y, M, H, eps, nullptr, 1, nullptr, pdl, stream);
}

int eetq_b200_decode_rmsnorm_p2p(const void* x, const void* w, void* y, int64_t M, int64_t H, float eps, const void* wait_flags,
                                 int world, const void* epoch, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w && y && M > 0 && H > 0, "decode_rmsnorm: bad argument");
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned(M)), dim3(512), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_kernel, static_cast<const __half*>(x), static_cast<const __half*>(w),
                                     static_cast<__half*>(y), int(H), eps, static_cast<const unsigned*>(wait_flags), world,
                                     static_cast<const int*>(epoch)));
    count_launch();
    return EETQ_B200_OK;
}

// Number of KV splits the fused attention kernel needs for a cache of max_ctx positions (<= 64 positions per CTA).
int64_t eetq_b200_decode_attention_splits(int64_t max_ctx) { return (max_ctx + ATT_ROWS - 1) / ATT_ROWS; }

// Fused RoPE + KV append + attention for one token at position *pos (reads [0, pos], writes cache row pos).
//   qkv [3H] raw projections (not modified); kcache/vcache [H/D][max_ctx][D] (head-major); partial: (H/D) * splits * 130 floats scratch; tickets: H/D ints, ZERO on
//   first use (the kernel leaves them zero); out [H].
int eetq_b200_decode_attention(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                               void* vcache, void* partial, void* tickets, void* out, int64_t H, int64_t D, int64_t max_ctx,
                               int pdl, void* stream)
{
    return eetq_b200_decode_attention_p2p(qkv, cos_t, sin_t, pos_i32, kcache, vcache, partial, tickets, out, H, D, max_ctx, nullptr, 1,
                                          nullptr, pdl, stream);
}

int eetq_b200_decode_attention_p2p(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                                   void* vcache, void* partial, void* tickets, void* out, int64_t H, int64_t D, int64_t max_ctx,
                                   const void* wait_flags, int world, const void* epoch, int pdl, void* stream)
{
    EB_CHECK_ARG(qkv && cos_t && sin_t && pos_i32 && kcache && vcache && partial && tickets && out,
                 "decode_attention: null pointer argument");
    EB_CHECK_ARG(D == ATT_D && H % D == 0, "decode_attention: head_dim must be 128");
    EB_CHECK_ARG(max_ctx >= 1 && max_ctx <= (1 << 20), "decode_attention: bad max_ctx");
    const int heads  = int(H / D);
    const int splits = int(eetq_b200_decode_attention_splits(max_ctx));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(unsigned(heads), unsigned(splits)), dim3(ATT_THREADS), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_fused_kernel, static_cast<const __half*>(qkv), static_cast<const __half*>(cos_t),
                                     static_cast<const __half*>(sin_t), static_cast<const int*>(pos_i32), static_cast<__half*>(kcache),
                                     static_cast<__half*>(vcache), static_cast<float*>(partial), static_cast<int*>(tickets),
                                     static_cast<__half*>(out), int(H), int(max_ctx), 1.0f / sqrtf(float(D)),
                                     static_cast<const unsigned*>(wait_flags), world, static_cast<const int*>(epoch)));
    count_launch();
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


// eetq_b200_w8a16_gemv_fused + an L2 prefetch of the KV cache rows [0, *pos) (layout [heads][max_ctx][128] fp16) that the
// attention kernel launched right after it will read: issued by the GEMV CTAs once their weight stream is in flight.
int eetq_b200_w8a16_gemv_fused_kvprefetch(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* norm_weight,
                                          float eps, int xmode, const void* residual, int64_t ldr, void* y, int64_t ldy, int64_t M,
                                          int64_t N, int64_t K, int dtype, const void* kcache, const void* vcache, const void* pos_i32,
                                          int64_t heads, int64_t max_ctx, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w_b200 && scales && y && kcache && vcache && pos_i32, "w8a16_gemv_fused_kvprefetch: null pointer argument");
    EB_CHECK_ARG(M >= 1 && M <= EETQ_B200_GEMV_MAX_M && K > 0 && N > 0 && K % 64 == 0 && N % 64 == 0, "w8a16_gemv_fused_kvprefetch: bad shape");
    GemvExtras ex;
    ex.norm_weight = norm_weight;
    ex.residual    = residual;
    ex.ldr         = ldr;
    ex.eps         = eps;
    ex.xmode       = xmode;
    ex.pf_k        = kcache;
    ex.pf_v        = vcache;
    ex.pf_pos      = static_cast<const int*>(pos_i32);
    ex.pf_heads    = int(heads);
    ex.pf_max_ctx  = int(max_ctx);
    return launch_gemv(x, ldx, w_b200, scales, nullptr, y, ldy, int(M), N, K, dtype, ex, pdl != 0, static_cast<cudaStream_t>(stream));
}

// Up to 4 dependent M=1 fp16 GEMVs in one launch (see w8a16_gemv_chain_kernel).  `phases` is an array of
// eetq_b200_gemv_phase (layout-identical to eetq_b200::GemvChainPhase); counters: >= nphases-1 uint32, zero on first use,
// PRIVATE to this chain position (they only ever increase); epoch: device int32 >= 1 that increases by 1 per launch.
int eetq_b200_w8a16_gemv_chain(const void* phases, int nphases, void* counters, const void* epoch, int pdl, void* stream)
{
    EB_CHECK_ARG(phases && counters && epoch, "w8a16_gemv_chain: null pointer argument");
    return launch_gemv_chain(static_cast<const GemvChainPhase*>(phases), nphases, static_cast<unsigned*>(counters),
                             static_cast<const int*>(epoch), pdl != 0, static_cast<cudaStream_t>(stream));
}

// Column-sharded variant with the all-gather fused into the epilogue over NVLink peer memory (see GemvP2P in common.cuh).
//   peer_y[r]    : rank r's copy of the FULL output vector, offset to this rank's first row (device-mapped peer pointer)
//   peer_flag[r] : address on rank r of flags[slot][this rank];  local_flags: this rank's flags[slot][0..world)
//   ticket       : local uint32, zero between calls;  epoch: device int32 that strictly increases every decode step
int eetq_b200_w8a16_gemv_fused_p2p(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* norm_weight,
                                   float eps, int xmode, const void* residual, int64_t ldr, int64_t M, int64_t N_local, int64_t K,
                                   int dtype, int world, const uint64_t* peer_y, const uint64_t* peer_flag, const void* local_flags,
                                   const void* wait_flags, void* ticket, const void* epoch, int64_t ldy, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w_b200 && scales && peer_y && peer_flag && local_flags && ticket && epoch, "gemv_fused_p2p: null pointer argument");
    EB_CHECK_ARG(world >= 2 && world <= 8, "gemv_fused_p2p: world must be in [2, 8]");
    EB_CHECK_ARG(M >= 1 && M <= EETQ_B200_GEMV_MAX_M, "gemv_fused_p2p: M must be in [1, %d]", EETQ_B200_GEMV_MAX_M);
    EB_CHECK_ARG(K > 0 && N_local > 0 && K % 64 == 0 && N_local % 64 == 0, "gemv_fused_p2p: K and N_local must be positive multiples of 64");
    GemvExtras ex;
    ex.norm_weight = norm_weight;
    ex.residual    = residual;
    ex.ldr         = ldr;
    ex.eps         = eps;
    ex.xmode       = xmode;
    ex.p2p.world   = world;
    for (int r = 0; r < world; ++r) {
        ex.p2p.peer_y[r]    = peer_y[r];
        ex.p2p.peer_flag[r] = peer_flag[r];
    }
    ex.p2p.local_flags = static_cast<const unsigned*>(local_flags);
    ex.p2p.wait_flags  = static_cast<const unsigned*>(wait_flags);
    ex.p2p.ticket      = static_cast<unsigned*>(ticket);
    ex.p2p.epoch       = static_cast<const int*>(epoch);
    void* y_self       = reinterpret_cast<void*>(peer_y[0]);  // unused in p2p mode (stores go through peer_y)
    return launch_gemv(x, ldx, w_b200, scales, nullptr, y_self, ldy, int(M), N_local, K, dtype, ex, pdl != 0,
                       static_cast<cudaStream_t>(stream));
}

}  // extern "C"
