// decode_ops.cu -- glue kernels either side of the w8a16 linears: embedding gather, RMSNorm,
// RoPE + KV-cache append, decode attention, lm_head + greedy argmax, and the prefill-side element-wise pieces.  They exist so
// that the headline metric (Llama-2-7B decode tokens/s) is bounded by the weight stream rather than by framework-op launches;
// they are NOT part of the reference's w8a16 boundary.  Reference counterparts, for behaviour:
//   rotary_embedding_neox_kernel  /root/reference/csrc/embedding_kernels/pos_encoding_kernels.cu:12-53   (also exported by name)
//   generalT5LayerNorm            /root/reference/csrc/layernorm_kernels/layernorm.cu:25-51             (also exported by name)
//   EETLlamaAttention (SDPA path) /root/reference/python/eetq/modules/llama_modules.py:68-149
// fp16.  Decode kernels are single-token; every launch can carry the programmatic-dependent-launch attribute.
//
// Multi-GPU (column-sharded linears): vectors that every rank needs (attention output, residual stream,
// MLP activation, arg-max candidates) travel as "LL" buffers -- each 8-byte word = {two fp16 values, 32-bit tag} stored with
// one 8-byte store straight into every peer's copy over NVLink (symmetric memory).  A word is valid when its tag equals the
// tag of the exchange (step counter x exchanges per step + exchange index), so data and flag arrive together: no fence, no
// separate flag, no collective call, and the consumer's prologue simply polls the words it is about to use.
#include <cstdlib>

#include "common.cuh"

namespace eetq_b200 {

namespace {

cudaError_t launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, dim3 grid, dim3 block, size_t smem, bool pdl,
                       cudaStream_t stream, int cluster_y = 0)
{
    cfg                  = cudaLaunchConfig_t{};
    cfg.gridDim          = grid;
    cfg.blockDim         = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream           = stream;
    int n                = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_y > 1) {
        attr[n].id               = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = 1;
        attr[n].val.clusterDim.y = unsigned(cluster_y);
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs    = attr;
    cfg.numAttrs = unsigned(n);
    return cudaSuccess;
}

// x[h] = table[token][h]; plain or LL output (every rank gathers the same row into its own LL buffer)
__global__ void __launch_bounds__(256) embed_kernel(const __half* __restrict__ table, const int64_t* __restrict__ token,
                                                     __half* __restrict__ x, int H, LLTag ll)
{
    // The embedding is the FIRST kernel of a decode step.  It releases its dependents only after the previous step's last
    // kernel (lm_head + arg-max, which advances *pos and the step counter) has completed, so every later kernel of this step
    // may read the position before its own dependency wait.
    trace_ev(TRACE_EMBED, 0);
    pdl_wait_prior_grids();
    trace_ev(TRACE_EMBED, 2);
    pdl_launch_dependents();
    const int64_t t = *token;
    if (ll.tag_base == nullptr) {
        const uint4* src = reinterpret_cast<const uint4*>(table + t * H);
        uint4* dst       = reinterpret_cast<uint4*>(x);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H / 8; i += gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
    else {
        const uint32_t tag     = ll_tag(ll);
        const uint32_t* src    = reinterpret_cast<const uint32_t*>(table + t * H);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(x);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H / 2; i += gridDim.x * blockDim.x)
            dst[i] = ll_pack(src[i], tag);
    }
}

// y = fp16( fp16(x_f32 * rsqrt(mean(x^2) + eps)) * w )   (HF LlamaRMSNorm arithmetic), one CTA per row.
// t5 != 0: the reference's generalT5LayerNorm arithmetic instead -- fp16( (x_f32 * rsqrt(..)) * w_f32 ), clamped to the fp16
// range (layernorm.cu:25-51, clamp_inf_for_half).
__global__ void __launch_bounds__(512) rmsnorm_kernel(const __half* __restrict__ x, int64_t ldx, const __half* __restrict__ w,
                                                       __half* __restrict__ y, int64_t ldy, int H, float eps, int t5)
{
    __shared__ float red[16];
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const __half* xr = x + int64_t(blockIdx.x) * ldx;
    __half* yr       = y + int64_t(blockIdx.x) * ldy;
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
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        if (t5) {
            float v = (__half2float(xr[i]) * r) * __half2float(w[i]);
            v       = fminf(fmaxf(v, -65504.f), 65504.f);
            yr[i]   = __float2half_rn(v);
        }
        else {
            yr[i] = __hmul(__float2half_rn(__half2float(xr[i]) * r), w[i]);
        }
    }
}

// Individually rounded fp16 operations.  The reference's scalar_t is c10::Half, whose operators go through fp32 and round
// after EVERY operation (its SASS, rebuilt here for sm_100a: FMUL, F2FP.F16, FADD, F2FP.F16 -- profiles/r02_ref_rotary_sass.txt);
// __hmul / __hsub on __half would be contracted into one HFMA by ptxas and differ in the last bit.
__device__ __forceinline__ __half rn_mul(__half a, __half b) { return __float2half_rn(__half2float(a) * __half2float(b)); }
__device__ __forceinline__ __half rn_add(__half a, __half b) { return __float2half_rn(__half2float(a) + __half2float(b)); }
__device__ __forceinline__ __half rn_sub(__half a, __half b) { return __float2half_rn(__half2float(a) - __half2float(b)); }

// In-place GPT-NeoX rotary embedding of query and key, the reference's op of the same name
// (pos_encoding_kernels.cu:12-53): one CTA per token, cos_sin_cache [max_position][rot_dim] = cos(rot/2) | sin(rot/2),
// fp16 arithmetic exactly as written there (two rounded products, one rounded difference / sum).
__global__ void rope_neox_kernel(const int64_t* __restrict__ positions, __half* __restrict__ query, __half* __restrict__ key,
                                 const __half* __restrict__ cos_sin_cache, int rot_dim, int stride, int num_heads, int head_size)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int token       = blockIdx.x;
    const int64_t pos     = positions[token];
    const __half* cache   = cos_sin_cache + pos * rot_dim;
    const int embed_dim   = rot_dim / 2;
    const int n           = num_heads * embed_dim;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int head = i / embed_dim, r = i - head * embed_dim;
        const int64_t ix = int64_t(token) * stride + head * head_size + r;
        const int64_t iy = ix + embed_dim;
        const __half c = cache[r], s = cache[embed_dim + r];
        const __half qx = query[ix], qy = query[iy];
        query[ix] = rn_sub(rn_mul(qx, c), rn_mul(qy, s));
        query[iy] = rn_add(rn_mul(qy, c), rn_mul(qx, s));
        const __half kx = key[ix], ky = key[iy];
        key[ix] = rn_sub(rn_mul(kx, c), rn_mul(ky, s));
        key[iy] = rn_add(rn_mul(ky, c), rn_mul(kx, s));
    }
}

// Prefill glue: for T tokens starting at cache position p0, rotate q in place (HF rotate_half convention, cos/sin tables
// [max_pos][D/2]) and write rotated k and v into the head-major KV cache [heads][max_ctx][D].  qkv rows are q | k | v.
__global__ void __launch_bounds__(128) rope_kv_write_kernel(__half* __restrict__ qkv, int64_t ld, const __half* __restrict__ cos_t,
                                                             const __half* __restrict__ sin_t, __half* __restrict__ kcache,
                                                             __half* __restrict__ vcache, int heads, int D, int max_ctx, int p0)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const int token = blockIdx.x, head = blockIdx.y;
    const int H     = heads * D;
    const int half  = D / 2;
    __half* row     = qkv + int64_t(token) * ld;
    const int pos   = p0 + token;
    for (int t = threadIdx.x; t < half; t += blockDim.x) {
        const float c = __half2float(cos_t[int64_t(pos) * half + t]), s = __half2float(sin_t[int64_t(pos) * half + t]);
        const int a = head * D + t, b = a + half;
        const float q0 = __half2float(row[a]), q1 = __half2float(row[b]);
        row[a] = __hadd(__float2half_rn(q0 * c), __float2half_rn(-q1 * s));
        row[b] = __hadd(__float2half_rn(q1 * c), __float2half_rn(q0 * s));
        const float k0 = __half2float(row[H + a]), k1 = __half2float(row[H + b]);
        const __half r0 = __hadd(__float2half_rn(k0 * c), __float2half_rn(-k1 * s));
        const __half r1 = __hadd(__float2half_rn(k1 * c), __float2half_rn(k0 * s));
        row[H + a] = r0;
        row[H + b] = r1;
        const int64_t crow = (int64_t(head) * max_ctx + pos) * D;
        kcache[crow + t] = r0;
        kcache[crow + t + half] = r1;
        vcache[crow + t]        = row[2 * H + a];
        vcache[crow + t + half] = row[2 * H + b];
    }
}

// Prefill glue: act[t][i] = fp16(silu(g)) * u.  interleaved = 0: gu rows are gate[I] | up[I]; 1: (g0,u0,g1,u1,...)
__global__ void __launch_bounds__(256) silu_mul_kernel(const __half* __restrict__ gu, int64_t ldg, __half* __restrict__ act,
                                                        int64_t lda, int I, int interleaved)
{
    pdl_launch_dependents();
    pdl_wait_prior_grids();
    const __half* row = gu + int64_t(blockIdx.y) * ldg;
    __half* out       = act + int64_t(blockIdx.y) * lda;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < I; i += gridDim.x * blockDim.x) {
        const __half g = interleaved ? row[2 * i] : row[i];
        const __half u = interleaved ? row[2 * i + 1] : row[I + i];
        const float gf = __half2float(g);
        out[i]         = __hmul(__float2half_rn(gf / (1.f + __expf(-gf))), u);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Decode attention: fused RoPE + KV append + attention for ONE query token, ONE launch per layer, no global scratch.
//   grid = (local heads, 4), cluster (1, 4, 1), 256 threads, head_dim 128.  The 4 CTAs of a cluster share one head.  The cache is
//   head-major, so a chunk of positions is ONE contiguous byte range of K and one of V: an elected thread moves them with two bulk
//   copies (TMA, cp.async.bulk) into shared memory and an mbarrier counts the bytes -- no registers and no load/store-unit slots
//   are tied up by the stream, the CTA stays small (72 registers) and the next GEMV's CTAs fit beside it.  A head's 8 staging
//   buffers (4 CTAs x 2) are sized from the position (read ahead of the dependency wait) so that they cover the whole context in
//   ONE pass up to 1280 positions; longer contexts loop (chunk c -> CTA c mod 4).  The copies are issued BEFORE
//   griddepcontrol.wait and overlap the tail of the q|k|v GEMV.
//   Thread (rowlane = t / 16, sub = t % 16) owns 8 dims of the rows of its row lane; each 16-lane group runs its own online
//   softmax with FHFMA products (no conversions, no block barriers inside the stream); the 16 row lanes are merged through shared
//   memory, the 4 CTAs through DISTRIBUTED shared memory: partials are pushed with st.async and counted by the receiver's
//   mbarrier (no cluster barrier, no release fence, no global scratch on the critical path).
//   The CTA whose chunk holds the newest position rotates k, appends k / v to the cache and uses them from shared memory.
//   Output: plain fp16 vector, or LL words pushed to every rank (head-sharded attention, see the file header).
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATT_D       = 128;
constexpr int ATT_THREADS = 256;
constexpr int ATT_LANES   = ATT_THREADS / 16;  // row lanes: 16-thread groups, each walks its own cache rows (a thread holds 8 dims)
constexpr int ATT_SPLITS  = 4;                 // CTAs per head == cluster size
constexpr int ATT_OWN     = ATT_D / ATT_SPLITS;  // output dims finished by one CTA of the cluster
constexpr int ATT_RMAX    = 10;  // cache rows per row lane in one staging buffer (a chunk = ATT_LANES x rpl <= 160 positions)
constexpr int ATT_BUF_HALF = ATT_LANES * ATT_RMAX * ATT_D * 2;  // bytes of K (or V) in one buffer: 40 KB
constexpr int ATT_SMEM     = 2 * 2 * ATT_BUF_HALF + 128;        // two buffers of K|V + alignment slack: 160 KB + 128

// acc(fp32) += a(fp16) * b(fp16): one FHFMA on sm_100a (exact product, single fp32 rounding) -- no fp16 -> fp32 conversions
__device__ __forceinline__ float fhfma(uint32_t a2, uint32_t b2, int hi, float acc)
{
    const uint16_t a = hi ? uint16_t(a2 >> 16) : uint16_t(a2 & 0xffffu);
    const uint16_t b = hi ? uint16_t(b2 >> 16) : uint16_t(b2 & 0xffffu);
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc) : "h"(a), "h"(b));
    return acc;
}
// 8 fp16 cache values x 8 fp16 query values (packed pairs), fp32 accumulation, two independent chains
__device__ __forceinline__ float dot8(const uint4& kv, const uint32_t (&q)[4])
{
    float a = 0.f, b = 0.f;
    a = fhfma(kv.x, q[0], 0, a); b = fhfma(kv.x, q[0], 1, b);
    a = fhfma(kv.y, q[1], 0, a); b = fhfma(kv.y, q[1], 1, b);
    a = fhfma(kv.z, q[2], 0, a); b = fhfma(kv.z, q[2], 1, b);
    a = fhfma(kv.w, q[3], 0, a); b = fhfma(kv.w, q[3], 1, b);
    return a + b;
}
// acc[0..7] += p * v[0..7], p as fp16 (the usual flash-attention choice for the P.V product), fp32 accumulation
__device__ __forceinline__ void axpy8(uint32_t p2, const uint4& vv, float (&acc)[8])
{
    acc[0] = fhfma(vv.x, p2, 0, acc[0]); acc[1] = fhfma(vv.x >> 16, p2, 0, acc[1]);
    acc[2] = fhfma(vv.y, p2, 0, acc[2]); acc[3] = fhfma(vv.y >> 16, p2, 0, acc[3]);
    acc[4] = fhfma(vv.z, p2, 0, acc[4]); acc[5] = fhfma(vv.z >> 16, p2, 0, acc[5]);
    acc[6] = fhfma(vv.w, p2, 0, acc[6]); acc[7] = fhfma(vv.w >> 16, p2, 0, acc[7]);
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank)
{
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr));
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// Remote shared-memory store that also signals: writes v into CTA `rank`'s copy of *local_ptr and completes 4 transaction bytes on
// that CTA's copy of *local_bar (st.async + mbarrier complete_tx) -- data and "it has arrived" travel together, so the receiver
// waits on its own mbarrier and no cluster-wide barrier or release fence sits on the critical path.
__device__ __forceinline__ void st_dsmem_f32_signal(float* local_ptr, unsigned long long* local_bar, uint32_t rank, float v)
{
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr));
    const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(local_bar));
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(b), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb)
                 : "memory");
}

struct AttnOut {
    __half* out;       // plain output [heads_local * 128] (world == 1) or nullptr
    LLPush push;       // LL output: every rank's attention vector, this rank's heads at element offset push.elem_off
    NextHint next;     // w != nullptr: the GEMV that consumes the attention output; its per-CTA head rows are requested into L2
};

__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const __half* __restrict__ qkv, const __half* __restrict__ cos_t,
                                                                   const __half* __restrict__ sin_t, const int* __restrict__ pos_p,
                                                                   __half* __restrict__ kcache, __half* __restrict__ vcache,
                                                                   int H, int max_ctx, float scale, const AttnOut ao)
{
    __shared__ __align__(16) __half q_s[ATT_D];  // rotated query, fp16 (HF evaluates RoPE in the model dtype)
    __shared__ __align__(16) __half knew[ATT_D];
    __shared__ __align__(16) __half vnew[ATT_D];
    __shared__ float grp_o[ATT_LANES][ATT_D];   // per row-lane partial numerators
    __shared__ float grp_ml[ATT_LANES][2];      // per row lane (max, denominator)
    __shared__ float mrg_o[ATT_SPLITS][ATT_OWN];  // partial numerators of MY output dims, one row pushed by every CTA of the cluster
    __shared__ float mrg_ml[ATT_SPLITS][2];  // (max, denominator) of every CTA of the cluster
    __shared__ __align__(8) unsigned long long mrg_bar;  // completes when all 8 x (16 + 2) floats above have landed
    __shared__ __align__(8) unsigned long long kv_bar[2];  // one per staging buffer: counts the bytes of its K and V bulk copies

    const int t       = threadIdx.x;
    const int sub     = t & 15;   // which 8-dim slice of the head
    const int rowlane = t >> 4;   // 0 .. ATT_LANES-1
    const int head    = blockIdx.x;
    const int split   = int(cluster_ctarank());  // == blockIdx.y

    trace_ev(TRACE_ATTN, 0);
    pdl_launch_dependents();
    if (t == 0) {
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mrg_bar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(uint32_t(ATT_SPLITS * (ATT_OWN + 2) * 4)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&kv_bar[0]))) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&kv_bar[1]))) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    cluster_arrive();  // "my mbarrier is armed": matched by the wait just before the first remote store (never blocks in practice)
    // *pos was advanced by the PREVIOUS step's last kernel, which completed before this step's first kernel released its
    // dependents (embed_kernel): the position (and below its rotary row) can be fetched ahead of the dependency wait.
    const int pos = *pos_p;
    // Chunk geometry: the 8 CTAs x 2 staging buffers of a head are 16 chunks in flight; a chunk holds 8 rowlanes x rpl positions,
    // rpl chosen so that the 16 chunks cover the whole context in ONE pass while it fits (<= 1280 positions) -- every CTA then does
    // the same work and nothing waits for a second, serialised round of loads.  Longer contexts loop (chunk c -> CTA c mod 8).
    constexpr int kSlots = 2 * ATT_SPLITS * ATT_LANES;  // row slots per round of the whole cluster (128)
    const int rpl        = min(ATT_RMAX, max(1, (pos + kSlots) / kSlots));
    const int chunk_rows = ATT_LANES * rpl;
    // The rows of a chunk are CONTIGUOUS in the head-major cache: one elected thread moves a chunk's K rows and V rows with two bulk
    // copies (TMA) into shared memory and an mbarrier counts the bytes.  No registers and no load/store-unit slots are tied up by the
    // stream, so the CTA stays small (the next GEMV's CTAs fit beside it and start THEIR weight stream during the attention).
    // Rows below the position were written by earlier STEPS (row `pos` itself is appended by this launch and taken from shared
    // memory), so the copies need nothing the preceding kernel produces.
    extern __shared__ uint8_t att_dyn[];
    const uint32_t kv_base = (static_cast<uint32_t>(__cvta_generic_to_shared(att_dyn)) + 127u) & ~127u;
    auto issue_chunk = [&](int buf, int chunk) {  // thread 0 only
        const int p0     = chunk * chunk_rows;
        const int n      = min(chunk_rows, pos - p0);  // rows strictly below the position
        const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&kv_bar[buf]));
        const uint32_t bytes = n > 0 ? uint32_t(n) * ATT_D * 2 : 0u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2u * bytes) : "memory");
        if (n > 0) {
            const __half* ks = kcache + (int64_t(head) * max_ctx + p0) * ATT_D;
            const __half* vs = vcache + (int64_t(head) * max_ctx + p0) * ATT_D;
            const uint32_t kd = kv_base + uint32_t(buf) * 2u * ATT_BUF_HALF;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kd), "l"(ks),
                         "r"(bytes), "r"(b)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kd + ATT_BUF_HALF),
                         "l"(vs), "r"(bytes), "r"(b)
                         : "memory");
        }
    };
    if (t == 0) {
        issue_chunk(0, split);
        issue_chunk(1, split + ATT_SPLITS);
    }
    // the attention stream is small (the layer's KV rows): use the idle HBM time to pull the head of the next GEMV's weights into L2
    if (ao.next.w != nullptr && t == ATT_THREADS - 32)
        l2_prefetch_next(ao.next, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y);
    float rope_c = 0.f, rope_s = 0.f;
    if (t < ATT_D / 2) {
        rope_c = __half2float(cos_t[int64_t(pos) * (ATT_D / 2) + t]);
        rope_s = __half2float(sin_t[int64_t(pos) * (ATT_D / 2) + t]);
    }
    trace_ev(TRACE_ATTN, 1);
    pdl_wait_prior_grids();  // qkv of the current token comes from the preceding GEMV
    trace_ev(TRACE_ATTN, 2);
    const int new_chunk = pos / chunk_rows;
    const bool owns_new = (new_chunk % ATT_SPLITS) == split;

    // RoPE of q (every CTA) and of the new k (owner CTA); HF rotate_half convention evaluated in fp16
    if (t < ATT_D / 2) {
        const float c = rope_c, s = rope_s;
        const int a = head * ATT_D + t, b = a + ATT_D / 2;
        const float q0 = __half2float(qkv[a]), q1 = __half2float(qkv[b]);
        q_s[t]             = __hadd(__float2half_rn(q0 * c), __float2half_rn(-q1 * s));
        q_s[t + ATT_D / 2] = __hadd(__float2half_rn(q1 * c), __float2half_rn(q0 * s));
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

    uint32_t qf[4];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(&q_s[sub * 8]);
        qf[0] = qv.x; qf[1] = qv.y; qf[2] = qv.z; qf[3] = qv.w;
    }
    trace_ev(TRACE_ATTN, 3);

    // online softmax state of this 16-lane group (identical in all 16 lanes; each lane owns 8 dims of the numerator)
    float m_run = -INFINITY, l_run = 0.f;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto wait_buf = [&](int buf, uint32_t parity) {
        const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(&kv_bar[buf]));
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(b), "r"(parity)
                : "memory");
        }
    };
    auto lds_row = [&](uint32_t base, int row) -> uint4 {  // this thread's 8 dims of a staged row (a warp reads 2 whole rows: no conflicts)
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + uint32_t(row) * (ATT_D * 2) + uint32_t(sub) * 16u));
        return v;
    };
    auto process = [&](int buf, int chunk) {
        const int p0       = chunk * chunk_rows;
        const uint32_t kb  = kv_base + uint32_t(buf) * 2u * ATT_BUF_HALF;
        const uint32_t vb  = kb + ATT_BUF_HALF;
        float sc[ATT_RMAX];
        float cmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < ATT_RMAX; ++i) {
            const int r = i * ATT_LANES + rowlane;
            const int j = p0 + r;
            float d     = 0.f;
            if (i < rpl && j <= pos) {  // uniform inside the 16-lane group
                const uint4 kv = (j == pos) ? *reinterpret_cast<const uint4*>(&knew[sub * 8])  // the row appended by this very launch
                                            : lds_row(kb, r);
                d = dot8(kv, qf);
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1)
                d += __shfl_xor_sync(0xffffffffu, d, o);
            sc[i] = (i < rpl && j <= pos) ? d * scale : -INFINITY;
            cmax  = fmaxf(cmax, sc[i]);
        }
        if (cmax == -INFINITY)
            return;  // nothing valid for this group in this chunk (uniform inside the 16-lane group)
        const float m_new = fmaxf(m_run, cmax);
        const float resc  = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
        l_run *= resc;
#pragma unroll
        for (int d = 0; d < 8; ++d)
            acc[d] *= resc;
#pragma unroll
        for (int i = 0; i < ATT_RMAX; ++i) {
            const int r = i * ATT_LANES + rowlane;
            const int j = p0 + r;
            if (sc[i] != -INFINITY) {
                // the probability is rounded to fp16 once and that SAME value feeds numerator and denominator
                const __half ph = __float2half_rn(__expf(sc[i] - m_new));
                l_run += __half2float(ph);
                const uint4 vv = (j == pos) ? *reinterpret_cast<const uint4*>(&vnew[sub * 8]) : lds_row(vb, r);
                axpy8(uint32_t(__half_as_ushort(ph)), vv, acc);
            }
        }
        m_run = m_new;
    };

    {
        int it = 0;
        for (int chunk = split; chunk * chunk_rows <= pos; chunk += ATT_SPLITS, ++it) {
            const int buf = it & 1;
            wait_buf(buf, uint32_t(it >> 1) & 1u);
            process(buf, chunk);
            const int nxt = chunk + 2 * ATT_SPLITS;
            if (nxt * chunk_rows <= pos) {  // contexts beyond one pass: refill this buffer once every thread has consumed it
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (t == 0)
                    issue_chunk(buf, nxt);
            }
        }
    }

    trace_ev(TRACE_ATTN, 4);
    // merge the row lanes of this CTA
#pragma unroll
    for (int d = 0; d < 8; ++d)
        grp_o[rowlane][sub * 8 + d] = acc[d];
    if (sub == 0) {
        grp_ml[rowlane][0] = m_run;
        grp_ml[rowlane][1] = l_run;
    }
    __syncthreads();
    if (t < ATT_D) {
        float mm = -INFINITY;
#pragma unroll
        for (int r = 0; r < ATT_LANES; ++r)
            mm = fmaxf(mm, grp_ml[r][0]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int r = 0; r < ATT_LANES; ++r) {
            const float mr = grp_ml[r][0];
            const float w  = (mr == -INFINITY) ? 0.f : __expf(mr - mm);
            num = fmaf(w, grp_o[r][t], num);
            den = fmaf(w, grp_ml[r][1], den);
        }
        // push: dim t is finished by CTA t / ATT_OWN of the cluster; every CTA needs every (max, denominator).  (The wait completes
        // the start-of-kernel phase: every sibling has armed its mbarrier -- it never blocks in practice.)
        cluster_wait();
        st_dsmem_f32_signal(&mrg_o[split][t % ATT_OWN], &mrg_bar, uint32_t(t / ATT_OWN), num);
        if (t < ATT_SPLITS) {
            st_dsmem_f32_signal(&mrg_ml[split][0], &mrg_bar, uint32_t(t), mm);
            st_dsmem_f32_signal(&mrg_ml[split][1], &mrg_bar, uint32_t(t), den);
        }
    }
    else {
        cluster_wait();  // every thread that arrived at the start-of-kernel phase also completes it
    }
    // The finishing threads wait on THIS CTA's mbarrier until all ATT_SPLITS x (ATT_OWN + 2) floats addressed to it have landed; the
    // other threads are done (nobody reads a sibling's shared memory, and every store's target is kept alive by this very wait).
    if (t < ATT_OWN) {
        const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&mrg_bar));
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
    trace_ev(TRACE_ATTN, 5);
    if (t < ATT_OWN) {
        const int d = split * ATT_OWN + t;
        float mm = -INFINITY;
#pragma unroll
        for (int r = 0; r < ATT_SPLITS; ++r)
            mm = fmaxf(mm, mrg_ml[r][0]);
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int r = 0; r < ATT_SPLITS; ++r) {
            const float mr = mrg_ml[r][0];
            const float w  = (mr == -INFINITY) ? 0.f : __expf(mr - mm);
            num = fmaf(w, mrg_o[r][t], num);
            den = fmaf(w, mrg_ml[r][1], den);
        }
        const __half o = __float2half_rn(num / den);
        if (ao.out != nullptr) {
            ao.out[head * ATT_D + d] = o;
        }
        else {
            // LL: lanes pair up (even lane carries dims d, d + 1) and push one 8-byte word per rank
            const uint32_t mine  = uint32_t(__half_as_ushort(o));
            const uint32_t other = __shfl_down_sync(0xffffffffu, mine, 1);
            if ((t & 1) == 0)
                ll_push_word(ao.push, (head * ATT_D + d) >> 1, mine | (other << 16));
        }
    }
    trace_ev(TRACE_ATTN, 6);
}

// ---------------------------------------------------------------------------------------------------------------
// lm_head + greedy arg-max: logits[v] = sum_k RMSNorm(x)[k] * W[v, k] over this rank's vocabulary rows, fp16 weights
// (the reference leaves lm_head unquantised, quantizer.py:40), fused final RMSNorm on the activation load, fp32
// accumulation (FHFMA), block arg-max, and a self-resetting ticket so that the LAST CTA picks the winner, (multi-GPU:
// exchanges candidates with the other ranks through LL words), writes the next token id, advances the position and
// the step counter.  One launch replaces RMSNorm + cuBLAS GEMV + arg-max + two copies.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LM_THREADS = 256;
constexpr int LM_R       = 4;   // vocabulary rows per register group
constexpr int LM_MAXKI   = 4;   // hidden <= 8192

struct LmArgs {
    const __half* x;          // [H] plain, or LL words (x_ll.tag_base != nullptr)
    LLTag x_ll;
    const __half* norm_w;     // [H]
    float eps;
    const __half* w;          // [V_local][H]
    int V_local, H, v_begin;  // this rank's first vocabulary row
    __half* logits;           // optional [V_total]: this rank writes its slice (tests / callers that want logits)
    float* cta_val;           // [grid] scratch
    int* cta_idx;             // [grid]
    unsigned* ticket;         // zero between launches
    int64_t* token;           // out: next token id
    int* pos;                 // += 1
    int* step;                // += 1 (the LL tag base)
    LLPush cand;              // multi-GPU: candidate exchange buffer (2 words per rank), world == 1: unused
    int world, rank;
};

template <int KITERS>
__global__ void __launch_bounds__(LM_THREADS, 2) lm_head_argmax_kernel(const LmArgs a)
{
    extern __shared__ float lm_partial[];  // [rows][8 warps]
    __shared__ float red_s[8];
    __shared__ float best_v[8];
    __shared__ int best_i[8];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, nchunks = H >> 3;  // 16-byte chunks of 8 fp16 weights
    const int row_begin = int((int64_t(blockIdx.x) * a.V_local) / gridDim.x);
    const int row_end   = int((int64_t(blockIdx.x + 1) * a.V_local) / gridDim.x);
    const int nrows     = row_end - row_begin;
    const int ngroups   = (nrows + LM_R - 1) / LM_R;

    trace_ev(TRACE_LMHEAD, 0);
    pdl_launch_dependents();
    uint4 wb[2][LM_R][KITERS];
    auto load_group = [&](uint4 (&buf)[LM_R][KITERS], int g) {
#pragma unroll
        for (int r = 0; r < LM_R; ++r) {
            const int row = row_begin + g * LM_R + r;
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const int c = tid + i * LM_THREADS;
                buf[r][i]   = (row < row_end && c < nchunks) ? ldg_stream_128(a.w + int64_t(row) * H + int64_t(c) * 8)
                                                             : make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    if (ngroups > 0) load_group(wb[0], 0);
    if (ngroups > 1) load_group(wb[1], 1);
    trace_ev(TRACE_LMHEAD, 1);
    pdl_wait_prior_grids();
    trace_ev(TRACE_LMHEAD, 2);

    // activation slice: 8 values per chunk, final RMSNorm applied in registers (HF arithmetic)
    uint32_t xh[KITERS][4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < KITERS; ++i) {
        const int c = tid + i * LM_THREADS;
        if (c < nchunks) {
            if (a.x_ll.tag_base != nullptr)
                ll_load_words<4>(reinterpret_cast<const unsigned long long*>(a.x) + int64_t(c) * 4, ll_tag(a.x_ll), xh[i]);
            else {
                const uint4 v = *reinterpret_cast<const uint4*>(a.x + int64_t(c) * 8);
                xh[i][0] = v.x; xh[i][1] = v.y; xh[i][2] = v.z; xh[i][3] = v.w;
            }
        }
        else {
            xh[i][0] = xh[i][1] = xh[i][2] = xh[i][3] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xh[i][j]));
            ss = fmaf(f.x, f.x, ss);
            ss = fmaf(f.y, f.y, ss);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red_s[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int wi = 0; wi < 8; ++wi)
        tot += red_s[wi];
    const float rn = rsqrtf(tot / float(H) + a.eps);
#pragma unroll
    for (int i = 0; i < KITERS; ++i) {
        const int c = tid + i * LM_THREADS;
        if (c < nchunks) {
            const uint4 nw = *reinterpret_cast<const uint4*>(a.norm_w + int64_t(c) * 8);
            const uint32_t nwr[4] = {nw.x, nw.y, nw.z, nw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f  = __half22float2(*reinterpret_cast<const __half2*>(&xh[i][j]));
                const __half2 n = __floats2half2_rn(f.x * rn, f.y * rn);
                const __half2 o = __hmul2(n, *reinterpret_cast<const __half2*>(&nwr[j]));
                xh[i][j]        = *reinterpret_cast<const uint32_t*>(&o);
            }
        }
    }

    auto compute_group = [&](uint4 (&buf)[LM_R][KITERS], int g) {
        float acc[LM_R];
#pragma unroll
        for (int r = 0; r < LM_R; ++r) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < KITERS; ++i) {
                const uint32_t wv[4] = {buf[r][i].x, buf[r][i].y, buf[r][i].z, buf[r][i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(s) : "h"(uint16_t(wv[j] & 0xffffu)), "h"(uint16_t(xh[i][j] & 0xffffu)));
                    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(s) : "h"(uint16_t(wv[j] >> 16)), "h"(uint16_t(xh[i][j] >> 16)));
                }
            }
            acc[r] = s;
        }
#pragma unroll
        for (int r = 0; r < LM_R; ++r) {
            float v = acc[r];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1)
                v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0)
                lm_partial[(g * LM_R + r) * 8 + warp] = v;
        }
    };
    for (int g = 0; g < ngroups; g += 2) {
        compute_group(wb[0], g);
        if (g + 2 < ngroups) load_group(wb[0], g + 2);
        if (g + 1 < ngroups) {
            compute_group(wb[1], g + 1);
            if (g + 3 < ngroups) load_group(wb[1], g + 3);
        }
    }
    __syncthreads();

    // logits of this CTA's rows (rounded to fp16 like the framework's matmul output) and the block arg-max
    float bv = -INFINITY;
    int bi   = 0x7fffffff;
    for (int r = tid; r < nrows; r += LM_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi)
            s += lm_partial[r * 8 + wi];
        const __half lh = __float2half_rn(s);
        const int v     = a.v_begin + row_begin + r;
        if (a.logits != nullptr)
            a.logits[v] = lh;
        const float lv = __half2float(lh);
        if (lv > bv || (lv == bv && v < bi)) {
            bv = lv;
            bi = v;
        }
    }
    auto better = [](float v1, int i1, float v2, int i2) { return v1 > v2 || (v1 == v2 && i1 < i2); };
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi   = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        best_v[warp] = bv;
        best_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
        for (int wi = 1; wi < 8; ++wi)
            if (better(best_v[wi], best_i[wi], bv, bi)) {
                bv = best_v[wi];
                bi = best_i[wi];
            }
        a.cta_val[blockIdx.x] = bv;
        a.cta_idx[blockIdx.x] = bi;
        __threadfence();
        const unsigned old = atomicAdd(a.ticket, 1u);
        is_last            = (old == gridDim.x - 1) ? 1 : 0;
        if (is_last)
            *a.ticket = 0;  // self-cleaning
    }
    __syncthreads();
    if (!is_last)
        return;
    // last CTA: winner over all CTAs of this rank (first maximum, like torch.argmax)
    __threadfence();
    bv = -INFINITY;
    bi = 0x7fffffff;
    for (int c = tid; c < int(gridDim.x); c += LM_THREADS) {
        const float v = __ldcg(a.cta_val + c);
        const int i   = __ldcg(a.cta_idx + c);
        if (better(v, i, bv, bi)) {
            bv = v;
            bi = i;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi   = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        best_v[warp] = bv;
        best_i[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
        for (int wi = 1; wi < 8; ++wi)
            if (better(best_v[wi], best_i[wi], bv, bi)) {
                bv = best_v[wi];
                bi = best_i[wi];
            }
        if (a.world > 1) {
            // publish this rank's candidate to every rank (words 2 r and 2 r + 1 of the candidate buffer), then gather
            const uint32_t tag = ll_tag(a.cand.tag);
            ll_push_word(a.cand, 2 * a.rank, __float_as_uint(bv));
            ll_push_word(a.cand, 2 * a.rank + 1, uint32_t(bi));
            const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(a.cand.local);
            for (int r = 0; r < a.world; ++r) {
                uint32_t w2[2];
                ll_load_words<2>(mine + 2 * r, tag, w2);
                const float v = __uint_as_float(w2[0]);
                const int i   = int(w2[1]);
                if (better(v, i, bv, bi)) {
                    bv = v;
                    bi = i;
                }
            }
        }
        *a.token = int64_t(bi);
        *a.pos += 1;
        *a.step += 1;
    }
}

}  // namespace

EB_TRACE_SETTER(trace_set_decode)

}  // namespace eetq_b200

using namespace eetq_b200;

namespace {
int fill_push(LLPush& p, int world, const uint64_t* peers, void* local, int64_t elem_off, const void* step, int per_step, int index)
{
    p = LLPush{};
    p.world = world;
    for (int r = 0; r < world && r < 8; ++r)
        p.peer[r] = peers[r];
    p.local       = static_cast<unsigned long long*>(local);
    p.elem_off    = int(elem_off);
    p.tag.tag_base = static_cast<const int*>(step);
    p.tag.per_step = per_step;
    p.tag.index    = index;
    return EETQ_B200_OK;
}
}  // namespace

extern "C" {

int eetq_b200_decode_embed(const void* table, const void* token_i64, void* x, int64_t H, const eetq_b200_ll* x_ll, int pdl, void* stream)
{
    EB_CHECK_ARG(table && token_i64 && x && H % 8 == 0, "decode_embed: bad argument");
    if (int rc = check_arch())
        return rc;
    LLTag tag{};
    if (x_ll != nullptr && x_ll->step != nullptr) {
        tag.tag_base = static_cast<const int*>(x_ll->step);
        tag.per_step = x_ll->per_step;
        tag.index    = x_ll->index;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned((H / 8 + 255) / 256)), dim3(256), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, embed_kernel, static_cast<const __half*>(table), static_cast<const int64_t*>(token_i64),
                                     static_cast<__half*>(x), int(H), tag));
    count_launch();
    return EETQ_B200_OK;
}

int eetq_b200_rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int64_t M, int64_t H, float eps, int pdl, void* stream)
{
    EB_CHECK_ARG(x && w && y && M > 0 && H > 0 && ldx >= H && ldy >= H, "rmsnorm: bad argument");
    if (int rc = check_arch())
        return rc;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(M)), dim3(512), 0, pdl != 0, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_kernel, static_cast<const __half*>(x), ldx, static_cast<const __half*>(w),
                                     static_cast<__half*>(y), ldy, int(H), eps, 0));
    count_launch();
    return EETQ_B200_OK;
}

// layernorm_forward of the reference (csrc/eetpy.cpp:19 -> layernorm.cu:88-110): fp16 [m, n] rows, T5-style RMS norm
int eetq_b200_layernorm_forward(const void* input, const void* gamma, void* out, int64_t m, int64_t n, float eps, void* stream)
{
    EB_CHECK_ARG(input && gamma && out && m > 0 && n > 0, "layernorm_forward: bad argument");
    if (int rc = check_arch())
        return rc;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(m)), dim3(512), 0, false, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rmsnorm_kernel, static_cast<const __half*>(input), n, static_cast<const __half*>(gamma),
                                     static_cast<__half*>(out), n, int(n), eps, 1));
    count_launch();
    return EETQ_B200_OK;
}

// rotary_embedding_neox of the reference (csrc/eetpy.cpp:18 -> pos_encoding_kernels.cu:55-87): in place on query and key
int eetq_b200_rotary_embedding_neox(const void* positions_i64, void* query, void* key, int64_t num_tokens, int64_t num_heads,
                                    int64_t head_size, const void* cos_sin_cache, int64_t rot_dim, void* stream)
{
    EB_CHECK_ARG(positions_i64 && query && key && cos_sin_cache, "rotary_embedding_neox: null pointer argument");
    EB_CHECK_ARG(num_tokens > 0 && num_heads > 0 && head_size > 0 && rot_dim > 0 && rot_dim % 2 == 0 && rot_dim <= head_size,
                 "rotary_embedding_neox: bad shape");
    if (int rc = check_arch())
        return rc;
    const int threads = int(num_heads * rot_dim / 2 < 512 ? num_heads * rot_dim / 2 : 512);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(num_tokens)), dim3(unsigned(threads < 32 ? 32 : threads)), 0, false, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rope_neox_kernel, static_cast<const int64_t*>(positions_i64), static_cast<__half*>(query),
                                     static_cast<__half*>(key), static_cast<const __half*>(cos_sin_cache), int(rot_dim),
                                     int(num_heads * head_size), int(num_heads), int(head_size)));
    count_launch();
    return EETQ_B200_OK;
}

int eetq_b200_prefill_rope_kv(void* qkv, int64_t ld, const void* cos_t, const void* sin_t, void* kcache, void* vcache, int64_t T,
                              int64_t heads, int64_t D, int64_t max_ctx, int64_t p0, void* stream)
{
    EB_CHECK_ARG(qkv && cos_t && sin_t && kcache && vcache && T > 0 && heads > 0 && D > 0 && D % 2 == 0, "prefill_rope_kv: bad argument");
    EB_CHECK_ARG(p0 >= 0 && p0 + T <= max_ctx && ld >= 3 * heads * D, "prefill_rope_kv: positions exceed the cache or bad stride");
    if (int rc = check_arch())
        return rc;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(T), unsigned(heads)), dim3(128), 0, false, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rope_kv_write_kernel, static_cast<__half*>(qkv), ld, static_cast<const __half*>(cos_t),
                                     static_cast<const __half*>(sin_t), static_cast<__half*>(kcache), static_cast<__half*>(vcache),
                                     int(heads), int(D), int(max_ctx), int(p0)));
    count_launch();
    return EETQ_B200_OK;
}

int eetq_b200_silu_mul(const void* gu, int64_t ldg, void* act, int64_t lda, int64_t T, int64_t I, int interleaved, void* stream)
{
    EB_CHECK_ARG(gu && act && T > 0 && I > 0 && ldg >= 2 * I && lda >= I, "silu_mul: bad argument");
    if (int rc = check_arch())
        return rc;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned((I + 255) / 256), unsigned(T)), dim3(256), 0, false, static_cast<cudaStream_t>(stream));
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, silu_mul_kernel, static_cast<const __half*>(gu), ldg, static_cast<__half*>(act), lda, int(I),
                                     interleaved));
    count_launch();
    return EETQ_B200_OK;
}

// Fused RoPE + KV append + attention for one token at position *pos (reads rows [0, pos], writes cache row pos).
//   qkv [3 * H_local] raw projections of this rank's heads (q | k | v, not modified); kcache/vcache [H_local/D][max_ctx][D]
//   (head-major); out [H_local] plain fp16, or NULL with `push` describing the LL exchange of the full attention vector.
int eetq_b200_decode_attention(const void* qkv, const void* cos_t, const void* sin_t, const void* pos_i32, void* kcache,
                               void* vcache, void* out, int64_t H_local, int64_t D, int64_t max_ctx, const eetq_b200_ll_push* push,
                               const void* next_w, int64_t next_n, int64_t next_k, int pdl, void* stream)
{
    EB_CHECK_ARG(next_w == nullptr || ((reinterpret_cast<uintptr_t>(next_w) & 15u) == 0 && next_n > 0 && next_k > 0 && next_k % 64 == 0),
                 "decode_attention: bad next_w hint");
    EB_CHECK_ARG(qkv && cos_t && sin_t && pos_i32 && kcache && vcache, "decode_attention: null pointer argument");
    EB_CHECK_ARG((out != nullptr) != (push != nullptr), "decode_attention: exactly one of out / push must be given");
    EB_CHECK_ARG(D == ATT_D && H_local % D == 0 && H_local > 0, "decode_attention: head_dim must be 128");
    EB_CHECK_ARG(max_ctx >= 1 && max_ctx <= (1 << 20), "decode_attention: bad max_ctx");
    if (int rc = check_arch())
        return rc;
    AttnOut ao{};
    ao.out = static_cast<__half*>(out);
    if (push != nullptr) {
        EB_CHECK_ARG(push->world >= 1 && push->world <= 8 && push->step != nullptr && push->peers != nullptr, "decode_attention: bad LL push");
        fill_push(ao.push, push->world, push->peers, push->local, push->elem_off, push->step, push->per_step, push->index);
    }
    static const bool l2_next = [] {
        const char* e = getenv("EETQ_B200_L2_NEXT");
        return !(e != nullptr && e[0] == '0');
    }();
    if (l2_next && next_w != nullptr)
        ao.next = make_next_hint(next_w, next_n, next_k);
    const int heads = int(H_local / D);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    {
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            EB_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
            attr_set[dev] = true;
        }
    }
    launch_cfg(cfg, attr, dim3(unsigned(heads), ATT_SPLITS), dim3(ATT_THREADS), ATT_SMEM, pdl != 0, static_cast<cudaStream_t>(stream), ATT_SPLITS);
    EB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attn_decode_kernel, static_cast<const __half*>(qkv), static_cast<const __half*>(cos_t),
                                     static_cast<const __half*>(sin_t), static_cast<const int*>(pos_i32), static_cast<__half*>(kcache),
                                     static_cast<__half*>(vcache), int(H_local), int(max_ctx), 1.0f / sqrtf(float(D)), ao));
    count_launch();
    return EETQ_B200_OK;
}

// Development aid: how many CTAs / 8-CTA clusters of the decode attention kernel fit on the current device at once.
int eetq_b200_decode_attention_occupancy(int heads, int* ctas_per_sm, int* max_clusters)
{
    EB_CHECK_ARG(ctas_per_sm && max_clusters && heads > 0, "decode_attention_occupancy: bad argument");
    EB_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    EB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, attn_decode_kernel, ATT_THREADS, ATT_SMEM));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(heads), ATT_SPLITS), dim3(ATT_THREADS), ATT_SMEM, false, nullptr, ATT_SPLITS);
    EB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, attn_decode_kernel, &cfg));
    return EETQ_B200_OK;
}

size_t eetq_b200_lm_head_scratch_bytes(void)
{
    const DeviceInfo& di = device_info();
    const int grid       = (di.ok ? di.sm_count : 148) * 2;
    return size_t(grid) * 8 + 64;
}

// logits = RMSNorm(x; norm_w, eps) @ W^T over V_local rows, arg-max -> *token; *pos += 1; *step += 1.
//   scratch: eetq_b200_lm_head_scratch_bytes() bytes, ZERO on first use (left clean).  x_ll / cand: NULL on one GPU.
int eetq_b200_lm_head_argmax(const void* x, const eetq_b200_ll* x_ll, const void* norm_w, float eps, const void* w, int64_t V_local,
                             int64_t H, int64_t v_begin, void* logits, void* scratch, void* token_i64, void* pos_i32, void* step_i32,
                             const eetq_b200_ll_push* cand, int rank, int pdl, void* stream)
{
    EB_CHECK_ARG(x && norm_w && w && scratch && token_i64 && pos_i32 && step_i32, "lm_head_argmax: null pointer argument");
    EB_CHECK_ARG(V_local > 0 && H > 0 && H % 8 == 0 && H <= LM_MAXKI * LM_THREADS * 8, "lm_head_argmax: bad shape (hidden <= %d)",
                 LM_MAXKI * LM_THREADS * 8);
    if (int rc = check_arch())
        return rc;
    const DeviceInfo& di = device_info();
    EB_CHECK_ARG(di.ok, "lm_head_argmax: device query failed");
    int grid = di.sm_count * 2;
    if (grid > V_local) grid = int(V_local);
    LmArgs a{};
    a.x = static_cast<const __half*>(x);
    if (x_ll != nullptr && x_ll->step != nullptr) {
        a.x_ll.tag_base = static_cast<const int*>(x_ll->step);
        a.x_ll.per_step = x_ll->per_step;
        a.x_ll.index    = x_ll->index;
    }
    a.norm_w  = static_cast<const __half*>(norm_w);
    a.eps     = eps;
    a.w       = static_cast<const __half*>(w);
    a.V_local = int(V_local);
    a.H       = int(H);
    a.v_begin = int(v_begin);
    a.logits  = static_cast<__half*>(logits);
    uint8_t* s = static_cast<uint8_t*>(scratch);
    a.ticket  = reinterpret_cast<unsigned*>(s);
    a.cta_val = reinterpret_cast<float*>(s + 64);
    a.cta_idx = reinterpret_cast<int*>(s + 64 + size_t(di.sm_count) * 2 * 4);
    a.token   = static_cast<int64_t*>(token_i64);
    a.pos     = static_cast<int*>(pos_i32);
    a.step    = static_cast<int*>(step_i32);
    a.world   = 1;
    a.rank    = rank;
    if (cand != nullptr && cand->world > 1) {
        EB_CHECK_ARG(cand->world <= 8 && cand->step != nullptr && cand->peers != nullptr && cand->local != nullptr, "lm_head_argmax: bad LL push");
        fill_push(a.cand, cand->world, cand->peers, cand->local, 0, cand->step, cand->per_step, cand->index);
        a.world = cand->world;
    }
    const int max_rows  = int((V_local + grid - 1) / grid);
    const size_t smem   = size_t((max_rows + LM_R - 1) / LM_R * LM_R) * 8 * sizeof(float);
    const int kiters    = int((H / 8 + LM_THREADS - 1) / LM_THREADS);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    launch_cfg(cfg, attr, dim3(unsigned(grid)), dim3(LM_THREADS), smem, pdl != 0, static_cast<cudaStream_t>(stream));
    cudaError_t e;
    switch (kiters) {
        case 1: e = cudaLaunchKernelEx(&cfg, lm_head_argmax_kernel<1>, a); break;
        case 2: e = cudaLaunchKernelEx(&cfg, lm_head_argmax_kernel<2>, a); break;
        case 3: e = cudaLaunchKernelEx(&cfg, lm_head_argmax_kernel<3>, a); break;
        default: e = cudaLaunchKernelEx(&cfg, lm_head_argmax_kernel<4>, a); break;
    }
    count_launch();
    if (e != cudaSuccess) {
        set_error("lm_head_argmax launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

}  // extern "C"
