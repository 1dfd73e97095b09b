// gemm_tc_v1.cu -- ROUND-1 kernel, kept only as an A/B baseline behind EETQ_B200_TC_IMPL=v1 (see gemm_tc.cu for the product).
// batched / prefill w8a16 GEMM on 5th-generation tensor cores (tcgen05 + TMEM + TMA, sm_100a).
//
// Replaces the reference prefill path
//   CutlassFpAIntBGemmRunner<half,uint8_t>::gemm   /root/reference/csrc/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:441-553
//   GemmFpAIntB + DqMmaMultistage                  /root/reference/csrc/cutlass_extensions/include/cutlass_extensions/gemm/kernel/fpA_intB_gemm.h:60-486,
//                                                  .../gemm/threadblock/dq_mma_multistage.h:98-591
// (mma.sync + cp.async + ldmatrix, compiled out for sm >= 90) with a Blackwell-native kernel; nothing of the
// CUTLASS 2.x structure is kept.
//
// Formulation (operands swapped so that small token counts stay efficient and the per-channel scale is per
// accumulator ROW):      D[n, t] = sum_k  A[n, k] * B[t, k]
//      A = dequantised weight tile  [128 features x 64 k]  fp16/bf16, K-major, 128B-swizzled   (written by the
//          dequant warps from the int8 tile that TMA staged -- b200 layout rows are already K-major)
//      B = activation tile          [BT tokens   x 64 k]  fp16/bf16, K-major, 128B-swizzled   (TMA straight from x)
//      D = fp32 accumulator in TMEM: lane = feature, column = token   (UMMA M = 128, N = BT, K = 16)
//
// Warp roles (352 threads):  warps 0..7 = dequantisers (two groups of 4 on alternating k-blocks), then epilogue |
//                            warp 8 = weight TMA producer | warp 9 = tcgen05.mma issuer + TMEM alloc/dealloc |
//                            warp 10 = activation TMA producer   (single-thread roles on the highest warp ids)
// Pipelines (mbarrier):      wfull/wempty[WS]   : TMA  <-> dequant   (int8, 256 k-bytes per stage = 4 MMA k-blocks)
//                            xfull/xempty[XS]   : TMA  <-> MMA       (activation tile, 64 k)
//                            a_full/a_empty[4]  : dequant <-> MMA (fp16 A tile)
//                            tmem_full          : MMA -> epilogue
//
// Arithmetic.  fp16: A = fp16(fp16(q) * s) with ONE rounding per weight and fp32 accumulation -- exactly the
// reference's K1 arithmetic (mma_tensorop_dequantizer.h:259-274, default_fpA_intB_traits.h:110), so results
// match it up to fp32 summation order.  bf16 (extension): A = bf16(q) exactly, scale applied in the fp32 epilogue.
//
// Small M is HBM-bound and N/128 tiles do not fill 148 SMs, so K is split over `splits` CTAs per tile; partial
// tiles go through an fp32 workspace; the `splits` CTAs of a tile meet at a counter and each reduces its share of
// the token columns in split order (deterministic); the last to leave resets the counters, so the workspace needs
// zeroing only once (the reference disables split-K
// altogether by passing a null workspace, fpA_intB_gemm_wrapper.cu:169-170).
//
// Roofline: bytes = K*N + 2N + 2MK + 2MN, flops = 2MNK; HBM-bound for M <~ 140, tensor-bound above.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace eetq_b200 {

namespace {

constexpr int BLOCK_N      = 128;  // output features per CTA  (UMMA M)
constexpr int BLOCK_K      = 64;   // k per pipeline stage (64 fp16 = one 128-byte swizzle row)
constexpr int UMMA_K       = 16;
constexpr int NUM_A_STAGES = 4;   // fp16 A tiles in flight between the dequant groups and the MMA issuer
constexpr int DQ_GROUPS    = 2;   // dequant warps work as 2 groups of 4 warps on alternating k-blocks
constexpr int DQ_WARPS     = 8;
constexpr int DQ_THREADS   = DQ_WARPS * 32;
constexpr int TC_THREADS   = 64 + DQ_THREADS + 32;  // warp 0 weight TMA, warp 1 MMA, 8 dequant/epilogue warps, warp 10 activation TMA
// Role -> warp mapping: the single-thread roles get the HIGHEST warp ids (the SM's issue arbiter favours higher warp
// ids among eligible warps of a sub-partition, so the MMA issuer and the TMA producers are never starved by dequant warps)
constexpr int W_PRODUCER_WARP = DQ_WARPS;      // 8
constexpr int MMA_WARP        = DQ_WARPS + 1;  // 9
constexpr int X_PRODUCER_WARP = DQ_WARPS + 2;  // 10
constexpr int W8_TILE      = BLOCK_N * BLOCK_K;       // 8192 B of int8
constexpr int A_TILE       = BLOCK_N * BLOCK_K * 2;   // 16384 B of fp16/bf16

// The int8 weights are staged 256 k-bytes at a time (one 256-byte-wide TMA box per stage): the b200
// layout is row-major, so a 64-byte-wide box would touch 128 DRAM pages for 8 KB -- 256 contiguous bytes per row is the
// widest box TMA allows for 1-byte elements.  Each weight stage therefore feeds 4 consecutive 64-k MMA blocks; the
// activation tiles keep their own (64-k) stage ring.
constexpr int W_SUB          = 4;                       // 64-k sub-blocks per weight stage
constexpr int W_STAGE        = W_SUB * W8_TILE;         // 32 KB
constexpr int W_HALF         = BLOCK_N * 128;           // one 128-byte-wide swizzled box = 16 KB
__host__ __device__ constexpr int w_stages_for(int bt) { return bt >= 256 ? 2 : (bt >= 64 ? 3 : 4); }
__host__ __device__ constexpr int x_stages_for(int bt) { return bt >= 256 ? 3 : (bt >= 128 ? 3 : (bt >= 64 ? 4 : 6)); }
__host__ __device__ constexpr int x_tile_bytes(int bt) { return bt * BLOCK_K * 2; }
__host__ __device__ constexpr int smem_bytes_for(int bt)
{
    return 1024 /*alignment slack*/ + w_stages_for(bt) * W_STAGE + x_stages_for(bt) * x_tile_bytes(bt) + NUM_A_STAGES * A_TILE
           + 512 /*barriers*/;
}
__host__ __device__ constexpr int tmem_cols_for(int bt) { return bt < 32 ? 32 : bt; }

// ------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t holder_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (fp16 or bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1) |
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);
    d |= uint64_t(1) << 16;
    d |= uint64_t(1024 >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(2) << 61;
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @[4,6); a/b format
// (0 = f16, 1 = bf16) @[7,10)/[10,13); a/b K-major = 0 @15/@16; N >> 3 @[17,23); M >> 4 @[24,29)
__host__ __device__ constexpr uint32_t make_idesc(int is_bf16, int umma_m, int umma_n)
{
    return (1u << 4) | (uint32_t(is_bf16) << 7) | (uint32_t(is_bf16) << 10) | (uint32_t(umma_n >> 3) << 17)
           | (uint32_t(umma_m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------- dequant
// 16 int8 (one uint4) -> 16 fp16/bf16 (two uint4).
//   fp16: PRMT each (biased) byte under the exponent of 1024 -> 1024+u, subtract 1152 -> q exactly,
//         multiply by the channel scale in fp16 (one rounding) == reference arithmetic.
//   bf16: PRMT into the mantissa of 2^23 (fp32), subtract, pack to bf16 (exact: |q| <= 128).
template <typename T>
__device__ __forceinline__ void dequant16(const uint4& in, uint32_t scale2, uint4& out0, uint4& out1)
{
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};  // biased bytes u = q + 128 (b200 layout)
    uint32_t o[8];
    if constexpr (DTypeOf<T>::value == EETQ_B200_F16) {
        const __half2 bias = __half2half2(__ushort_as_half(0x6480));  // 1152
        const __half2 s2   = *reinterpret_cast<const __half2*>(&scale2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t lo = __byte_perm(w[i], 0x64646464u, 0x5140);
            uint32_t hi = __byte_perm(w[i], 0x64646464u, 0x5342);
            __half2 qlo = __hsub2(*reinterpret_cast<__half2*>(&lo), bias);
            __half2 qhi = __hsub2(*reinterpret_cast<__half2*>(&hi), bias);
            qlo         = __hmul2(qlo, s2);
            qhi         = __hmul2(qhi, s2);
            o[2 * i]     = *reinterpret_cast<uint32_t*>(&qlo);
            o[2 * i + 1] = *reinterpret_cast<uint32_t*>(&qhi);
        }
    }
    else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float f0 = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7650)) - 8388736.f;
            const float f1 = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7651)) - 8388736.f;
            const float f2 = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7652)) - 8388736.f;
            const float f3 = __uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7653)) - 8388736.f;
            __nv_bfloat162 lo = __floats2bfloat162_rn(f0, f1);
            __nv_bfloat162 hi = __floats2bfloat162_rn(f2, f3);
            o[2 * i]     = *reinterpret_cast<uint32_t*>(&lo);
            o[2 * i + 1] = *reinterpret_cast<uint32_t*>(&hi);
        }
    }
    out0 = make_uint4(o[0], o[1], o[2], o[3]);
    out1 = make_uint4(o[4], o[5], o[6], o[7]);
}

struct TcParams {
    const void* scales;
    const void* bias;
    void* y;
    int64_t ldy;
    int M, N, K;
    int splits;
    int* tile_counters;   // [n_tiles * t_tiles], zero on entry, zero on exit
    float* partials;      // [splits][n_tiles * t_tiles][BT][128]
};

// ------------------------------------------------------------------------------------------------- kernel
template <typename T, int BT>
__global__ void __launch_bounds__(TC_THREADS, 1)
    w8a16_gemm_tc_v1_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const TcParams p)
{
    constexpr int WS         = w_stages_for(BT);
    constexpr int XS         = x_stages_for(BT);
    constexpr int X_TILE     = x_tile_bytes(BT);
    constexpr int TMEM_COLS  = tmem_cols_for(BT);
    constexpr bool SCALE_IN_A = DTypeOf<T>::value == EETQ_B200_F16;
    constexpr uint32_t IDESC = make_idesc(DTypeOf<T>::value == EETQ_B200_BF16, BLOCK_N, BT);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w8_base   = smem_base;                                   // WS x 32 KB (two swizzled 16 KB halves each)
    const uint32_t x_base    = w8_base + WS * W_STAGE;                       // XS x X_TILE
    const uint32_t a_base    = x_base + XS * X_TILE;                         // 4 x 16 KB
    const uint32_t bar_base  = a_base + NUM_A_STAGES * A_TILE;
    const uint32_t wfull_bar  = bar_base;                 // WS x 8
    const uint32_t wempty_bar = wfull_bar + WS * 8;       // WS x 8
    const uint32_t xfull_bar  = wempty_bar + WS * 8;      // XS x 8
    const uint32_t xempty_bar = xfull_bar + XS * 8;       // XS x 8
    const uint32_t afull_bar  = xempty_bar + XS * 8;      // 4 x 8
    const uint32_t aempty_bar = afull_bar + NUM_A_STAGES * 8;
    const uint32_t tfull_bar = aempty_bar + NUM_A_STAGES * 8;
    const uint32_t tmem_holder = tfull_bar + 8;
    const uint32_t flag_holder = tmem_holder + 4;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));         // generic pointer to smem_base

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int n_tile = blockIdx.x;
    const int t_tile = blockIdx.y;
    const int split  = blockIdx.z;
    const int kb_total = p.K / BLOCK_K;
    const int kb_begin = int((int64_t(split) * kb_total) / p.splits);
    const int kb_end   = int((int64_t(split + 1) * kb_total) / p.splits);
    const int num_kb   = kb_end - kb_begin;

    // ------------------------------------------------------------------ one-time setup
    if (warp == W_PRODUCER_WARP && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < WS; ++s) {
            mbar_init(wfull_bar + 8 * s, 1);
            mbar_init(wempty_bar + 8 * s, DQ_WARPS);  // every dequant warp reads two sub-blocks of each weight stage
        }
        for (int s = 0; s < XS; ++s) {
            mbar_init(xfull_bar + 8 * s, 1);
            mbar_init(xempty_bar + 8 * s, 1);
        }
        for (int a = 0; a < NUM_A_STAGES; ++a) {
            mbar_init(afull_bar + 8 * a, DQ_WARPS / DQ_GROUPS);  // one elected arrive per warp of the owning group
            mbar_init(aempty_bar + 8 * a, 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP)
        tmem_alloc(tmem_holder, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_holder - smem_base));

    pdl_launch_dependents();
    pdl_wait_prior_grids();  // x (and y / workspace) may be produced by the previous kernel in the stream

    if (warp == W_PRODUCER_WARP) {
        // ============================================================== TMA producer
        if (lane == 0) {
            // weight stream: free-running (bounded only by its own ring), one 256-k stage per 4 MMA k-blocks
            for (int wi = 0; wi * W_SUB < num_kb; ++wi) {
                const int k0       = (kb_begin + wi * W_SUB) * BLOCK_K;
                const int ws       = wi % WS;
                const uint32_t wph = (wi / WS) & 1;
                mbar_wait(wempty_bar + 8 * ws, wph ^ 1);
                mbar_arrive_expect_tx(wfull_bar + 8 * ws, W_STAGE);
                // k beyond K is zero-filled by TMA (and never converted: sub-blocks past num_kb are skipped)
                tma_load_2d(w8_base + ws * W_STAGE, &map_w, wfull_bar + 8 * ws, k0, n_tile * BLOCK_N);
            }
        }
    }
    else if (warp == X_PRODUCER_WARP) {
        // ============================================================== activation TMA producer (own warp, so a full
        // activation ring never stalls the weight prefetch)
        if (lane == 0) {
            for (int it = 0; it < num_kb; ++it) {
                const int k0       = (kb_begin + it) * BLOCK_K;
                const int xs       = it % XS;
                const uint32_t xph = (it / XS) & 1;
                mbar_wait(xempty_bar + 8 * xs, xph ^ 1);
                mbar_arrive_expect_tx(xfull_bar + 8 * xs, X_TILE);
                tma_load_2d(x_base + xs * X_TILE, &map_x, xfull_bar + 8 * xs, k0, t_tile * BT);
            }
        }
    }
    else if (warp == MMA_WARP) {
        // ============================================================== MMA issuer (one thread)
        if (lane == 0) {
            for (int it = 0; it < num_kb; ++it) {
                const int s       = it % XS;
                const uint32_t ph = (it / XS) & 1;
                const int a       = it % NUM_A_STAGES;
                const uint32_t aph = (it / NUM_A_STAGES) & 1;
                mbar_wait(xfull_bar + 8 * s, ph);     // activation tile landed
                mbar_wait(afull_bar + 8 * a, aph);    // dequantised weight tile written
                tc_fence_after();
                const uint64_t a_desc = make_kmajor_sw128_desc(a_base + a * A_TILE);
                const uint64_t b_desc = make_kmajor_sw128_desc(x_base + s * X_TILE);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (>>4) address field
                    umma_f16(tmem_base, a_desc + uint64_t(2 * k), b_desc + uint64_t(2 * k), IDESC, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(xempty_bar + 8 * s);    // frees the activation stage
                umma_commit(aempty_bar + 8 * a);    // frees the A stage
            }
            umma_commit(tfull_bar);                  // accumulator complete
        }
    }
    else {
        // ============================================================== dequantisers, then epilogue
        const int dt = threadIdx.x;       // 0..255 (warps 0..7)
        // Two groups of 4 warps convert alternating k-blocks, so one group's smem round trips and barrier hops overlap
        // the other's; inside a group every thread batches its 4 LDS.128 before converting (ILP) and each warp
        // signals the MMA issuer with ONE elected mbarrier arrive.
        constexpr int GROUP_THREADS     = DQ_THREADS / DQ_GROUPS;             // 128
        constexpr int CHUNKS_PER_THREAD = (W8_TILE / 16) / GROUP_THREADS;     // 4
        const int grp = dt / GROUP_THREADS;
        const int gt  = dt % GROUP_THREADS;
        // this thread always converts the same rows: fetch their channel scales once
        uint32_t scale2[CHUNKS_PER_THREAD];
#pragma unroll
        for (int j = 0; j < CHUNKS_PER_THREAD; ++j) {
            scale2[j] = 0;
            if constexpr (SCALE_IN_A) {
                const int n      = n_tile * BLOCK_N + ((gt + GROUP_THREADS * j) >> 2);
                const __half sv  = (n < p.N) ? static_cast<const __half*>(p.scales)[n] : __ushort_as_half(0);
                const __half2 s2 = __half2half2(sv);
                scale2[j]        = *reinterpret_cast<const uint32_t*>(&s2);
            }
        }
        for (int it = grp; it < num_kb; it += DQ_GROUPS) {
            const int wi       = it / W_SUB;
            const int sub_k    = it % W_SUB;          // which 64-k slice of the 256-k weight stage
            const int ws       = wi % WS;
            const uint32_t wph = (wi / WS) & 1;
            const int a        = it % NUM_A_STAGES;
            const uint32_t aph = (it / NUM_A_STAGES) & 1;
            mbar_wait(wfull_bar + 8 * ws, wph);       // int8 stage landed
            // stage = [128 rows][256 B] as ONE un-swizzled TMA box: 256-byte rows are the widest TMA allows for 1-byte
            // elements and halve the number of DRAM requests per stage (the TMA unit, not HBM, bounds a weight stream
            // made of 128-byte requests); the 2-way LDS bank conflict this costs is negligible next to that
            const uint8_t* w8 = smem_gen + (w8_base - smem_base) + ws * W_STAGE;
            uint8_t* at       = smem_gen + (a_base - smem_base) + a * A_TILE;
            uint4 in[CHUNKS_PER_THREAD];
#pragma unroll
            for (int j = 0; j < CHUNKS_PER_THREAD; ++j) {
                const int c   = gt + GROUP_THREADS * j;
                const int row = c >> 2;
                in[j] = *reinterpret_cast<const uint4*>(w8 + row * 256 + sub_k * 64 + (c & 3) * 16);
            }
            mbar_wait(aempty_bar + 8 * a, aph ^ 1);   // A stage free (the MMA that read it has completed)
#pragma unroll
            for (int j = 0; j < CHUNKS_PER_THREAD; ++j) {
                const int c   = gt + GROUP_THREADS * j;  // 16-byte chunk index in the [128][64 B] slice
                const int row = c >> 2;
                const int kc  = c & 3;
                uint4 o0, o1;
                dequant16<T>(in[j], scale2[j], o0, o1);
                // 128B swizzle: 16-byte chunk index XOR (row & 7)
                uint8_t* rowp = at + row * 128;
                *reinterpret_cast<uint4*>(rowp + (((2 * kc) ^ (row & 7)) << 4))     = o0;
                *reinterpret_cast<uint4*>(rowp + (((2 * kc + 1) ^ (row & 7)) << 4)) = o1;
            }
            fence_proxy_async_smem();            // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(afull_bar + 8 * a);
                // the int8 bytes have been consumed (converted): after this group's last sub-block of the weight stage
                // (sub-blocks 2 / 3, or its final k-block) hand the stage back to the TMA producer -- once per warp
                if (sub_k >= W_SUB - DQ_GROUPS || it + DQ_GROUPS >= num_kb)
                    mbar_arrive(wempty_bar + 8 * ws);
            }
        }

        // ---------------------------------------------------------- epilogue
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        const int quad   = warp & 3;                 // TMEM lane quadrant this warp may read
        const int half_i = warp >> 2;                // two warps share a quadrant: split the columns
        const int n      = n_tile * BLOCK_N + quad * 32 + lane;
        const bool n_ok  = n < p.N;
        constexpr int COLS_PER_WARP = BT / 2;        // BT >= 32 -> multiple of 16; BT == 16 handled below
        constexpr int CHUNK = 16;
        const int col_begin = (BT >= 32) ? half_i * COLS_PER_WARP : 0;
        const int col_end   = (BT >= 32) ? col_begin + COLS_PER_WARP : ((half_i == 0) ? BT : 0);

        float scale_f = 1.f, bias_f = 0.f;
        if (n_ok) {
            if constexpr (!SCALE_IN_A)
                scale_f = to_float(static_cast<const T*>(p.scales)[n]);
            if (p.bias != nullptr)
                bias_f = to_float(static_cast<const T*>(p.bias)[n]);
        }
        T* y = static_cast<T*>(p.y);
        const int tile_id = t_tile * gridDim.x + n_tile;

        if (p.splits == 1) {
            for (int c0 = col_begin; c0 < col_end; c0 += CHUNK) {
                uint32_t r[16];
                tmem_ld_x16(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    const int t = t_tile * BT + c0 + j;
                    if (n_ok && t < p.M)
                        y[int64_t(t) * p.ldy + n] = from_float<T>(__uint_as_float(r[j]) * scale_f + bias_f);
                }
            }
        }
        else {
            // split-K: publish this CTA's partial tile, last arriver reduces in split order
            float* my_part = p.partials + (int64_t(split) * (gridDim.x * gridDim.y) + tile_id) * (BT * BLOCK_N);
            for (int c0 = col_begin; c0 < col_end; c0 += CHUNK) {
                uint32_t r[16];
                tmem_ld_x16(tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < CHUNK; ++j)
                    my_part[(c0 + j) * BLOCK_N + quad * 32 + lane] = __uint_as_float(r[j]);
            }
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(DQ_THREADS) : "memory");
            // All `splits` CTAs of this tile are co-resident (grid <= 148 CTAs, one per SM): meet at an arrive counter, then
            // EVERY CTA reduces its own 1/splits share of the token columns (in split order -> deterministic) instead of
            // leaving one CTA to walk the whole tile through a chain of dependent L2 round trips.
            int* arrive = p.tile_counters + tile_id;
            int* depart = p.tile_counters + 512 + tile_id;
            if (dt == 0) {
                atomicAdd(arrive, 1);
                while (*reinterpret_cast<volatile int*>(arrive) < p.splits) {
                }
                __threadfence();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(DQ_THREADS) : "memory");
            const int cs      = BT / p.splits;                 // columns reduced by this CTA (splits is a power of two <= 8)
            const int my_c0   = split * cs;
            const int per_grp = (cs >= 2) ? cs / 2 : cs;       // the two warps of a lane quadrant share the columns
            const int c_begin = my_c0 + ((cs >= 2) ? half_i * per_grp : 0);
            const int c_end   = (cs >= 2 || half_i == 0) ? c_begin + per_grp : c_begin;
            const int64_t tile_stride = int64_t(gridDim.x) * gridDim.y * (BT * BLOCK_N);
            const float* base = p.partials + int64_t(tile_id) * (BT * BLOCK_N) + quad * 32 + lane;
            for (int c = c_begin; c < c_end; c += 4) {
                float v[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int sp = 0; sp < 8; ++sp)
                        v[u][sp] = (sp < p.splits && c + u < c_end) ? __ldcg(base + sp * tile_stride + (c + u) * BLOCK_N) : 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float acc = 0.f;
#pragma unroll
                    for (int sp = 0; sp < 8; ++sp)
                        acc += v[u][sp];
                    const int t = t_tile * BT + c + u;
                    if (c + u < c_end && n_ok && t < p.M)
                        y[int64_t(t) * p.ldy + n] = from_float<T>(acc * scale_f + bias_f);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(DQ_THREADS) : "memory");
            if (dt == 0) {
                // last CTA to leave resets both counters so the workspace stays clean for the next call
                if (atomicAdd(depart, 1) == p.splits - 1) {
                    *reinterpret_cast<volatile int*>(arrive) = 0;
                    *reinterpret_cast<volatile int*>(depart) = 0;
                }
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------- host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess
            && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

struct MapKey {
    const void* ptr;
    uint64_t d0, d1, stride;
    uint32_t b0, b1;
    int dtype, swizzle;
    bool operator==(const MapKey& o) const
    {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && dtype == o.dtype
               && swizzle == o.swizzle;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        size_t h = std::hash<const void*>()(k.ptr);
        auto mix = [&h](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.stride); mix(k.b0); mix(k.b1); mix(uint64_t(k.dtype)); mix(uint64_t(k.swizzle));
        return h;
    }
};

// 2-D tensor map over a row-major [d1][d0] matrix (d0 contiguous) with row pitch `stride` bytes; cached.
int get_tensor_map(const void* ptr, CUtensorMapDataType dt, int dtype_tag, uint64_t d0, uint64_t d1, uint64_t stride, uint32_t b0,
                   uint32_t b1, CUtensorMapSwizzle swz, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    MapKey key{ptr, d0, d1, stride, b0, b1, dtype_tag, int(swz)};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return EETQ_B200_OK;
        }
    }
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
        return EETQ_B200_ECUDA;
    }
    cuuint64_t dims[2]    = {d0, d1};
    cuuint64_t strides[1] = {stride};
    cuuint32_t box[2]     = {b0, b1};
    cuuint32_t estr[2]    = {1, 1};
    CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu, box %u x %u)", int(r),
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)stride, b0, b1);
        return EETQ_B200_ECUDA;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096)
        cache.clear();
    cache.emplace(key, *out);
    return EETQ_B200_OK;
}

struct TcConfig {
    int bt;        // tokens per CTA tile (UMMA N)
    int n_tiles;   // ceil(N / 128)
    int t_tiles;   // ceil(M / bt)
    int splits;    // split-K factor
};

TcConfig choose_config(int64_t M, int64_t N, int64_t K)
{
    // One CTA per SM (smem), 148 slots.  Dequantisation work scales with the number of token tiles, so up to 256
    // tokens ride in ONE tile and spare SMs are filled by splitting K; above that, 256-token tiles.
    TcConfig c{};
    c.bt      = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
    c.n_tiles = int((N + BLOCK_N - 1) / BLOCK_N);
    c.t_tiles = int((M + c.bt - 1) / c.bt);
    const int tiles = c.n_tiles * c.t_tiles;
    const int kb    = int(K / BLOCK_K);
    int s           = (device_info().ok ? device_info().sm_count : 148) / tiles;
    const int max_s = kb / 8 > 0 ? kb / 8 : 1;  // at least 8 k-blocks (512 k) per split
    if (s > max_s) s = max_s;
    if (s > 8) s = 8;
    if (s < 1) s = 1;
    while (s & (s - 1)) --s;  // power of two: every split CTA reduces an equal share of the token columns
    c.splits = s;
    return c;
}

template <typename T, int BT>
int launch_tc(const CUtensorMap& map_w, const CUtensorMap& map_x, const TcParams& p, const TcConfig& cfg, bool pdl, cudaStream_t stream)
{
    auto kernel = w8a16_gemm_tc_v1_kernel<T, BT>;
    constexpr int smem = smem_bytes_for(BT);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        EB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim          = dim3(unsigned(cfg.n_tiles), unsigned(cfg.t_tiles), unsigned(cfg.splits));
    lc.blockDim         = dim3(TC_THREADS);
    lc.dynamicSmemBytes = smem;
    lc.stream           = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs                                           = attr;
    lc.numAttrs                                        = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&lc, kernel, map_w, map_x, p);
    count_launch();
    if (e != cudaSuccess) {
        set_error("gemm_tc launch failed: %s", cudaGetErrorString(e));
        return EETQ_B200_ECUDA;
    }
    return EETQ_B200_OK;
}

template <typename T>
int launch_tc_bt(const CUtensorMap& map_w, const CUtensorMap& map_x, const TcParams& p, const TcConfig& cfg, bool pdl, cudaStream_t stream)
{
    switch (cfg.bt) {
        case 16: return launch_tc<T, 16>(map_w, map_x, p, cfg, pdl, stream);
        case 32: return launch_tc<T, 32>(map_w, map_x, p, cfg, pdl, stream);
        case 64: return launch_tc<T, 64>(map_w, map_x, p, cfg, pdl, stream);
        case 128: return launch_tc<T, 128>(map_w, map_x, p, cfg, pdl, stream);
        default: return launch_tc<T, 256>(map_w, map_x, p, cfg, pdl, stream);
    }
}

// The tile counters live in a FIXED-size region at the start of the workspace so that calls with different
// (M, N, K) -- hence different partial-buffer layouts -- can share one zero-initialised workspace: only the counter
// region must stay zero between calls, and every kernel leaves it zero.  Split-K is only chosen when
// tiles <= 296, so 4 KiB (1024 counters) is always enough.
constexpr size_t kCounterRegionBytes = 4096;
size_t counters_bytes(const TcConfig&) { return kCounterRegionBytes; }

}  // namespace

size_t gemm_tc_v1_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    if (M <= 0 || N <= 0 || K < BLOCK_K)
        return 0;
    const TcConfig c = choose_config(M, N, K);
    if (c.splits == 1)
        return 0;
    return counters_bytes(c) + size_t(c.splits) * c.n_tiles * c.t_tiles * c.bt * BLOCK_N * sizeof(float);
}

int launch_gemm_tc_v1(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, void* y, int64_t ldy,
                   int64_t M, int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes, bool pdl,
                   cudaStream_t stream)
{
    TcConfig cfg = choose_config(M, N, K);
    const size_t need = gemm_tc_v1_workspace_bytes(M, N, K);
    if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
        // no (or too small a) workspace: fall back to a single split rather than failing
        cfg.splits = 1;
    }
    CUtensorMap map_w, map_x;
    if (int rc = get_tensor_map(w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 100, uint64_t(K), uint64_t(N), uint64_t(K), 256, BLOCK_N,
                                CU_TENSOR_MAP_SWIZZLE_NONE, &map_w))
        return rc;
    const CUtensorMapDataType xdt = dtype == EETQ_B200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (int rc = get_tensor_map(x, xdt, dtype, uint64_t(K), uint64_t(M), uint64_t(ldx) * 2, BLOCK_K, uint32_t(cfg.bt),
                                CU_TENSOR_MAP_SWIZZLE_128B, &map_x))
        return rc;

    TcParams p{};
    p.scales = scales;
    p.bias   = bias;
    p.y      = y;
    p.ldy    = ldy;
    p.M      = int(M);
    p.N      = int(N);
    p.K      = int(K);
    p.splits = cfg.splits;
    if (cfg.splits > 1) {
        p.tile_counters = static_cast<int*>(workspace);
        p.partials      = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + counters_bytes(cfg));
    }
    if (dtype == EETQ_B200_F16)
        return launch_tc_bt<__half>(map_w, map_x, p, cfg, pdl, stream);
    return launch_tc_bt<__nv_bfloat16>(map_w, map_x, p, cfg, pdl, stream);
}

}  // namespace eetq_b200
