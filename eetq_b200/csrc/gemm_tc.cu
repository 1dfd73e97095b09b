// gemm_tc.cu -- batched / prefill w8a16 GEMM on 5th-generation tensor cores (tcgen05 + TMEM + TMA, sm_100a).
//
// Replaces the reference prefill path
//   CutlassFpAIntBGemmRunner<half,uint8_t>::gemm   /root/reference/csrc/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:441-553
//   GemmFpAIntB + DqMmaMultistage                  /root/reference/csrc/cutlass_extensions/include/cutlass_extensions/gemm/kernel/fpA_intB_gemm.h:60-486,
//                                                  .../gemm/threadblock/dq_mma_multistage.h:98-591
// (mma.sync + cp.async + ldmatrix, compiled out for sm >= 90) with a Blackwell-native kernel; nothing of the
// CUTLASS 2.x structure is kept.
//
// Formulation (operands swapped so that small token counts stay efficient and the per-channel scale is a property of
// an accumulator ROW):      D[n, t] = sum_k  A[n, k] * B[t, k]
//      A = dequantised weights [128 features x 256 k] fp16/bf16 **in tensor memory** (tcgen05.st by the dequant warps;
//          the MMA reads its A operand straight from TMEM -- no shared-memory round trip, no async-proxy fence)
//      B = activation tile     [BT tokens x 64 k]  K-major, 128B-swizzled shared memory   (TMA straight from x)
//      D = fp32 accumulator in TMEM: lane = feature, column = token   (UMMA M = 128, N = BT, K = 16)
//
// Persistent stream-K schedule.  The work is cut into UNITS of (128-feature tile, 256-k stage).  The grid is one CTA per SM
// (never more CTAs than can be resident); CTA g walks the contiguous unit range [g*U/G, (g+1)*U/G), i.e. the tail of one
// tile, whole tiles, and the head of another.  A tile that is shared by several CTAs is finished by its OWNER (the CTA that
// holds the tile's first stage -- for which the tile is the LAST segment of its range); the other contributors meet
// the tile as the FIRST segment of their range, dump their fp32 partial to a workspace slot and raise a flag, long before
// the owner gets there.  The owner adds the partials in k order (deterministic) and writes y.  A CTA therefore only ever
// waits for first-segment partials of higher-indexed CTAs, which those produce without waiting for anybody: progress
// needs no grid-wide co-residency (this replaces round 1's spin-wait split-K; the reference disables split-K altogether by
// passing a null workspace, fpA_intB_gemm_wrapper.cu:169-170).
//
// Warp roles ((7 + DQW) warps):  warps 0..3   epilogue (TMEM -> registers -> y / workspace), one per TMEM lane quadrant
//                                warps 4..4+DQW-1  dequantisers: int8 smem -> fp16/bf16 registers -> TMEM (tcgen05.st)
//                                warp 4+DQW   weight TMA producer | 5+DQW  tcgen05.mma issuer + TMEM alloc | 6+DQW  activation TMA
//                                (single-thread roles on the HIGHEST warp ids: the issue arbiter favours them)
// Pipelines (mbarrier):  wfull/wempty[WS]  TMA <-> dequant   (int8, 256 k per stage = two 128-byte-wide swizzled boxes)
//                        xfull/xempty[XS]  TMA <-> MMA       (activations)
//                        afull/aempty[2]   dequant <-> MMA   (A operand in TMEM, 256 k = 16 UMMAs per hand-off)
//                        tfull/tempty[ND]  MMA <-> epilogue  (accumulator; double-buffered when BT <= 128, so the epilogue
//                                                             of one segment overlaps the main loop of the next)
//
// Arithmetic.  fp16: A = fp16(fp16(q) * s) with ONE rounding per weight and fp32 accumulation -- exactly the
// reference's K1 arithmetic (mma_tensorop_dequantizer.h:259-274, default_fpA_intB_traits.h:110), so results
// match it up to fp32 summation order.  bf16 (extension): A = bf16(q) exactly, scale applied in the fp32 epilogue.
//
// Roofline: bytes = K*N + 2N + 2MK + 2MN, flops = 2MNK; HBM-bound for M <~ 140, tensor-bound above.
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace eetq_b200 {

namespace {

constexpr int BLOCK_N      = 128;  // output features per tile  (UMMA M)
constexpr int STAGE_K      = 256;  // k per weight stage / per A operand stage (one dequant -> MMA hand-off)
constexpr int SUB_K        = 64;   // k per activation TMA box (64 fp16 = one 128-byte swizzle row)
constexpr int SUBS         = STAGE_K / SUB_K;
constexpr int UMMA_K       = 16;
constexpr int W_BOX_BYTES  = BLOCK_N * 128;       // one 128-byte-wide swizzled int8 box = 16 KB
constexpr int W_STAGE      = 2 * W_BOX_BYTES;     // 32 KB
constexpr int A_STAGES     = 2;
constexpr int A_STAGE_COLS = STAGE_K / 2;         // 128 TMEM columns (two 16-bit values per 32-bit column)
constexpr int A_COL0       = 256;                 // TMEM columns [256, 512) hold the A ring, [0, 256) the accumulators
constexpr int TMEM_COLS    = 512;
constexpr int EPI_WARPS    = 4;
constexpr int EPI_THREADS  = EPI_WARPS * 32;
constexpr int kFlagRegionBytes = 4096;            // one int flag per CTA (<= 1024 CTAs)
constexpr int TRACE_SLOTS  = 64;

__host__ __device__ constexpr int tc_threads(int dqw) { return (EPI_WARPS + dqw + 3) * 32; }
__host__ __device__ constexpr int w_stages_for(int bt) { return bt <= 64 ? 5 : (bt <= 128 ? 4 : 3); }
// activation ring: for small token tiles one ring stage carries all four 64-k boxes of a weight stage (one wait per
// hand-off); for large tiles a stage is a single 64-k box
__host__ __device__ constexpr int x_sub_for(int bt) { return bt <= 32 ? 4 : 1; }
__host__ __device__ constexpr int x_stages_for(int bt) { return bt <= 16 ? 4 : (bt <= 32 ? 3 : (bt <= 64 ? 6 : (bt <= 128 ? 4 : 3))); }
__host__ __device__ constexpr int x_stage_bytes(int bt) { return x_sub_for(bt) * bt * SUB_K * 2; }
__host__ __device__ constexpr int smem_bytes_for(int bt)
{
    return 1024 /*alignment slack*/ + w_stages_for(bt) * W_STAGE + x_stages_for(bt) * x_stage_bytes(bt) + 512 /*barriers*/;
}
__host__ __device__ constexpr int acc_buffers_for(int bt) { return bt <= 128 ? 2 : 1; }

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
// Every wait in this kernel is followed by warp-collective `.sync.aligned` instructions (tcgen05.st / tcgen05.ld / tcgen05.wait)
// or by an elected-lane issue.  Lanes leave the try_wait spin loop on different iterations whenever the wait really blocks, and
// nothing re-converges them by itself: a tcgen05.st.sync.aligned executed by a partial warp silently drops lanes (seen as single
// wrong feature rows, only in the MMA-bound regime where the dequant warps block on `aempty`).  So: wait, then re-converge.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity)
{
    mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::f16 (fp16 or bf16 inputs, fp32 accumulate), A operand read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread (the dequantised A operand: 64 fp16/bf16 per lane)
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one elected lane of a fully converged warp (the whole warp runs the surrounding loop, so every value that feeds the
// tcgen05 / TMA instructions stays provably warp-uniform and lives in uniform registers; round 1 ran these loops under
// `if (lane == 0)`, which made the compiler wrap EVERY tcgen05.mma / commit / TMA in a per-instruction "which lanes hold which
// value" loop -- ~120 cycles per issue, the 0.39 us per 64-k block floor of the round-1 kernel)
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }
template <int THREADS>
__device__ __forceinline__ void joint_bar_sync() { asm volatile("bar.sync 2, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
// 16 int8 (one uint4 of biased bytes u = q + 128, b200 layout) -> 16 fp16/bf16 in 8 registers, k order preserved.
//   fp16: PRMT each byte under the exponent of 1024 -> 1024+u, subtract 1152 -> q exactly,
//         multiply by the channel scale in fp16 (one rounding) == reference arithmetic.
//   bf16: PRMT into the mantissa of 2^23 (fp32), subtract, pack to bf16 (exact: |q| <= 128).
template <typename T>
__device__ __forceinline__ void dequant16(const uint4& in, uint32_t scale2, uint32_t* o)
{
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
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
            // q is an integer with at most 8 significant bits: the upper half of its fp32 pattern IS its bf16 pattern
            o[2 * i]     = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
            o[2 * i + 1] = __byte_perm(__float_as_uint(f2), __float_as_uint(f3), 0x7632);
        }
    }
}

struct TcParams {
    const void* scales;
    const void* bias;
    const void* residual;  // optional [M, N] added to the output in the epilogue (row stride ldr)
    int64_t ldr;
    void* y;
    int64_t ldy;
    int M, N, K;
    int n_tiles, t_tiles;
    int spt;           // 256-k stages per tile
    int units_q, units_r;  // CTA g owns units [g q + min(g, r), ...): the first r CTAs take q + 1 (tile-aligned mode: tiles)
    int tile_aligned;  // 1: every CTA owns whole tiles (no workspace needed)
    int* flags;        // [grid] zero on entry, zero on exit
    float* slots;      // [grid][BT][128] fp32 partial tiles
    unsigned long long* trace;  // optional [grid][TRACE_SLOTS] clock samples (instrumented build only)
};

__device__ __forceinline__ int unit_begin(const TcParams& p, int g)
{
    const int b = g * p.units_q + min(g, p.units_r);
    return p.tile_aligned ? b * p.spt : b;
}
// partial-tile slot layout: [BT / 4][128 rows][4 columns] fp32 -- a thread's 4 consecutive columns are one 16-byte access and
// a warp's accesses are contiguous
__device__ __forceinline__ int slot_index(int c, int row) { return ((c >> 2) * BLOCK_N + row) * 4; }

__device__ __forceinline__ unsigned long long clk64()
{
    unsigned long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
    return c;
}
__device__ __forceinline__ unsigned long long gtime_ns()
{
    unsigned long long c;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(c));
    return c;
}

// ------------------------------------------------------------------------------------------------- kernel
struct Segment {
    int tile, s0, s1, n_tile, t_tile;
};
__device__ __forceinline__ Segment segment_at(const TcParams& p, int u, int u1)
{
    Segment sg;
    sg.tile   = u / p.spt;
    sg.s0     = u - sg.tile * p.spt;
    sg.s1     = min(p.spt, sg.s0 + (u1 - u));
    sg.t_tile = sg.tile / p.n_tiles;
    sg.n_tile = sg.tile - sg.t_tile * p.n_tiles;
    return sg;
}

// Epilogue of one segment for the columns {16 (part + i nparts)} of accumulator buffer d, by the calling warp (TMEM lane
// quadrant `quad`).  kind 0: whole tile -> y;  1: contributor -> workspace slot;  2: owner -> own accumulator + the slots of
// CTAs g+1 .. g_last, in that (= ascending k) order -> y.
template <typename T, int BT, bool SCALE_IN_A, int FB>
__device__ __forceinline__ void epilogue_columns(const TcParams& p, const Segment& sg, int kind, uint32_t tmem_d, int quad, int lane,
                                                 int part, int nparts, int g, int g_last, uint32_t tempty_bar_or_0)
{
    const int row   = quad * 32 + lane;
    const int n     = sg.n_tile * BLOCK_N + row;
    const bool n_ok = n < p.N;
    float scale_f = 1.f, bias_f = 0.f;
    if (kind != 1 && n_ok) {
        if constexpr (!SCALE_IN_A)
            scale_f = to_float(static_cast<const T*>(p.scales)[n]);
        if (p.bias != nullptr)
            bias_f = to_float(static_cast<const T*>(p.bias)[n]);
    }
    T* y           = static_cast<T*>(p.y);
    float* my_slot = p.slots + size_t(g) * (BT * BLOCK_N);
    constexpr int NCHUNK = BT / 16;
#pragma unroll 1
    for (int ci = part; ci < NCHUNK; ci += nparts) {
        const int c0 = ci * 16;
        // owner: request the contributors' partials for this chunk first (independent loads, FB contributors per batch in
        // flight), then read the own accumulator while they travel
        float4 w[FB][4];
        int gg0 = g + 1;
        if (kind == 2) {
#pragma unroll
            for (int b = 0; b < FB; ++b)
                if (gg0 + b <= g_last) {
                    const float4* src = reinterpret_cast<const float4*>(p.slots + size_t(gg0 + b) * (BT * BLOCK_N) + slot_index(c0, row));
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        w[b][q4] = __ldcg(src + q4 * BLOCK_N);
                }
        }
        uint32_t r[16];
        tmem_ld_x16(tmem_d + (uint32_t(quad * 32) << 16) + uint32_t(c0), r);
        tmem_ld_wait();
        if (tempty_bar_or_0 != 0 && ci + nparts >= NCHUNK) {
            // all of this accumulator is in registers: let the MMA issuer reuse it for the next segment
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty_bar_or_0);
        }
        if (kind == 1) {
            float4* dst = reinterpret_cast<float4*>(my_slot + slot_index(c0, row));
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
                dst[q4 * BLOCK_N] = make_float4(__uint_as_float(r[4 * q4]), __uint_as_float(r[4 * q4 + 1]), __uint_as_float(r[4 * q4 + 2]),
                                                __uint_as_float(r[4 * q4 + 3]));
            continue;
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            v[j] = __uint_as_float(r[j]);
        if (kind == 2) {
            while (true) {
#pragma unroll
                for (int b = 0; b < FB; ++b)
                    if (gg0 + b <= g_last) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            v[4 * q4] += w[b][q4].x;
                            v[4 * q4 + 1] += w[b][q4].y;
                            v[4 * q4 + 2] += w[b][q4].z;
                            v[4 * q4 + 3] += w[b][q4].w;
                        }
                    }
                gg0 += FB;
                if (gg0 > g_last)
                    break;
#pragma unroll
                for (int b = 0; b < FB; ++b)
                    if (gg0 + b <= g_last) {
                        const float4* src = reinterpret_cast<const float4*>(p.slots + size_t(gg0 + b) * (BT * BLOCK_N) + slot_index(c0, row));
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            w[b][q4] = __ldcg(src + q4 * BLOCK_N);
                    }
            }
        }
        // The 4 epilogue warps are ONE warp per scheduler, i.e. latency-bound on every dependent instruction: keep the per-element
        // work to convert + store + pointer bump (the straightforward indexed form cost ~25 instructions per element and made the
        // epilogue of a 256-token tile as long as its whole main loop).
        const int t0   = sg.t_tile * BT + c0;
        const int tcnt = n_ok ? min(16, p.M - t0) : 0;  // valid token rows of this chunk (<= 0: nothing to store)
        T* yp          = y + int64_t(t0) * p.ldy + n;
        if (p.residual == nullptr) {
            if (tcnt == 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    *yp = from_float<T>(v[j] * scale_f + bias_f);
                    yp += p.ldy;
                }
            }
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j < tcnt)
                        *yp = from_float<T>(v[j] * scale_f + bias_f);
                    yp += p.ldy;
                }
            }
        }
        else {
            const T* rp = static_cast<const T*>(p.residual) + int64_t(t0) * p.ldr + n;
            T rv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {  // all residual loads first (independent), then the adds and stores
                rv[j] = (j < tcnt) ? *rp : from_float<T>(0.f);
                rp += p.ldr;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j < tcnt)
                    *yp = from_float<T>(to_float(from_float<T>(v[j] * scale_f + bias_f)) + to_float(rv[j]));
                yp += p.ldy;
            }
        }
    }
}

template <typename T, int BT, int DQW, bool TRACE>
__global__ void __launch_bounds__(tc_threads(DQW), 1)
    w8a16_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const TcParams p)
{
    constexpr int WS          = w_stages_for(BT);
    constexpr int XS          = x_stages_for(BT);
    constexpr int XSUB        = x_sub_for(BT);      // 64-k boxes per activation ring stage
    constexpr int XSTEPS      = SUBS / XSUB;        // activation ring stages per weight stage
    constexpr int X_STAGE     = x_stage_bytes(BT);
    constexpr int X_BOX       = BT * SUB_K * 2;
    constexpr int ND          = acc_buffers_for(BT);
    constexpr bool SCALE_IN_A = DTypeOf<T>::value == EETQ_B200_F16;
    constexpr uint32_t IDESC  = make_idesc(DTypeOf<T>::value == EETQ_B200_BF16, BLOCK_N, BT);
    constexpr int NP          = DQW / 4;            // dequant warps per TMEM lane quadrant: each takes 1/NP of a stage's k
    constexpr int CH          = 16 / NP;            // 16-byte chunks per thread per stage (8 or 4)
    constexpr int W_PRODUCER_WARP = EPI_WARPS + DQW;
    constexpr int MMA_WARP        = EPI_WARPS + DQW + 1;
    constexpr int X_PRODUCER_WARP = EPI_WARPS + DQW + 2;
    constexpr int JOINT_THREADS   = (EPI_WARPS + DQW) * 32;  // epilogue + dequant warps: together they finish an owned tile
    static_assert(DQW == 8 || DQW == 16, "8 or 16 dequant warps");
    static_assert(ND * BT <= A_COL0, "accumulators must fit below the A ring");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base  = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w8_base    = smem_base;                      // WS x 32 KB (two 16 KB swizzled boxes each)
    const uint32_t x_base     = w8_base + WS * W_STAGE;          // XS x X_STAGE
    const uint32_t bar_base   = x_base + XS * X_STAGE;
    const uint32_t wfull_bar  = bar_base;
    const uint32_t wempty_bar = wfull_bar + WS * 8;
    const uint32_t xfull_bar  = wempty_bar + WS * 8;
    const uint32_t xempty_bar = xfull_bar + XS * 8;
    const uint32_t afull_bar  = xempty_bar + XS * 8;
    const uint32_t aempty_bar = afull_bar + A_STAGES * 8;
    const uint32_t tfull_bar  = aempty_bar + A_STAGES * 8;
    const uint32_t tempty_bar = tfull_bar + 2 * 8;
    const uint32_t tmem_holder = tempty_bar + 2 * 8;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to smem_base

    const int warp = int(warp_uniform(threadIdx.x >> 5));
    const int lane = threadIdx.x & 31;
    const int g    = blockIdx.x;
    const int G    = gridDim.x;
    const int u0   = unit_begin(p, g);
    const int u1   = unit_begin(p, g + 1);

    unsigned long long* tr = nullptr;
    if constexpr (TRACE) {
        tr = p.trace + size_t(g) * TRACE_SLOTS;
        if (threadIdx.x == 0) {
            tr[48] = gtime_ns();
            tr[49] = clk64();
        }
    }

    // ------------------------------------------------------------------ one-time setup
    if (warp == W_PRODUCER_WARP && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < WS; ++s) {
            mbar_init(wfull_bar + 8 * s, 1);
            mbar_init(wempty_bar + 8 * s, DQW);
        }
        for (int s = 0; s < XS; ++s) {
            mbar_init(xfull_bar + 8 * s, 1);
            mbar_init(xempty_bar + 8 * s, 1);
        }
        for (int a = 0; a < A_STAGES; ++a) {
            mbar_init(afull_bar + 8 * a, DQW);
            mbar_init(aempty_bar + 8 * a, 1);
        }
        for (int d = 0; d < 2; ++d) {
            mbar_init(tfull_bar + 8 * d, 1);
            mbar_init(tempty_bar + 8 * d, EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == MMA_WARP)
        tmem_alloc(tmem_holder, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = warp_uniform(*reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_holder - smem_base)));

    pdl_launch_dependents();
    if constexpr (TRACE) {
        if (threadIdx.x == 0)
            tr[50] = clk64();
    }

    // The last segment of a CTA's range may be an OWNED tile whose remaining k range was computed by the following CTAs:
    // its epilogue (own accumulator + their partials) is done by the epilogue AND dequant warps together at the very end.
    // (Closed form -- the segment walk this replaces sat in front of the role dispatch and delayed the first TMA by ~0.6 us.)
    const int last_tile0    = ((u1 - 1) / p.spt) * p.spt;  // first unit of the tile that holds this CTA's last unit
    const bool last_is_fixup = (u1 > u0) && last_tile0 >= u0 && u1 < last_tile0 + p.spt;

    if (warp == W_PRODUCER_WARP) {
        // ============================================================== weight TMA producer (whole warp loops, one lane issues)
        // Weights never depend on the previous kernel in the stream: no griddepcontrol.wait here, the stream starts at once
        int wcount = 0;
        for (int u = u0; u < u1;) {
            const Segment sg = segment_at(p, u, u1);
            for (int st = sg.s0; st < sg.s1; ++st, ++wcount) {
                const int ws       = wcount % WS;
                const uint32_t wph = (wcount / WS) & 1;
                mbar_wait_warp(wempty_bar + 8 * ws, wph ^ 1);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(wfull_bar + 8 * ws, W_STAGE);
                    // k beyond K (last stage when K % 256 != 0) is zero-filled by TMA; the matching activations are
                    // zero-filled too, so those products vanish
                    tma_load_2d(w8_base + ws * W_STAGE, &map_w, wfull_bar + 8 * ws, st * STAGE_K, sg.n_tile * BLOCK_N);
                    tma_load_2d(w8_base + ws * W_STAGE + W_BOX_BYTES, &map_w, wfull_bar + 8 * ws, st * STAGE_K + 128, sg.n_tile * BLOCK_N);
                }
                __syncwarp();
                if constexpr (TRACE) {
                    if (wcount == 0 && lane == 0) tr[51] = clk64();
                }
            }
            u += sg.s1 - sg.s0;
        }
    }
    else if (warp == X_PRODUCER_WARP) {
        // ============================================================== activation TMA producer
        pdl_wait_prior_grids();  // x may be produced by the previous kernel in the stream
        int xcount = 0;
        for (int u = u0; u < u1;) {
            const Segment sg = segment_at(p, u, u1);
            for (int st = sg.s0; st < sg.s1; ++st) {
#pragma unroll 1
                for (int xi = 0; xi < XSTEPS; ++xi, ++xcount) {
                    const int xs       = xcount % XS;
                    const uint32_t xph = (xcount / XS) & 1;
                    mbar_wait_warp(xempty_bar + 8 * xs, xph ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(xfull_bar + 8 * xs, X_STAGE);
#pragma unroll
                        for (int j = 0; j < XSUB; ++j)
                            tma_load_2d(x_base + xs * X_STAGE + j * X_BOX, &map_x, xfull_bar + 8 * xs,
                                        st * STAGE_K + (xi * XSUB + j) * SUB_K, sg.t_tile * BT);
                    }
                    __syncwarp();
                }
            }
            u += sg.s1 - sg.s0;
        }
    }
    else if (warp == MMA_WARP) {
        // ============================================================== MMA issuer (whole warp loops, one lane issues)
        int acount = 0, xcount = 0, seg = 0;
        for (int u = u0; u < u1; ++seg) {
            const Segment sg = segment_at(p, u, u1);
            const int d      = seg % ND;
            mbar_wait_warp(tempty_bar + 8 * d, ((seg / ND) & 1) ^ 1);  // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_addr = tmem_base + uint32_t(d * BT);
            for (int st = sg.s0; st < sg.s1; ++st, ++acount) {
                const int a        = acount % A_STAGES;
                const uint32_t aph = (acount / A_STAGES) & 1;
                mbar_wait_warp(afull_bar + 8 * a, aph);  // 256 k of dequantised weights sit in TMEM
                tc_fence_after();
                const uint32_t a_addr = tmem_base + uint32_t(A_COL0 + a * A_STAGE_COLS);
#pragma unroll 1
                for (int xi = 0; xi < XSTEPS; ++xi, ++xcount) {
                    const int xs       = xcount % XS;
                    const uint32_t xph = (xcount / XS) & 1;
                    mbar_wait_warp(xfull_bar + 8 * xs, xph);  // activation boxes landed
                    tc_fence_after();
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < XSUB; ++j) {
                            const int sub         = xi * XSUB + j;
                            const uint64_t b_desc = make_kmajor_sw128_desc(x_base + xs * X_STAGE + j * X_BOX);
#pragma unroll
                            for (int k = 0; k < SUB_K / UMMA_K; ++k) {
                                // A: 16 k = 8 TMEM columns per UMMA; B: +32 bytes inside the swizzle atom = +2 in the (>>4) field
                                umma_f16_ts(d_addr, a_addr + uint32_t(sub * (SUB_K / 2) + k * (UMMA_K / 2)), b_desc + uint64_t(2 * k), IDESC,
                                            (st > sg.s0 || sub > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit(xempty_bar + 8 * xs);  // frees the activation stage once these MMAs have read it
                        if (xi == XSTEPS - 1) {
                            umma_commit(aempty_bar + 8 * a);  // frees the A stage
                            if (st == sg.s1 - 1)
                                umma_commit(tfull_bar + 8 * d);  // accumulator of this segment complete
                        }
                    }
                    __syncwarp();
                }
                if constexpr (TRACE) {
                    if (acount < 12 && lane == 0) tr[16 + acount] = clk64();
                }
            }
            u += sg.s1 - sg.s0;
        }
        if constexpr (TRACE) {
            if (lane == 0) tr[30] = clk64();
        }
    }
    else {
        const bool is_epi = warp < EPI_WARPS;
        if (!is_epi) {
            // ========================================================== dequantisers: smem int8 -> registers -> TMEM
            const int dw   = warp - EPI_WARPS;   // 0 .. DQW-1
            const int quad = dw & 3;             // == warp % 4: the TMEM lane quadrant this warp may touch
            const int part = dw >> 2;            // which 1/NP of the stage's k range
            const int row  = quad * 32 + lane;   // feature row inside the tile
            const int c_first = part * CH;       // first 16-byte chunk (of 16 per 256-k row)
            const int box     = c_first >> 3;    // which 128-byte-wide box
            const int cb      = c_first & 7;     // first chunk inside the box
            const uint32_t row_off   = uint32_t(box * W_BOX_BYTES + row * 128);
            const uint32_t lane_addr = uint32_t(quad * 32) << 16;
            int wcount = 0, acount = 0;
            for (int u = u0; u < u1;) {
                const Segment sg = segment_at(p, u, u1);
                uint32_t scale2  = 0;
                if constexpr (SCALE_IN_A) {
                    const int n      = sg.n_tile * BLOCK_N + row;
                    const __half sv  = (n < p.N) ? static_cast<const __half*>(p.scales)[n] : __ushort_as_half(0);
                    const __half2 s2 = __half2half2(sv);
                    scale2           = *reinterpret_cast<const uint32_t*>(&s2);
                }
                for (int st = sg.s0; st < sg.s1; ++st, ++wcount, ++acount) {
                    const int ws       = wcount % WS;
                    const uint32_t wph = (wcount / WS) & 1;
                    const int a        = acount % A_STAGES;
                    const uint32_t aph = (acount / A_STAGES) & 1;
                    mbar_wait_warp(wfull_bar + 8 * ws, wph);  // int8 stage landed
                    const uint32_t rp = w8_base + ws * W_STAGE + row_off;
                    uint4 in[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        in[j] = lds128(rp + ((((cb + j) ^ (row & 7))) << 4));  // 128B swizzle: chunk ^ (row & 7)
                    // Convert BEFORE releasing the weight stage and before waiting for the TMEM stage: (1) the stage may only be handed
                    // back to TMA once the loads have RETURNED, not merely been issued (an mbarrier arrive does not wait for outstanding
                    // shared-memory loads; the conversion consumes them), (2) in the MMA-bound regime the arithmetic then overlaps the
                    // wait for `aempty` and only the TMEM stores remain on the critical path.
                    uint32_t o[CH / 4][32];
#pragma unroll
                    for (int h = 0; h < CH / 4; ++h) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dequant16<T>(in[h * 4 + j], scale2, &o[h][8 * j]);
                    }
                    // pin the converted values here: the release below must not be scheduled ahead of the loads' consumers
#pragma unroll
                    for (int h = 0; h < CH / 4; ++h) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("mov.b32 %0, %0;" : "+r"(o[h][j])::"memory");
                    }
                    // generic-proxy reads of the stage are ordered before the async-proxy (TMA) writes that follow the release
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(wempty_bar + 8 * ws);
                    mbar_wait_warp(aempty_bar + 8 * a, aph ^ 1);  // the MMAs that read this A stage have completed
                    tc_fence_after();
                    const uint32_t a_col = uint32_t(A_COL0 + a * A_STAGE_COLS + c_first * 8);
#pragma unroll
                    for (int h = 0; h < CH / 4; ++h)
                        tmem_st_x32(tmem_base + lane_addr + a_col + uint32_t(h * 32), o[h]);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(afull_bar + 8 * a);
                    if constexpr (TRACE) {
                        if (dw == 0 && lane == 0 && acount < 12) tr[acount] = clk64();
                    }
                }
                u += sg.s1 - sg.s0;
            }
        }
        else {
            // ========================================================== epilogue warps 0..3 (TMEM lane quadrant = warp)
            const int et = threadIdx.x;  // 0..127
            pdl_wait_prior_grids();      // y / workspace may still be in use by the previous kernel in the stream
            int seg = 0;
            for (int u = u0; u < u1; ++seg) {
                const Segment sg = segment_at(p, u, u1);
                u += sg.s1 - sg.s0;
                const int d            = seg % ND;
                const bool contributor = sg.s0 > 0;  // somebody else owns this tile: dump the partial
                const bool fixup       = sg.s0 == 0 && sg.s1 < p.spt;
                if (fixup)
                    break;  // always the last segment: handled jointly below
                mbar_wait_warp(tfull_bar + 8 * d, (seg / ND) & 1);
                tc_fence_after();
                if constexpr (TRACE) {
                    if (et == 0 && seg < 4) tr[32 + 2 * seg] = clk64();
                }
                epilogue_columns<T, BT, SCALE_IN_A, (DQW == 8 ? 4 : 2)>(p, sg, contributor ? 1 : 0, tmem_base + uint32_t(d * BT), warp, lane, 0, 1, g, g,
                                                    tempty_bar + 8 * d);
                if (contributor) {
                    __threadfence();
                    epi_bar_sync();
                    if (et == 0)
                        st_release_gpu(p.flags + g, 1);
                }
                if constexpr (TRACE) {
                    if (et == 0 && seg < 4) tr[33 + 2 * seg] = clk64();
                }
            }
        }
        // ============================================================== joint finish of an owned, shared tile
        if (last_is_fixup) {
            if (!is_epi)
                pdl_wait_prior_grids();
            // the last segment and its position in the accumulator ring
            int u = u0, seg = -1;
            Segment sg{};
            while (u < u1) {
                sg = segment_at(p, u, u1);
                u += sg.s1 - sg.s0;
                ++seg;
            }
            const int d = seg % ND;
            // contributors: the CTAs g+1 .. g_last whose ranges start inside this tile
            int g_last         = g;
            const int tile_end = (sg.tile + 1) * p.spt;
            while (g_last + 1 < G && unit_begin(p, g_last + 1) < tile_end)
                ++g_last;
            const int jt = threadIdx.x;  // 0 .. JOINT_THREADS-1 (warps 0 .. 4+DQW-1)
            // wait (normally not at all: the contributors met this tile FIRST) until every contributor has published
            if (jt >= 1 && jt <= g_last - g) {
                while (ld_acquire_gpu(p.flags + g + jt) == 0) {
                }
            }
            mbar_wait_warp(tfull_bar + 8 * d, (seg / ND) & 1);
            tc_fence_after();
            joint_bar_sync<JOINT_THREADS>();
            if constexpr (TRACE) {
                if (jt == 0) tr[40] = clk64();
            }
            epilogue_columns<T, BT, SCALE_IN_A, (DQW == 8 ? 4 : 2)>(p, sg, 2, tmem_base + uint32_t(d * BT), warp & 3, lane, warp >> 2, (EPI_WARPS + DQW) / 4, g,
                                                g_last, 0u);
            joint_bar_sync<JOINT_THREADS>();  // every thread has finished reading the slots
            if (jt >= 1 && jt <= g_last - g)
                p.flags[g + jt] = 0;  // leave the workspace clean for the next call
            if constexpr (TRACE) {
                if (jt == 0) tr[41] = clk64();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if constexpr (TRACE) {
        if (threadIdx.x == 0) {
            tr[52] = clk64();
            tr[53] = gtime_ns();
            tr[54] = (unsigned long long)(u1 - u0);
        }
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

int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return (e != nullptr && e[0] != '\0') ? atoi(e) : dflt;
}

struct MapKey {
    const void* ptr;
    uint64_t d0, d1, stride;
    uint32_t b0, b1;
    int dtype, swizzle, device;
    bool operator==(const MapKey& o) const
    {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1 && dtype == o.dtype
               && swizzle == o.swizzle && device == o.device;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        size_t h = std::hash<const void*>()(k.ptr);
        auto mix = [&h](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.stride); mix(k.b0); mix(k.b1); mix(uint64_t(k.dtype)); mix(uint64_t(k.swizzle)); mix(uint64_t(k.device));
        return h;
    }
};

// 2-D tensor map over a row-major [d1][d0] matrix (d0 contiguous) with row pitch `stride` bytes.  A tensor map encodes
// nothing but (address, shape, box): the cache key holds all of them plus the device, so a recycled address with the same
// geometry yields the identical (still valid) descriptor.
int get_tensor_map(const void* ptr, CUtensorMapDataType dt, int dtype_tag, uint64_t d0, uint64_t d1, uint64_t stride, uint32_t b0,
                   uint32_t b1, CUtensorMapSwizzle swz, CUtensorMapL2promotion promo, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    MapKey key{ptr, d0, d1, stride, b0, b1, dtype_tag, int(swz) | (int(promo) << 8), dev};
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
    CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu, box %u x %u)", int(r),
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)stride, b0, b1);
        return EETQ_B200_ECUDA;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192)
        cache.clear();
    cache.emplace(key, *out);
    return EETQ_B200_OK;
}

struct TcConfig {
    int bt;       // tokens per tile (UMMA N)
    int n_tiles;  // ceil(N / 128)
    int t_tiles;  // ceil(M / bt)
    int spt;      // 256-k stages per tile
    int grid;     // persistent CTAs
    bool split;   // some tile is shared between CTAs (needs the workspace)
};

int resident_ctas()
{
    const DeviceInfo& di = device_info();
    return di.ok ? di.sm_count : 148;  // 1 CTA per SM (shared memory); never launch more than can be resident
}

TcConfig choose_config(int64_t M, int64_t N, int64_t K, bool have_workspace)
{
    TcConfig c{};
    const int sms = resident_ctas();
    c.n_tiles     = int((N + BLOCK_N - 1) / BLOCK_N);
    c.spt         = int((K + STAGE_K - 1) / STAGE_K);
    int bt        = M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256;
    // dequantisation work scales with the number of token tiles, so up to 256 tokens ride in one tile -- unless the
    // tiles are so few that every tile will be shared by several CTAs: then 128-token tiles keep the fp32 partials
    // that travel through the workspace small and give the accumulator double buffering
    if (bt == 256 && int64_t(c.n_tiles) * ((M + 255) / 256) * 2 <= sms && have_workspace)
        bt = 128;
    const int force_bt = env_int("EETQ_B200_TC_BT", 0);
    if (force_bt == 16 || force_bt == 32 || force_bt == 64 || force_bt == 128 || force_bt == 256)
        bt = force_bt;
    c.bt            = bt;
    c.t_tiles       = int((M + bt - 1) / bt);
    const int tiles = c.n_tiles * c.t_tiles;
    const int64_t U = int64_t(tiles) * c.spt;
    // Whole tiles per CTA (no partial exchange) win when one wave of tiles already fills most of the machine: the partial dump +
    // owner fix-up of the stream-K schedule costs a fixed 2-4 us.  Measured cut-offs (kbench, profiles/r02_kbench_tc.json).
    {
        const int waves   = (tiles + sms - 1) / sms;
        const double fill = double(tiles) / (double(waves) * sms);
        const bool aligned_wins = bt <= 64 ? fill >= 0.55 : (fill >= 0.85 && c.spt <= 24);
        if (aligned_wins)
            have_workspace = false;
    }
    if (have_workspace) {
        // at least two stages per CTA when there is that much work, and never more than ~64 CTAs sharing one tile (the
        // owner polls one flag per contributor with its 128 epilogue threads)
        const int64_t min_units = (c.spt + 63) / 64 > 2 ? (c.spt + 63) / 64 : 2;
        int64_t g = U >= min_units ? U / min_units : 1;
        if (g > sms) g = sms;
        c.grid  = int(g);
        c.split = true;
    }
    else {
        c.grid  = tiles < sms ? tiles : sms;
        c.split = false;
    }
    return c;
}

template <typename T, int BT, int DQW, bool TRACE>
int launch_tc(const CUtensorMap& map_w, const CUtensorMap& map_x, const TcParams& p, const TcConfig& cfg, bool pdl, cudaStream_t stream)
{
    auto kernel = w8a16_gemm_tc_kernel<T, BT, DQW, TRACE>;
    constexpr int smem = smem_bytes_for(BT);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        EB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    cudaLaunchConfig_t lc{};
    lc.gridDim          = dim3(unsigned(cfg.grid));
    lc.blockDim         = dim3(tc_threads(DQW));
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

template <typename T, int DQW, bool TRACE>
int launch_tc_bt(const CUtensorMap& map_w, const CUtensorMap& map_x, const TcParams& p, const TcConfig& cfg, bool pdl, cudaStream_t stream)
{
    switch (cfg.bt) {
        case 16: return launch_tc<T, 16, DQW, TRACE>(map_w, map_x, p, cfg, pdl, stream);
        case 32: return launch_tc<T, 32, DQW, TRACE>(map_w, map_x, p, cfg, pdl, stream);
        case 64: return launch_tc<T, 64, DQW, TRACE>(map_w, map_x, p, cfg, pdl, stream);
        case 128: return launch_tc<T, 128, DQW, TRACE>(map_w, map_x, p, cfg, pdl, stream);
        default: return launch_tc<T, 256, DQW, TRACE>(map_w, map_x, p, cfg, pdl, stream);
    }
}

}  // namespace

// Workspace = [flag region][grid x BT x 128 fp32 partial slots].  The flag region must be ZERO before the first call;
// every call leaves it zero (so one zero-initialised buffer serves all shapes and is only zeroed once).
size_t gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    if (M <= 0 || N <= 0 || K <= 0)
        return 0;
    const TcConfig c = choose_config(M, N, K, true);
    return size_t(kFlagRegionBytes) + size_t(c.grid) * c.bt * BLOCK_N * sizeof(float);
}

int launch_gemm_tc(const void* x, int64_t ldx, const int8_t* w, const void* scales, const void* bias, const void* residual,
                   int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                   size_t workspace_bytes, bool pdl, unsigned long long* trace, cudaStream_t stream)
{
    static const int nosplit_knob = env_int("EETQ_B200_TC_NOSPLIT", 0);  // diagnostics: whole tiles per CTA
    const bool have_ws = nosplit_knob == 0 && workspace != nullptr && workspace_bytes >= gemm_tc_workspace_bytes(M, N, K);
    // no (or too small a) workspace: every CTA takes whole tiles -- correct, just less evenly balanced
    const TcConfig cfg = choose_config(M, N, K, have_ws);
    if (cfg.grid > kFlagRegionBytes / int(sizeof(int))) {
        set_error("gemm_tc: grid %d exceeds the flag region", cfg.grid);
        return EETQ_B200_EINVAL;
    }
    static const int promo_knob = env_int("EETQ_B200_TC_L2PROMO", 256);
    const CUtensorMapL2promotion wpromo = promo_knob == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                         : promo_knob == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                            : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUtensorMap map_w, map_x;
    // weights: [N rows][K bytes]; one box = 128 rows x 128 bytes, 128B-swizzled so that the dequant warps' row-per-lane
    // 16-byte reads are bank-conflict free; L2 promotion 256 B makes the two boxes of a stage one DRAM burst per row
    if (int rc = get_tensor_map(w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 100, uint64_t(K), uint64_t(N), uint64_t(K), 128, BLOCK_N,
                                CU_TENSOR_MAP_SWIZZLE_128B, wpromo, &map_w))
        return rc;
    const CUtensorMapDataType xdt = dtype == EETQ_B200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (int rc = get_tensor_map(x, xdt, dtype, uint64_t(K), uint64_t(M), uint64_t(ldx) * 2, SUB_K, uint32_t(cfg.bt),
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, &map_x))
        return rc;

    TcParams p{};
    p.scales       = scales;
    p.bias         = bias;
    p.residual     = residual;
    p.ldr          = ldr;
    p.y            = y;
    p.ldy          = ldy;
    p.M            = int(M);
    p.N            = int(N);
    p.K            = int(K);
    p.n_tiles      = cfg.n_tiles;
    p.t_tiles      = cfg.t_tiles;
    p.spt          = cfg.spt;
    {
        const int total = cfg.split ? cfg.n_tiles * cfg.t_tiles * cfg.spt : cfg.n_tiles * cfg.t_tiles;
        p.units_q       = total / cfg.grid;
        p.units_r       = total % cfg.grid;
    }
    p.tile_aligned = cfg.split ? 0 : 1;
    p.trace        = trace;
    if (cfg.split) {
        p.flags = static_cast<int*>(workspace);
        p.slots = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kFlagRegionBytes);
    }
    // 8 dequant warps (two per TMEM lane quadrant).  16 were measured within +-3 % at every shape and are not built.
    constexpr int kDqw = 8;
    if (trace != nullptr) {
        if (dtype != EETQ_B200_F16) {
            set_error("gemm_tc: the instrumented build exists for fp16 only");
            return EETQ_B200_EINVAL;
        }
        return launch_tc_bt<__half, kDqw, true>(map_w, map_x, p, cfg, pdl, stream);
    }
    if (dtype == EETQ_B200_F16)
        return launch_tc_bt<__half, kDqw, false>(map_w, map_x, p, cfg, pdl, stream);
    return launch_tc_bt<__nv_bfloat16, kDqw, false>(map_w, map_x, p, cfg, pdl, stream);
}

int gemm_tc_trace_slots() { return TRACE_SLOTS; }
int gemm_tc_grid_for(int64_t M, int64_t N, int64_t K) { return choose_config(M, N, K, true).grid; }

}  // namespace eetq_b200
