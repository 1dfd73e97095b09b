/*
 * w8a16_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar CPU restatement of the one hot path of NetEase-FuXi/EETQ that this repository
 * replaces (per-output-channel symmetric INT8 weight quantisation, the sm75-sm89 weight layout,
 * and the w8a16 GEMM arithmetic).  It is the *checker* for the CUDA kernels: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (eetq_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity pin: the reference ships no golden vectors; this file is pinned instead
 * against the reference C++ itself, compiled unmodified into oracle/_ref/libref_oracle.so
 * (tests/test_oracle.py, tests/golden/).
 *
 * Each function names the reference lines whose *behaviour* it restates; the code is written
 * from the closed forms in SURVEY.md §8a, not from the reference's loop structure.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef _Float16 f16;

/* ---------------------------------------------------------------------------------------------
 * Q1  per-column symmetric int8 quantisation.
 * Reference: csrc/cutlass_kernels/cutlass_preprocessors.cc:608-649
 *   amax[n]  = max_k |float(w[k,n])|                                  (:623-628)
 *   s32[n]   = amax[n] * (1/128)            in fp32                   (:610, :633)
 *   stored scale = ComputeType(s32[n])      (fp16 or fp32)            (:634)
 *   q[k,n]   = int8(clamp(round(float(w)/s32[n]), -128, 127))         (:644-648)
 *   round() is C round-half-away-from-zero; the division uses the FP32 scale, not the stored one.
 *   An all-zero column gives 0/0 = NaN; std::min(127.f, NaN) == 127.f on the reference's
 *   comparison order, so the reference emits q = 127 with scale 0.  We reproduce that.
 * ------------------------------------------------------------------------------------------- */
static inline int8_t quant_one(float w, float s32)
{
    float r = roundf(w / s32);
    float lo = (r < 127.f) ? r : 127.f;      /* NaN < 127 is false -> 127, like std::min(127.f, r) */
    float hi = (-128.f < lo) ? lo : -128.f;  /* std::max(-128.f, lo) */
    return (int8_t)hi;
}

void oracle_quantize_f32(const float* w, size_t K, size_t N, int8_t* q_kn, float* scales_f32)
{
    for (size_t n = 0; n < N; ++n) {
        float amax = 0.f;
        for (size_t k = 0; k < K; ++k) {
            float a = fabsf(w[k * N + n]);
            if (amax < a) amax = a;
        }
        scales_f32[n] = amax * (1.f / 128.f);
    }
    for (size_t k = 0; k < K; ++k)
        for (size_t n = 0; n < N; ++n)
            q_kn[k * N + n] = quant_one(w[k * N + n], scales_f32[n]);
}

/* fp16 weights: arithmetic identical (everything is promoted to fp32); the stored scale is fp16. */
void oracle_quantize_f16(const f16* w, size_t K, size_t N, int8_t* q_kn, f16* scales_f16, float* scales_f32)
{
    for (size_t n = 0; n < N; ++n) {
        float amax = 0.f;
        for (size_t k = 0; k < K; ++k) {
            float a = fabsf((float)w[k * N + n]);
            if (amax < a) amax = a;
        }
        scales_f32[n] = amax * (1.f / 128.f);
        scales_f16[n] = (f16)scales_f32[n];
    }
    for (size_t k = 0; k < K; ++k)
        for (size_t n = 0; n < N; ++n)
            q_kn[k * N + n] = quant_one((float)w[k * N + n], scales_f32[n]);
}

/* ---------------------------------------------------------------------------------------------
 * Q3  the reference's weight layout for sm75..sm89, int8.
 * Reference: preprocess_weights_for_mixed_gemm, cutlass_preprocessors.cc:497-534
 *   = permute_B_rows (:137-195) o transpose (:201-320) o 2-column interleave (:432-495)
 *     o +128 bias & byte swizzle (:337-358).
 * Closed form: the K*N output bytes, viewed as [N/2][K/64][2][4][16]
 * = (column pair, k tile, column-in-pair c, 16-group g, position p), hold
 *   uint8(q[k, n] + 128),  n = 2*pair + c,  k = 64*ktile + 16*g + (p>>1) + 8*(p&1).
 * Needs K % 64 == 0 and N % 64 == 0 (returns -1 otherwise; the reference aborts, :230/:455).
 * ------------------------------------------------------------------------------------------- */
int oracle_ref_layout(const int8_t* q_kn, size_t K, size_t N, uint8_t* out)
{
    if (K % 64 || N % 64) return -1;
    size_t o = 0;
    for (size_t pair = 0; pair < N / 2; ++pair)
        for (size_t kt = 0; kt < K / 64; ++kt)
            for (size_t c = 0; c < 2; ++c)
                for (size_t g = 0; g < 4; ++g)
                    for (size_t p = 0; p < 16; ++p) {
                        size_t n = 2 * pair + c;
                        size_t k = 64 * kt + 16 * g + (p >> 1) + 8 * (p & 1);
                        out[o++] = (uint8_t)((int)q_kn[k * N + n] + 128);
                    }
    return 0;
}

/* inverse of the above: reference-layout bytes -> row-major int8 [K,N] */
int oracle_ref_layout_inv(const uint8_t* in, size_t K, size_t N, int8_t* q_kn)
{
    if (K % 64 || N % 64) return -1;
    size_t o = 0;
    for (size_t pair = 0; pair < N / 2; ++pair)
        for (size_t kt = 0; kt < K / 64; ++kt)
            for (size_t c = 0; c < 2; ++c)
                for (size_t g = 0; g < 4; ++g)
                    for (size_t p = 0; p < 16; ++p) {
                        size_t n = 2 * pair + c;
                        size_t k = 64 * kt + 16 * g + (p >> 1) + 8 * (p & 1);
                        q_kn[k * N + n] = (int8_t)((int)in[o++] - 128);
                    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * The layout THIS repository's kernels consume ("b200 layout"):
 * biased bytes, output-feature-major:  out[n*K + k] = uint8(q[k, n] + 128)   (nn.Linear.weight order).
 * Not a reference function -- it is restated here so tests can check the CUDA packer bit-for-bit.
 * ------------------------------------------------------------------------------------------- */
void oracle_b200_layout(const int8_t* q_kn, size_t K, size_t N, uint8_t* out_nk)
{
    for (size_t n = 0; n < N; ++n)
        for (size_t k = 0; k < K; ++k)
            out_nk[n * K + k] = (uint8_t)((int)q_kn[k * N + n] + 128);
}

/* ---------------------------------------------------------------------------------------------
 * K1/G1  the w8a16 GEMM arithmetic (the parity target).
 * Reference: csrc/cutlass_extensions/.../warp/mma_tensorop_dequantizer.h:259-274 (fp16 w*s, one
 * rounding per weight), default_fpA_intB_traits.h:110 (fp32 accumulation), fpA_intB_gemm_template.h:133
 * (fp16 output):   y[m,n] = fp16( sum_k fp32(x[m,k]) * fp32( fp16( fp16(q[k,n]) * s[n] ) ) ).
 * Summation order here is k-ascending in fp32 (the reference's order is tile dependent), hence
 * the 1e-3 norm-relative parity tolerance rather than bit equality.
 * ------------------------------------------------------------------------------------------- */
void oracle_gemm_f16(const f16* x, const int8_t* q_kn, const f16* s, f16* y, size_t M, size_t N, size_t K)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n) {
        for (size_t m = 0; m < M; ++m) {
            float acc = 0.f;
            for (size_t k = 0; k < K; ++k) {
                f16 wd = (f16)((f16)q_kn[k * N + n] * s[n]);
                acc += (float)x[m * K + k] * (float)wd;
            }
            y[m * N + n] = (f16)acc;
        }
    }
}

/* double-precision version of the same sum (ground truth for tolerance studies) */
void oracle_gemm_f64(const f16* x, const int8_t* q_kn, const f16* s, double* y, size_t M, size_t N, size_t K)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n) {
        for (size_t m = 0; m < M; ++m) {
            double acc = 0.0;
            for (size_t k = 0; k < K; ++k) {
                f16 wd = (f16)((f16)q_kn[k * N + n] * s[n]);
                acc += (double)x[m * K + k] * (double)wd;
            }
            y[m * N + n] = acc;
        }
    }
}
