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

/* ---------------------------------------------------------------------------------------------
 * Packed int4 (QuantType::PACKED_INT4_WEIGHT_ONLY) -- reachable from the reference's Python as
 * quant_weights(w, torch.quint4x2, ..) and preprocess_weights(w, is_int4=True) (csrc/eetpy.cpp:11-17).
 * Quantiser: cutlass_preprocessors.cc:608-669
 *   s32[n] = amax[n] * (1/8);  v = (int)round(float(w)/s32[n]);  q = max(-8, min(7, v));
 *   two values per byte: low nibble = even column, high nibble = odd column.
 *   The float->int conversion happens BEFORE the clamp; for NaN (0/0 in an all-zero column, or a NaN weight)
 *   x86-64's cvttss2si returns INT_MIN, which clamps to -8.  Written out explicitly here so the restatement
 *   does not depend on the host's conversion behaviour.
 * ------------------------------------------------------------------------------------------- */
static inline int quant_one4(float w, float s32)
{
    float r = roundf(w / s32);
    if (r != r) return -8; /* NaN -> INT_MIN -> -8 */
    int v = (r > 100.f) ? 100 : (r < -100.f) ? -100 : (int)r;
    return v < -8 ? -8 : (v > 7 ? 7 : v);
}

void oracle_quantize4_f32(const float* w, size_t K, size_t N, uint8_t* packed_kn2, float* scales_f32)
{
    for (size_t n = 0; n < N; ++n) {
        float amax = 0.f;
        for (size_t k = 0; k < K; ++k) {
            float a = fabsf(w[k * N + n]);
            if (amax < a) amax = a;
        }
        scales_f32[n] = amax * (1.f / 8.f);
    }
    for (size_t k = 0; k < K; ++k)
        for (size_t j = 0; j < N / 2; ++j) {
            int a = quant_one4(w[k * N + 2 * j], scales_f32[2 * j]);
            int b = quant_one4(w[k * N + 2 * j + 1], scales_f32[2 * j + 1]);
            packed_kn2[k * (N / 2) + j] = (uint8_t)((a & 15) | ((b & 15) << 4));
        }
}

void oracle_quantize4_f16(const f16* w, size_t K, size_t N, uint8_t* packed_kn2, f16* scales_f16, float* scales_f32)
{
    for (size_t n = 0; n < N; ++n) {
        float amax = 0.f;
        for (size_t k = 0; k < K; ++k) {
            float a = fabsf((float)w[k * N + n]);
            if (amax < a) amax = a;
        }
        scales_f32[n] = amax * (1.f / 8.f);
        scales_f16[n] = (f16)scales_f32[n];
    }
    for (size_t k = 0; k < K; ++k)
        for (size_t j = 0; j < N / 2; ++j) {
            int a = quant_one4((float)w[k * N + 2 * j], scales_f32[2 * j]);
            int b = quant_one4((float)w[k * N + 2 * j + 1], scales_f32[2 * j + 1]);
            packed_kn2[k * (N / 2) + j] = (uint8_t)((a & 15) | ((b & 15) << 4));
        }
}

static inline int get_nib(const uint8_t* p, size_t i) { return (p[i >> 1] >> (4 * (i & 1))) & 15; }
static inline void put_nib(uint8_t* p, size_t i, int v)
{
    p[i >> 1] = (uint8_t)((p[i >> 1] & (0xF0 >> (4 * (i & 1)))) | ((v & 15) << (4 * (i & 1))));
}

/* Reference int4 layout for sm75..sm89 (preprocess_weights_for_mixed_gemm, cutlass_preprocessors.cc:497-534, int4 case):
 *   row permutation in groups of 32 k (:137-195), element transpose (:201-320), ColumnMajorTileInterleave<64,4> (:432-495),
 *   +8 bias and nibble interleave inside each 32-bit word (:360-418).
 * Written per OUTPUT nibble: nibble index o = (((n4*(K/64) + kt)*4 + c)*8 + v)*8 + d holds
 *   (q[perm(64*kt + 8*v + e), 4*n4 + c] + 8) & 15,  e = d<4 ? 2d : 2(d-4)+1,
 *   perm(k') = 32*(k'/32) + 8*((t%8)/2) + t%2 + 2*(t/8),  t = k' % 32.
 * Input: packed row-major [K][N/2]; output K*N/2 bytes. */
int oracle_ref_layout4(const uint8_t* packed_kn2, size_t K, size_t N, uint8_t* out)
{
    if (K % 64 || N % 64) return -1;
    memset(out, 0, K * N / 2);
    size_t o = 0;
    for (size_t n4 = 0; n4 < N / 4; ++n4)
        for (size_t kt = 0; kt < K / 64; ++kt)
            for (size_t c = 0; c < 4; ++c)
                for (size_t v = 0; v < 8; ++v)
                    for (size_t d = 0; d < 8; ++d, ++o) {
                        size_t e  = d < 4 ? 2 * d : 2 * (d - 4) + 1;
                        size_t kp = 64 * kt + 8 * v + e;
                        size_t t  = kp % 32;
                        size_t k  = 32 * (kp / 32) + 8 * ((t % 8) / 2) + t % 2 + 2 * (t / 8);
                        size_t n  = 4 * n4 + c;
                        int q     = get_nib(packed_kn2, k * N + n); /* two's-complement nibble */
                        put_nib(out, o, (q + 8) & 15);
                    }
    return 0;
}

/* The int4 layout THIS repository's kernels consume: output-feature-major rows of K/2 bytes; in each
 * 32-bit word (8 consecutive k) nibble p < 4 holds q[8j+2p]+8 and nibble 4+p holds q[8j+2p+1]+8. */
void oracle_b200_layout4(const uint8_t* packed_kn2, size_t K, size_t N, uint8_t* out)
{
    memset(out, 0, K * N / 2);
    for (size_t n = 0; n < N; ++n)
        for (size_t k = 0; k < K; ++k) {
            size_t j = k / 8, r = k % 8;
            size_t nib = (r & 1) ? 4 + r / 2 : r / 2;
            int q = get_nib(packed_kn2, k * N + n);
            put_nib(out, (n * (K / 8) + j) * 8 + nib, (q + 8) & 15);
        }
}

/* w4a16 GEMM arithmetic: identical to oracle_gemm_f16 with q in [-8, 7] read from the packed row-major form
 * (the reference's Int4b decode kernel dequantises the same way: fp16(q) * s, kernel.h:68-116, :355-377). */
void oracle_gemm4_f16(const f16* x, const uint8_t* packed_kn2, const f16* s, f16* y, size_t M, size_t N, size_t K)
{
#pragma omp parallel for schedule(static)
    for (size_t n = 0; n < N; ++n) {
        for (size_t m = 0; m < M; ++m) {
            float acc = 0.f;
            for (size_t k = 0; k < K; ++k) {
                int q  = get_nib(packed_kn2, k * N + n);
                q      = q >= 8 ? q - 16 : q;
                f16 wd = (f16)((f16)q * s[n]);
                acc += (float)x[m * K + k] * (float)wd;
            }
            y[m * N + n] = (f16)acc;
        }
    }
}
