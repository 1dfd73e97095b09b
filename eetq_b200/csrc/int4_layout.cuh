// int4_layout.cuh -- index arithmetic of the packed-int4 layouts (host + device).
//
// Three byte formats of a logical int4 matrix q[K][N] (values -8..7), nibble index i = byte i/2, bits 4*(i&1):
//   packed row-major  nibble k*N + n           = q[k,n] & 15 (two's complement)       -- the reference's "unprocessed" tensor,
//                                                                                         cutlass_preprocessors.cc:651-669
//   reference sm80    nibble (((n4*(K/64) + kt)*4 + c)*8 + v)*8 + d = q[perm(64kt + 8v + e(d)), 4*n4 + c] + 8
//                     e(d) = d < 4 ? 2d : 2(d-4)+1;  perm swaps the two 2-bit fields of k%32 (t = 8a + 2b + c0 -> 8b + 2a + c0,
//                     an involution) -- closed form of preprocess_weights_for_mixed_gemm for PACKED_INT4_WEIGHT_ONLY
//                     (cutlass_preprocessors.cc:497-534 = :137-195 o :201-320 o :432-495 o :360-418), pinned against the
//                     compiled reference through oracle/w8a16_oracle.py::ref_layout4
//   b200 int4         nibble (n*(K/8) + k/8)*8 + pos(k%8) = q[k,n] + 8,  pos(r) = r even ? r/2 : 4 + r/2
// The functions are __host__ __device__ so that tests/test_int4_host.py runs the kernels' own arithmetic on the CPU.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define EB_HD __host__ __device__ __forceinline__
#else
#define EB_HD inline
#endif

namespace eetq_b200 {

enum { NIB_PACK4 = 0, NIB_UNPACK4 = 1, NIB_FROM_REF4 = 2, NIB_TO_REF4 = 3 };

// position of k-offset r (0..7) inside a 32-bit word of the b200 int4 layout, and its inverse
EB_HD int b200_nib(int r) { return (r & 1) ? 4 + (r >> 1) : (r >> 1); }
EB_HD int b200_koff(int d) { return (d < 4) ? 2 * d : 2 * (d - 4) + 1; }

// eight biased nibbles (one per byte, k ascending) -> one word of the b200 int4 layout
EB_HD uint32_t b200_pack_word(const uint8_t* u8)
{
    uint32_t v = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
        v |= uint32_t(u8[r] & 15u) << (4 * b200_nib(r));
    return v;
}

EB_HD int64_t ref4_perm(int64_t k)
{
    const int t = int(k & 31);
    return (k & ~int64_t(31)) + 8 * ((t >> 1) & 3) + 2 * (t >> 3) + (t & 1);
}
EB_HD int64_t ref4_nibble_index(int64_t k, int64_t n, int64_t K)
{
    const int64_t kp = ref4_perm(k);  // storage position k' with perm(k') = k (perm is its own inverse)
    return ((((n >> 2) * (K >> 6) + (kp >> 6)) * 4 + (n & 3)) * 8 + ((kp >> 3) & 7)) * 8 + b200_nib(int(kp & 7));
}
EB_HD int64_t b200_nibble_index(int64_t k, int64_t n, int64_t K) { return (n * (K >> 3) + (k >> 3)) * 8 + b200_nib(int(k & 7)); }

// output word `wi` of a layout conversion, assembled from the eight source nibbles the mapping names
template <int MODE>
EB_HD uint32_t nibble_layout_word(const uint8_t* src, int64_t K, int64_t N, int64_t wi)
{
    uint32_t out = 0;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        int64_t k, n, si;
        if (MODE == NIB_PACK4 || MODE == NIB_FROM_REF4) {
            // destination = b200 int4: word wi = (n, j), nibble d holds k = 8j + koff(d)
            n  = wi / (K >> 3);
            k  = 8 * (wi % (K >> 3)) + b200_koff(d);
            si = (MODE == NIB_PACK4) ? k * N + n : ref4_nibble_index(k, n, K);
        }
        else if (MODE == NIB_UNPACK4) {
            // destination = packed row-major: word wi = (k, 8 columns), nibble d is column 8*(wi % (N/8)) + d
            k  = wi / (N >> 3);
            n  = 8 * (wi % (N >> 3)) + d;
            si = b200_nibble_index(k, n, K);
        }
        else {
            // destination = reference layout: word wi = (n4, kt, c, v), nibble d holds storage position k' = 64kt + 8v + koff(d)
            const int v      = int(wi & 7);
            const int c      = int((wi >> 3) & 3);
            const int64_t kt = (wi >> 5) % (K >> 6);
            const int64_t n4 = (wi >> 5) / (K >> 6);
            n  = 4 * n4 + c;
            k  = ref4_perm(64 * kt + 8 * v + b200_koff(d));
            si = b200_nibble_index(k, n, K);
        }
        uint32_t nib = (uint32_t(src[si >> 1]) >> (4 * int(si & 1))) & 15u;
        if (MODE == NIB_PACK4 || MODE == NIB_UNPACK4)
            nib ^= 8u;  // two's complement <-> biased by +8
        out |= nib << (4 * d);
    }
    return out;
}

// PRMT semantics for selectors without the sign-replicate bit (host fallback of __byte_perm)
EB_HD uint32_t byte_perm_hd(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t ab = (uint64_t(b) << 32) | a;
    uint32_t r        = 0;
    for (int i = 0; i < 4; ++i)
        r |= uint32_t((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

// one b200 int4 word (8 k) -> two words of the b200 int8 layout: byte = u4 + 120 = q + 128
EB_HD void widen4to8_word(uint32_t w, uint32_t& k0123, uint32_t& k4567)
{
    // bytes of lo = nibbles (0,2,4,6) = k (0,4,1,5); bytes of hi = nibbles (1,3,5,7) = k (2,6,3,7)
    const uint32_t lo = w & 0x0f0f0f0fu;
    const uint32_t hi = (w >> 4) & 0x0f0f0f0fu;
    k0123 = byte_perm_hd(lo, hi, 0x6420) + 0x78787878u;
    k4567 = byte_perm_hd(lo, hi, 0x7531) + 0x78787878u;
}

}  // namespace eetq_b200
