// int4_host.cc -- TEST INFRASTRUCTURE: runs the product's own int4 index arithmetic (eetq_b200/csrc/int4_layout.cuh, the
// functions the CUDA kernels call) on the CPU so that tests/test_int4_host.py can compare it with the oracle without a GPU.
#include <cstddef>
#include <cstdint>

#include "../../eetq_b200/csrc/int4_layout.cuh"

using namespace eetq_b200;

extern "C" int host_nibble_layout(int mode, const uint8_t* src, int64_t K, int64_t N, uint32_t* dst)
{
    const int64_t words = K * N / 8;
    for (int64_t wi = 0; wi < words; ++wi) {
        switch (mode) {
            case NIB_PACK4: dst[wi] = nibble_layout_word<NIB_PACK4>(src, K, N, wi); break;
            case NIB_UNPACK4: dst[wi] = nibble_layout_word<NIB_UNPACK4>(src, K, N, wi); break;
            case NIB_FROM_REF4: dst[wi] = nibble_layout_word<NIB_FROM_REF4>(src, K, N, wi); break;
            case NIB_TO_REF4: dst[wi] = nibble_layout_word<NIB_TO_REF4>(src, K, N, wi); break;
            default: return -1;
        }
    }
    return 0;
}

extern "C" void host_widen4to8(const uint32_t* src, int64_t words, uint32_t* dst)
{
    for (int64_t i = 0; i < words; ++i)
        widen4to8_word(src[i], dst[2 * i], dst[2 * i + 1]);
}

// eight biased nibbles per word, one per byte -> b200 int4 words
extern "C" void host_pack_words(const uint8_t* u8, int64_t words, uint32_t* dst)
{
    for (int64_t i = 0; i < words; ++i)
        dst[i] = b200_pack_word(u8 + 8 * i);
}
