// ref_shim.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Glue that lets the UNMODIFIED reference quantiser / weight preprocessor
// (/root/reference/csrc/cutlass_kernels/cutlass_preprocessors.cc) run on a box with no
// Ampere GPU.  The reference asks the CUDA runtime for the compute capability
// (csrc/utils/cuda_utils.h:63-72) and throws for arch >= 90 (cutlass_preprocessors.cc:113-128),
// so we answer "8.0" from two local stubs (the library is linked with -Bsymbolic so these win
// over any libcudart already in the process).  The extern "C" wrappers catch the reference's
// std::runtime_error so a ctypes caller sees an error code instead of abort().
//
// Built by oracle/Makefile into oracle/_ref/libref_oracle.so; sources are compiled where they
// lie under /root/reference -- nothing is copied into this repository.
#include <cuda_runtime_api.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <exception>
#include <vector>

#include "cutlass_kernels/cutlass_preprocessors.h"

extern "C" cudaError_t cudaGetDevice(int* device)
{
    *device = 0;
    return cudaSuccess;
}

extern "C" cudaError_t cudaDeviceGetAttribute(int* value, cudaDeviceAttr attr, int)
{
    *value = (attr == cudaDevAttrComputeCapabilityMajor) ? 8 : 0;
    return cudaSuccess;
}

namespace ft = fastertransformer;

// quant_weights(w[K,N] fp16) -> processed (reference sm80 layout), unprocessed row-major [K,N], scales fp16[N]
extern "C" int ref_quant_fp16(int8_t* processed, int8_t* unprocessed, void* scales, const void* w, size_t K, size_t N)
{
    try {
        ft::symmetric_quantize<half, half>(processed, unprocessed, static_cast<half*>(scales),
                                           static_cast<const half*>(w), {K, N}, ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_quant_fp16: %s\n", e.what());
        return 1;
    }
}

// same for fp32 weights / fp32 scales
extern "C" int ref_quant_fp32(int8_t* processed, int8_t* unprocessed, float* scales, const float* w, size_t K, size_t N)
{
    try {
        ft::symmetric_quantize<float, float>(processed, unprocessed, scales, w, {K, N},
                                             ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_quant_fp32: %s\n", e.what());
        return 1;
    }
}

// preprocess_weights(int8 row-major [K,N]) -> reference sm80 layout
extern "C" int ref_preprocess(int8_t* out, const int8_t* in, size_t K, size_t N)
{
    try {
        ft::preprocess_weights(out, in, K, N, /*is_int4=*/false, /*arch=*/80);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_preprocess: %s\n", e.what());
        return 1;
    }
}

// ---- packed int4 (QuantType::PACKED_INT4_WEIGHT_ONLY): reachable from the reference's Python as
// quant_weights(w, torch.quint4x2, ...) and preprocess_weights(w, is_int4=True)  (csrc/eetpy.cpp:11-17) ----
// quant_weights(w[K,N] fp16, quint4x2) -> processed / unprocessed packed [K, N/2] (two int4 per byte), scales fp16[N]
extern "C" int ref_quant4_fp16(int8_t* processed, int8_t* unprocessed, void* scales, const void* w, size_t K, size_t N)
{
    try {
        ft::symmetric_quantize<half, half>(processed, unprocessed, static_cast<half*>(scales),
                                           static_cast<const half*>(w), {K, N}, ft::QuantType::PACKED_INT4_WEIGHT_ONLY);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_quant4_fp16: %s\n", e.what());
        return 1;
    }
}

extern "C" int ref_quant4_fp32(int8_t* processed, int8_t* unprocessed, float* scales, const float* w, size_t K, size_t N)
{
    try {
        ft::symmetric_quantize<float, float>(processed, unprocessed, scales, w, {K, N},
                                             ft::QuantType::PACKED_INT4_WEIGHT_ONLY);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_quant4_fp32: %s\n", e.what());
        return 1;
    }
}

// preprocess_weights(packed int4 row-major, K x N ELEMENTS = K*N/2 bytes, is_int4=true) -> reference sm80 int4 layout
extern "C" int ref_preprocess4(int8_t* out, const int8_t* in, size_t K, size_t N)
{
    try {
        ft::preprocess_weights(out, in, K, N, /*is_int4=*/true, /*arch=*/80);
        return 0;
    }
    catch (std::exception& e) {
        fprintf(stderr, "ref_preprocess4: %s\n", e.what());
        return 1;
    }
}
