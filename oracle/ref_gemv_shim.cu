// ref_gemv_shim.cu -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product library).
//
// C wrapper around the reference decode GEMV (TensorRT-LLM weightOnlyBatchedGemv as vendored by
// EETQ: /root/reference/csrc/weightOnlyBatchedGemv/kernelLauncher.cu + ...Bs{1..4}Int8b.cu),
// compiled UNMODIFIED for sm_100a by oracle/Makefile into oracle/_ref/libref_gemv.so.
// It is the only reference kernel that can execute on a B200 (the CUTLASS path throws for
// sm >= 90, fpA_intB_gemm_template.h:433-436), valid for M <= 4, and it expects weights in the
// reference sm80 interleaved layout (produced by libref_oracle.so / oracle.ref_layout()).
// Mirrors the call made at /root/reference/csrc/cutlass_kernels/fpA_intB_gemm_wrapper.cu:154-159.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "weightOnlyBatchedGemv/kernelLauncher.h"

static int ref_gemv_any(const void* x, const void* w_ref_layout, const void* scales, void* y, int m, int n, int k,
                        tensorrt_llm::kernels::WeightOnlyQuantType qtype, void* stream)
{
    namespace trt = tensorrt_llm::kernels;
    if (m < 1 || m > 4)
        return 2;
    trt::WeightOnlyParams params{static_cast<const uint8_t*>(w_ref_layout),
                                 scales,
                                 nullptr,
                                 x,
                                 nullptr,
                                 nullptr,
                                 y,
                                 m,
                                 n,
                                 k,
                                 0,
                                 qtype,
                                 trt::WeightOnlyType::PerChannel,
                                 trt::WeightOnlyActivationFunctionType::Identity,
                                 trt::WeightOnlyActivationType::FP16};
    trt::weight_only_batched_gemv_launcher(params, static_cast<cudaStream_t>(stream));
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int ref_w8a16_gemv(const void* x, const void* w_ref_layout, const void* scales, void* y, int m, int n, int k,
                              void* stream)
{
    return ref_gemv_any(x, w_ref_layout, scales, y, m, n, k, tensorrt_llm::kernels::WeightOnlyQuantType::Int8b, stream);
}

// same kernel template, Int4b weights in the reference's int4 layout (kernel.h:68-116); not selectable from the reference's
// Python (fpA_intB_gemm_wrapper.cu:154-159 hard-codes Int8b) but compiled into its extension
extern "C" int ref_w4a16_gemv(const void* x, const void* w_ref_layout, const void* scales, void* y, int m, int n, int k,
                              void* stream)
{
    return ref_gemv_any(x, w_ref_layout, scales, y, m, n, k, tensorrt_llm::kernels::WeightOnlyQuantType::Int4b, stream);
}
