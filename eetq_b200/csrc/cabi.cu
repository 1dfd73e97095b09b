// cabi.cu -- the extern "C" boundary of libeetq_b200.so (declared in include/eetq_b200.h).
//
// Stands in for the reference's torch/pybind boundary:
//   /root/reference/csrc/eetpy.cpp:7-19                                  (the 4 hot-path symbols)
//   /root/reference/csrc/cutlass_kernels/fpA_intB_gemm_wrapper.cu:28-202 (marshalling + M-based dispatch)
// Differences by design: raw pointers + explicit stream instead of torch::Tensor; argument validation the
// reference lacks; error codes + thread-local message instead of C++ exceptions/asserts;
// per-device attributes cached once instead of re-queried per forward (fpA_intB_gemm_template.h:390-397).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "int4_layout.cuh"

namespace eetq_b200 {

namespace {
thread_local char tls_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(uint64_t(n), std::memory_order_relaxed); }

const DeviceInfo& device_info()
{
    static DeviceInfo infos[64];
    static std::once_flag flags[64];
    static DeviceInfo bad;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return bad;
    std::call_once(flags[dev], [dev]() {
        DeviceInfo d;
        bool ok = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
        ok      = ok && cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess;
        ok      = ok && cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess;
        ok      = ok && cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess;
        d.ok    = ok;
        infos[dev] = d;
    });
    return infos[dev];
}

int check_arch()
{
    const DeviceInfo& di = device_info();
    if (!di.ok) {
        set_error("could not query the current CUDA device");
        return EETQ_B200_ECUDA;
    }
    if (di.cc_major != 10) {
        set_error("libeetq_b200 is built for sm_100a only; current device is sm_%d%d", di.cc_major, di.cc_minor);
        return EETQ_B200_EARCH;
    }
    return EETQ_B200_OK;
}

int check_kn(const char* who, int64_t K, int64_t N)
{
    if (K <= 0 || N <= 0 || (K % 64) != 0 || (N % 64) != 0) {
        set_error("%s: K (%lld) and N (%lld) must be positive multiples of 64", who, (long long)K, (long long)N);
        return EETQ_B200_EINVAL;
    }
    if (K > (1 << 20) || N > (1 << 24)) {
        set_error("%s: K (%lld) or N (%lld) too large", who, (long long)K, (long long)N);
        return EETQ_B200_EINVAL;
    }
    return EETQ_B200_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// shared argument validation of every w8a16 forward entry point
int check_forward_args(const char* who, const void* x, int64_t ldx, const void* w, const void* scales, const void* y, int64_t ldy,
                       int64_t M, int64_t N, int64_t K, int dtype, int64_t n_out)
{
    if (n_out < 0)
        n_out = N;  // output row width (N / 2 for the SiLU*up pair epilogue)
    EB_CHECK_ARG(x && w && scales && y, "%s: null pointer argument", who);
    EB_CHECK_ARG(dtype == EETQ_B200_F16 || dtype == EETQ_B200_BF16, "%s: dtype must be F16 or BF16 (got %d)", who, dtype);
    EB_CHECK_ARG(M >= 0 && M <= (int64_t(1) << 24), "%s: bad M=%lld", who, (long long)M);
    if (int rc = check_kn(who, K, N))
        return rc;
    EB_CHECK_ARG(ldx >= K && ldy >= n_out, "%s: ldx (%lld) < K or ldy (%lld) < output width (%lld)", who, (long long)ldx, (long long)ldy,
                 (long long)n_out);
    EB_CHECK_ARG((ldx % 8) == 0 && (ldy % 8) == 0, "%s: ldx and ldy must be multiples of 8 elements", who);
    EB_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y), "%s: x, w and y must be 16-byte aligned", who);
    return EETQ_B200_OK;
}

}  // namespace eetq_b200

using namespace eetq_b200;

extern "C" {

const char* eetq_b200_last_error(void) { return tls_error; }
int eetq_b200_version(void) { return EETQ_B200_VERSION; }
uint64_t eetq_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int eetq_b200_quantize(const void* w_kn, int w_dtype, int64_t K, int64_t N, int8_t* q_b200, void* scales, float* s32,
                       int8_t* q_kn, void* stream)
{
    EB_CHECK_ARG(w_kn && q_b200 && scales && s32, "quantize: null pointer argument");
    EB_CHECK_ARG(aligned16(w_kn) && aligned16(q_b200) && (q_kn == nullptr || aligned16(q_kn)),
                 "quantize: pointers must be 16-byte aligned");
    if (int rc = check_kn("quantize", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_quantize(w_kn, w_dtype, K, N, q_b200, scales, s32, q_kn, static_cast<cudaStream_t>(stream));
}

int eetq_b200_pack(const int8_t* q_kn, int64_t K, int64_t N, int8_t* q_b200, void* stream)
{
    EB_CHECK_ARG(q_kn && q_b200, "pack: null pointer argument");
    EB_CHECK_ARG(aligned16(q_kn) && aligned16(q_b200), "pack: pointers must be 16-byte aligned");
    if (int rc = check_kn("pack", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_transpose_bytes(q_kn, K, N, q_b200, static_cast<cudaStream_t>(stream));
}

int eetq_b200_unpack(const int8_t* q_b200, int64_t K, int64_t N, int8_t* q_kn, void* stream)
{
    EB_CHECK_ARG(q_kn && q_b200, "unpack: null pointer argument");
    EB_CHECK_ARG(aligned16(q_kn) && aligned16(q_b200), "unpack: pointers must be 16-byte aligned");
    if (int rc = check_kn("unpack", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_transpose_bytes(q_b200, N, K, q_kn, static_cast<cudaStream_t>(stream));
}

int eetq_b200_from_ref_layout(const uint8_t* w_ref, int64_t K, int64_t N, int8_t* q_b200, void* stream)
{
    EB_CHECK_ARG(w_ref && q_b200, "from_ref_layout: null pointer argument");
    EB_CHECK_ARG(aligned16(w_ref) && aligned16(q_b200), "from_ref_layout: pointers must be 16-byte aligned");
    if (int rc = check_kn("from_ref_layout", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_from_ref_layout(w_ref, K, N, q_b200, static_cast<cudaStream_t>(stream));
}

int eetq_b200_to_ref_layout(const int8_t* q_b200, int64_t K, int64_t N, uint8_t* w_ref, void* stream)
{
    EB_CHECK_ARG(w_ref && q_b200, "to_ref_layout: null pointer argument");
    EB_CHECK_ARG(aligned16(w_ref) && aligned16(q_b200), "to_ref_layout: pointers must be 16-byte aligned");
    if (int rc = check_kn("to_ref_layout", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_to_ref_layout(q_b200, K, N, w_ref, static_cast<cudaStream_t>(stream));
}

size_t eetq_b200_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    if (M <= 0 || N <= 0 || K <= 0)
        return 0;
    return gemm_tc_workspace_bytes(M, N, K);
}

namespace {
// EETQ_B200_GEMV_MMA=0 (A/B measurements) keeps 3..8 rows on the SIMT kernel instead of the mma.sync streaming kernel
bool gemv_mma_on()
{
    static const bool v = [] {
        const char* e = getenv("EETQ_B200_GEMV_MMA");
        return !(e != nullptr && e[0] == '0');
    }();
    return v;
}

int gemm_dispatch(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias, const void* residual,
                  int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                  size_t workspace_bytes, int flags, unsigned long long* trace, void* stream)
{
    if (int rc = check_forward_args("w8a16_gemm", x, ldx, w_b200, scales, y, ldy, M, N, K, dtype))
        return rc;
    {
        const int forced = flags & (EETQ_B200_FLAG_FORCE_GEMV | EETQ_B200_FLAG_FORCE_TC | EETQ_B200_FLAG_FORCE_MMA);
        EB_CHECK_ARG((forced & (forced - 1)) == 0, "w8a16_gemm: the FORCE_* flags are exclusive");
    }
    EB_CHECK_ARG(residual == nullptr || (ldr >= N && (ldr % 8) == 0 && aligned16(residual)), "w8a16_gemm: bad residual stride/alignment");
    if (M == 0)
        return EETQ_B200_OK;  // empty batch: nothing to enqueue
    if (int rc = check_arch())
        return rc;

    const bool pdl = (flags & EETQ_B200_FLAG_PDL) != 0;
    bool use_gemv  = M <= 4;  // same cut as the reference's SMALL_M_FAST_PATH (fpA_intB_gemm_wrapper.h:4)
    if (flags & EETQ_B200_FLAG_FORCE_GEMV) {
        EB_CHECK_ARG(M <= EETQ_B200_GEMV_MAX_M, "w8a16_gemm: FORCE_GEMV needs M <= %d", EETQ_B200_GEMV_MAX_M);
        use_gemv = true;
    }
    else if (M >= 3 && K * N >= (int64_t(32) << 20)) {
        // measured (profiles/r02_kbench_*.json): on the large Llama shapes the tcgen05 kernel with a 16-token tile streams the
        // weights faster than the mma.sync kernel and than the reference GEMV for 3 and 4 rows (4096x11008: 14.1 vs 16.5 / 15.9 us;
        // 11008x4096: 15.7 vs 19.9 / 15.1 us); the small 4096x4096 shape stays on the streaming kernels (8.5 vs 10.6 us)
        use_gemv = false;
    }
    if ((flags & EETQ_B200_FLAG_FORCE_TC) || trace != nullptr)
        use_gemv = false;

    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (flags & EETQ_B200_FLAG_FORCE_MMA) {
        EB_CHECK_ARG(trace == nullptr && gemv_mma_supported(int(M < 9 ? M : 9), K, 8), "w8a16_gemm: FORCE_MMA needs M <= 8 (and K that fits)");
        return launch_gemv_mma(x, ldx, w_b200, scales, bias, residual, ldr, y, ldy, int(M), N, K, dtype, 8, pdl, s);
    }
    // 3..8 rows: the mma.sync streaming kernel (M-independent instruction count); 1 and 2 rows: the SIMT kernel measured faster
    // (A/B runs of this round, DESIGN.md section 5: 4096x4096 M = 3 / 4: 8.1-8.6 us vs 8.7 / 8.9 us SIMT)
    if (use_gemv && M >= 3 && gemv_mma_on() && gemv_mma_supported(int(M), K, 8))
        return launch_gemv_mma(x, ldx, w_b200, scales, bias, residual, ldr, y, ldy, int(M), N, K, dtype, 8, pdl, s);
    if (use_gemv) {
        GemvExtras ex;
        ex.residual = residual;
        ex.ldr      = ldr;
        return launch_gemv(x, ldx, w_b200, scales, bias, y, ldy, int(M), N, K, dtype, ex, pdl, s);
    }
    return launch_gemm_tc(x, ldx, w_b200, scales, bias, residual, ldr, y, ldy, M, N, K, dtype, workspace, workspace_bytes, pdl, trace, s);
}
}  // namespace

// The decode GEMV (M <= 8) with its fusions exposed; see eetq_b200_gemv_opts in the header.
int eetq_b200_w8a16_gemv_fused(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias, void* y,
                               int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, const eetq_b200_gemv_opts* o, int pdl,
                               void* stream)
{
    EB_CHECK_ARG(o != nullptr, "w8a16_gemv_fused: null options");
    const bool x_ll = o->x_ll != nullptr && o->x_ll->step != nullptr;
    const bool push = o->push != nullptr && o->push->world > 0;
    // LL buffers are 8-byte words, not element vectors: validate what applies
    void* y_chk     = push ? const_cast<int8_t*>(w_b200) : y;
    const int64_t kx = (o->xmode == GEMV_X_SILU_MUL) ? 2 * K : K;
    const int64_t n_out = (o->epi == GEMV_EPI_SILU_PAIRS) ? N / 2 : N;
    if (int rc = check_forward_args("w8a16_gemv_fused", x, x_ll ? K : ldx, w_b200, scales, y_chk, push ? N : ldy, M, N, K, dtype, n_out))
        return rc;
    EB_CHECK_ARG(M >= 1 && M <= EETQ_B200_GEMV_MAX_M, "w8a16_gemv_fused: M must be in [1, %d]", EETQ_B200_GEMV_MAX_M);
    EB_CHECK_ARG(o->xmode >= 0 && o->xmode <= 2 && (o->xmode != GEMV_X_RMSNORM || (o->norm_weight != nullptr && aligned16(o->norm_weight))),
                 "w8a16_gemv_fused: bad xmode / norm_weight");
    EB_CHECK_ARG(x_ll || ldx >= kx, "w8a16_gemv_fused: ldx too small for the activation mode");
    EB_CHECK_ARG(o->epi == GEMV_EPI_PLAIN || o->epi == GEMV_EPI_SILU_PAIRS, "w8a16_gemv_fused: bad epilogue mode");
    EB_CHECK_ARG(!(x_ll && o->xmode == GEMV_X_SILU_MUL), "w8a16_gemv_fused: LL input cannot be combined with the SiLU-mul prologue");
    if (int rc = check_arch())
        return rc;
    GemvExtras ex;
    ex.norm_weight = o->norm_weight;
    ex.eps         = o->eps;
    ex.xmode       = o->xmode;
    ex.epi         = o->epi;
    ex.residual    = o->residual;
    ex.ldr         = o->ldr;
    if (o->next_w != nullptr) {
        EB_CHECK_ARG(aligned16(o->next_w), "w8a16_gemv_fused: next_w must be 16-byte aligned");
        if (int rc = check_kn("w8a16_gemv_fused (next_w)", o->next_k, o->next_n))
            return rc;
        ex.next_w = o->next_w;
        ex.next_n = o->next_n;
        ex.next_k = o->next_k;
    }
    if (x_ll) {
        ex.x_ll.tag_base = static_cast<const int*>(o->x_ll->step);
        ex.x_ll.per_step = o->x_ll->per_step;
        ex.x_ll.index    = o->x_ll->index;
    }
    if (o->residual_ll != nullptr && o->residual_ll->step != nullptr) {
        EB_CHECK_ARG(o->residual != nullptr, "w8a16_gemv_fused: residual_ll without a residual buffer");
        ex.res_ll.tag_base = static_cast<const int*>(o->residual_ll->step);
        ex.res_ll.per_step = o->residual_ll->per_step;
        ex.res_ll.index    = o->residual_ll->index;
        ex.res_off         = int(o->residual_off);
    }
    if (push) {
        EB_CHECK_ARG(o->push->world <= 8 && o->push->peers != nullptr && o->push->step != nullptr, "w8a16_gemv_fused: bad LL push");
        ex.push.world = o->push->world;
        for (int r = 0; r < o->push->world; ++r)
            ex.push.peer[r] = o->push->peers[r];
        ex.push.local        = static_cast<unsigned long long*>(o->push->local);
        ex.push.elem_off     = int(o->push->elem_off);
        ex.push.tag.tag_base = static_cast<const int*>(o->push->step);
        ex.push.tag.per_step = o->push->per_step;
        ex.push.tag.index    = o->push->index;
    }
    return launch_gemv(x, ldx, w_b200, scales, bias, y, ldy, int(M), N, K, dtype, ex, pdl != 0, static_cast<cudaStream_t>(stream));
}

int eetq_b200_w8a16_gemm_ex(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                            void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace,
                            size_t workspace_bytes, int flags, void* stream)
{
    return gemm_dispatch(x, ldx, w_b200, scales, bias, nullptr, 0, y, ldy, M, N, K, dtype, workspace, workspace_bytes, flags, nullptr,
                         stream);
}

int eetq_b200_w8a16_gemm_residual(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, const void* bias,
                                  const void* residual, int64_t ldr, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype,
                                  void* workspace, size_t workspace_bytes, int flags, void* stream)
{
    return gemm_dispatch(x, ldx, w_b200, scales, bias, residual, ldr, y, ldy, M, N, K, dtype, workspace, workspace_bytes, flags, nullptr,
                         stream);
}

int eetq_b200_w8a16_gemm_trace(const void* x, int64_t ldx, const int8_t* w_b200, const void* scales, void* y, int64_t ldy, int64_t M,
                               int64_t N, int64_t K, void* workspace, size_t workspace_bytes, void* trace, size_t trace_bytes,
                               void* stream)
{
    EB_CHECK_ARG(trace != nullptr, "w8a16_gemm_trace: null trace buffer");
    const size_t need = size_t(gemm_tc_grid_for(M, N, K)) * gemm_tc_trace_slots() * sizeof(unsigned long long);
    EB_CHECK_ARG(trace_bytes >= need, "w8a16_gemm_trace: trace buffer too small (%zu < %zu)", trace_bytes, need);
    return gemm_dispatch(x, ldx, w_b200, scales, nullptr, nullptr, 0, y, ldy, M, N, K, EETQ_B200_F16, workspace, workspace_bytes,
                         EETQ_B200_FLAG_FORCE_TC, static_cast<unsigned long long*>(trace), stream);
}

int eetq_b200_w8a16_gemm_trace_info(int64_t M, int64_t N, int64_t K, int* grid, int* slots)
{
    EB_CHECK_ARG(grid && slots && M > 0 && N > 0 && K > 0, "w8a16_gemm_trace_info: bad argument");
    *grid  = gemm_tc_grid_for(M, N, K);
    *slots = gemm_tc_trace_slots();
    return EETQ_B200_OK;
}

int eetq_b200_w8a16_gemm(const void* x, const int8_t* w_b200, const void* scales, const void* bias, void* y, int64_t M,
                         int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes, void* stream)
{
    return eetq_b200_w8a16_gemm_ex(x, K, w_b200, scales, bias, y, N, M, N, K, dtype, workspace, workspace_bytes,
                                   EETQ_B200_FLAG_DEFAULT, stream);
}

// ---- packed int4 ------------------------------------------------------------------------------------------------------------
int eetq_b200_quantize4(const void* w_kn, int w_dtype, int64_t K, int64_t N, uint8_t* q4_b200, void* scales, float* s32,
                        uint8_t* q4_kn, void* stream)
{
    EB_CHECK_ARG(w_kn && q4_b200 && scales && s32, "quantize4: null pointer argument");
    EB_CHECK_ARG(aligned16(w_kn) && aligned16(q4_b200) && (q4_kn == nullptr || aligned16(q4_kn)),
                 "quantize4: pointers must be 16-byte aligned");
    if (int rc = check_kn("quantize4", K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_quantize4(w_kn, w_dtype, K, N, q4_b200, scales, s32, q4_kn, static_cast<cudaStream_t>(stream));
}

namespace {
int nibble_layout_entry(const char* who, int mode, const uint8_t* src, int64_t K, int64_t N, uint8_t* dst, void* stream)
{
    EB_CHECK_ARG(src && dst, "%s: null pointer argument", who);
    EB_CHECK_ARG(src != dst, "%s: in-place conversion is not supported", who);
    EB_CHECK_ARG(aligned16(src) && aligned16(dst), "%s: pointers must be 16-byte aligned", who);
    if (int rc = check_kn(who, K, N))
        return rc;
    if (int rc = check_arch())
        return rc;
    return launch_nibble_layout(mode, src, K, N, dst, static_cast<cudaStream_t>(stream));
}
}  // namespace

int eetq_b200_pack4(const uint8_t* q4_kn, int64_t K, int64_t N, uint8_t* q4_b200, void* stream)
{
    return nibble_layout_entry("pack4", NIB_PACK4, q4_kn, K, N, q4_b200, stream);
}
int eetq_b200_unpack4(const uint8_t* q4_b200, int64_t K, int64_t N, uint8_t* q4_kn, void* stream)
{
    return nibble_layout_entry("unpack4", NIB_UNPACK4, q4_b200, K, N, q4_kn, stream);
}
int eetq_b200_from_ref_layout4(const uint8_t* w4_ref, int64_t K, int64_t N, uint8_t* q4_b200, void* stream)
{
    return nibble_layout_entry("from_ref_layout4", NIB_FROM_REF4, w4_ref, K, N, q4_b200, stream);
}
int eetq_b200_to_ref_layout4(const uint8_t* q4_b200, int64_t K, int64_t N, uint8_t* w4_ref, void* stream)
{
    return nibble_layout_entry("to_ref_layout4", NIB_TO_REF4, q4_b200, K, N, w4_ref, stream);
}

namespace {
// the tcgen05 scratch comes first (its flag page must stay where the caller zeroed it), the widened weights after it
size_t w4_tc_scratch_bytes(int64_t M, int64_t N, int64_t K) { return (gemm_tc_workspace_bytes(M, N, K) + 255) & ~size_t(255); }
}  // namespace

size_t eetq_b200_w4a16_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    if (M <= EETQ_B200_GEMV4_SIMT_MAX_M || N <= 0 || K <= 0)
        return 0;
    if (M <= EETQ_B200_GEMV4_MAX_M && gemv_mma_on() && gemv_mma_supported(int(M), K, 4))
        return 0;  // streamed by the mma.sync kernel, nothing is widened
    return w4_tc_scratch_bytes(M, N, K) + size_t(N) * size_t(K);
}

int eetq_b200_w4a16_gemm(const void* x, int64_t ldx, const uint8_t* q4_b200, const void* scales, const void* bias, void* y,
                         int64_t ldy, int64_t M, int64_t N, int64_t K, int dtype, void* workspace, size_t workspace_bytes,
                         int flags, void* stream)
{
    if (int rc = check_forward_args("w4a16_gemm", x, ldx, q4_b200, scales, y, ldy, M, N, K, dtype))
        return rc;
    EB_CHECK_ARG(!(flags & EETQ_B200_FLAG_FORCE_TC), "w4a16_gemm: FORCE_TC is not supported");
    EB_CHECK_ARG(!((flags & EETQ_B200_FLAG_FORCE_GEMV) && (flags & EETQ_B200_FLAG_FORCE_MMA)), "w4a16_gemm: the FORCE_* flags are exclusive");
    if (M == 0)
        return EETQ_B200_OK;
    if (int rc = check_arch())
        return rc;
    const bool pdl     = (flags & EETQ_B200_FLAG_PDL) != 0;
    cudaStream_t s     = static_cast<cudaStream_t>(stream);
    const int8_t* w4   = reinterpret_cast<const int8_t*>(q4_b200);
    if (flags & EETQ_B200_FLAG_FORCE_MMA) {
        EB_CHECK_ARG(gemv_mma_supported(int(M < 9 ? M : 9), K, 4), "w4a16_gemm: FORCE_MMA needs M <= 8 and K %% 128 == 0");
        return launch_gemv_mma(x, ldx, w4, scales, bias, nullptr, 0, y, ldy, int(M), N, K, dtype, 4, pdl, s);
    }
    if (flags & EETQ_B200_FLAG_FORCE_GEMV)
        EB_CHECK_ARG(M <= EETQ_B200_GEMV4_SIMT_MAX_M, "w4a16_gemm: FORCE_GEMV needs M <= %d", EETQ_B200_GEMV4_SIMT_MAX_M);
    // measured (profiles/r02_kbench_mma2.json): one row streams fastest through the SIMT kernel (4096x4096: 5.7 vs 6.8 us); from two
    // rows on the mma.sync kernel wins, by up to 1.9x at K = 11008 / 4 rows, and it also beats widening + tcgen05 up to 8 rows
    if (M >= 2 && M <= EETQ_B200_GEMV4_MAX_M && !(flags & EETQ_B200_FLAG_FORCE_GEMV) && gemv_mma_on() && gemv_mma_supported(int(M), K, 4))
        return launch_gemv_mma(x, ldx, w4, scales, bias, nullptr, 0, y, ldy, int(M), N, K, dtype, 4, pdl, s);
    if (M <= EETQ_B200_GEMV4_SIMT_MAX_M) {
        GemvExtras ex;
        ex.wbits = 4;
        return launch_gemv(x, ldx, w4, scales, bias, y, ldy, int(M), N, K, dtype, ex, pdl, s);
    }
    const size_t scratch = w4_tc_scratch_bytes(M, N, K);
    if (workspace == nullptr || !aligned16(workspace) || workspace_bytes < scratch + size_t(N) * size_t(K)) {
        set_error("w4a16_gemm: M = %lld needs a workspace of %zu bytes (got %zu)", (long long)M, scratch + size_t(N) * size_t(K),
                  workspace_bytes);
        return EETQ_B200_EWORKSPACE;
    }
    int8_t* widened = static_cast<int8_t*>(workspace) + scratch;
    if (int rc = launch_widen4to8(q4_b200, K, N, widened, s))
        return rc;
    return launch_gemm_tc(x, ldx, widened, scales, bias, nullptr, 0, y, ldy, M, N, K, dtype, workspace, scratch, false, nullptr, s);
}

// Development aid (trace builds only; a no-op otherwise): point the kernels' timeline recorder at a device buffer of
// 2 + 2 * capacity 64-bit words ([0] = record count, zero it before use).  Returns 1 when the library records, else 0.
int eetq_b200_set_timeline(void* buffer, uint64_t capacity)
{
    trace_set_gemv(buffer, static_cast<unsigned int>(capacity));
    trace_set_decode(buffer, static_cast<unsigned int>(capacity));
#ifdef EETQ_B200_TRACE
    return 1;
#else
    return 0;
#endif
}

int eetq_b200_w8a16_gemm_host(const void* x_host, void* x_dev, const int8_t* w_b200, const void* scales,
                              const void* bias, void* y_dev, void* y_host, int64_t M, int64_t N, int64_t K, int dtype,
                              void* workspace, size_t workspace_bytes, void* stream)
{
    EB_CHECK_ARG(x_host && x_dev && y_dev && y_host, "w8a16_gemm_host: null pointer argument");
    if (int rc = check_forward_args("w8a16_gemm_host", x_dev, K, w_b200, scales, y_dev, N, M, N, K, dtype))
        return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (M == 0)
        return EETQ_B200_OK;
    EB_CHECK_CUDA(cudaMemcpyAsync(x_dev, x_host, size_t(M) * K * 2, cudaMemcpyHostToDevice, s));
    if (int rc = eetq_b200_w8a16_gemm(x_dev, w_b200, scales, bias, y_dev, M, N, K, dtype, workspace, workspace_bytes,
                                      stream))
        return rc;
    EB_CHECK_CUDA(cudaMemcpyAsync(y_host, y_dev, size_t(M) * N * 2, cudaMemcpyDeviceToHost, s));
    return EETQ_B200_OK;
}

}  // extern "C"
