// C-ABI entry points for the GEMM family and the element-wise requant epilogues.
#include "common.cuh"

namespace fpcc {

int launch_pairs_simt(const PairArgs &a, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s);
int launch_conv_simt(const int8_t *feats, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                     int64_t ld, int n_out, const int32_t *zp_comp, const EpiParams &ep, void *out, cudaStream_t s);
// tensor-core (tcgen05) paths, igemm_tc.cu; return FPCC_ERR_UNSUPPORTED when the shape is not covered
int launch_conv_tc(const int8_t *feats, int n_in, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                   int64_t ld, int n_out, const EpiParams &ep, void *out, cudaStream_t s);
int launch_pairs_tc(const PairArgs &a, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s);
bool tc_enabled();

// ---------------------------------------------------------------------------------------------
// element-wise requant / prelu (src/element_wise/*.cu).  One thread per 4 consecutive channels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) requant_kernel(const int32_t *__restrict__ in, int64_t total, int ch, EpiParams ep,
                                                      void *__restrict__ out) {
    const bool has_slope = ep.slope != nullptr;
    const int32_t slope = has_slope ? ep.slope[0] : 0;
    const int64_t zp = ep.zp[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % ch);
        int64_t o = epi_value(in[i], ep.bias ? ep.bias[c] : 0, has_slope, slope, ep.mul[ep.mul_is_scalar ? 0 : c], zp, ep.shift);
        epi_store(out, i, o, ep.out_type, nullptr, false, 0);
    }
}

__global__ void __launch_bounds__(256) prelu_kernel(const int32_t *__restrict__ in, int64_t total, const int32_t *__restrict__ slope,
                                                    int32_t *__restrict__ out) {
    const int32_t sl = slope[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = clamp_i32(prelu_q25((int64_t)in[i], sl));
}

static int ew_grid(int64_t total) {
    int64_t b = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    return (int)(b < cap ? b : cap);
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_requant(const int32_t *in, int64_t rows, int ch, const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(in && out, "requant: NULL pointer");
    FPCC_REQUIRE(rows > 0 && ch > 0, "requant: rows and channels must be positive");
    int rc = check_epilogue(e, false);
    if (rc) return rc;
    int64_t total = rows * ch;
    requant_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(in, total, ch, to_params(e), out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_prelu_i32(const int32_t *in, int64_t numel, const int32_t *slope, int32_t *out, void *stream) {
    FPCC_REQUIRE(in && out && slope && numel > 0, "prelu: bad arguments");
    prelu_kernel<<<ew_grid(numel), 256, 0, (cudaStream_t)stream>>>(in, numel, slope, out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_gemm_i8(const int8_t *A, const int8_t *B, const int32_t *C, int c_mode, int32_t *D, int M, int N, int K,
                            void *stream) {
    FPCC_REQUIRE(A && B && D, "gemm_i8: NULL pointer");
    FPCC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_i8: bad sizes");
    FPCC_REQUIRE(c_mode >= 0 && c_mode <= 2 && (c_mode == 0 || C), "gemm_i8: bad C operand");
    PairArgs a = {A, B, nullptr, nullptr, nullptr, 1, M, N, K, 1, C, c_mode, 0};
    EpiParams ep = {};
    return launch_pairs_simt(a, ep, D, ceil_div(M, 64), (cudaStream_t)stream);
}

extern "C" int fpcc_gather_gemm_scatter_i8(const int8_t *A, const int8_t *B, int32_t *D, const int32_t *gather_idx,
                                           const int32_t *scatter_idx, int L, int N, int K, void *stream) {
    FPCC_REQUIRE(A && B && D && gather_idx && scatter_idx, "gather_gemm_scatter_i8: NULL pointer");
    FPCC_REQUIRE(L > 0 && N > 0 && K > 0, "gather_gemm_scatter_i8: bad sizes");
    PairArgs a = {A, B, gather_idx, scatter_idx, nullptr, 1, L, N, K, 2, nullptr, 0, 0};
    EpiParams ep = {};
    return launch_pairs_simt(a, ep, D, ceil_div(L, 64), (cudaStream_t)stream);
}

extern "C" int fpcc_spconv_i8(const int8_t *in_feats, int n_in, int c_in, const int8_t *weight, int kvol, int c_out,
                              const int32_t *nbr_table, int64_t ld, int n_out, const int32_t *zp_comp,
                              const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(in_feats && weight && nbr_table && out, "spconv_i8: NULL pointer");
    FPCC_REQUIRE(n_in > 0 && n_out > 0 && c_in > 0 && c_out > 0 && kvol > 0 && ld >= n_out, "spconv_i8: bad sizes");
    int rc = check_epilogue(e, true);
    if (rc) return rc;
    EpiParams ep = to_params(e);
    if (tc_enabled() && !zp_comp) {
        rc = launch_conv_tc(in_feats, n_in, c_in, weight, kvol, c_out, nbr_table, ld, n_out, ep, out, (cudaStream_t)stream);
        if (rc != FPCC_ERR_UNSUPPORTED) return rc;
    }
    return launch_conv_simt(in_feats, c_in, weight, kvol, c_out, nbr_table, ld, n_out, zp_comp, ep, out, (cudaStream_t)stream);
}

extern "C" int fpcc_linear_i8(const int8_t *A, int m, int k, const int8_t *W, int n, const int32_t *sel_row,
                              const int32_t *sel_out, const int32_t *sel_offsets, int n_groups, int n_sel,
                              const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(A && W && out, "linear_i8: NULL pointer");
    FPCC_REQUIRE(m > 0 && n > 0 && k > 0, "linear_i8: bad sizes");
    int rc = check_epilogue(e, true);
    if (rc) return rc;
    EpiParams ep = to_params(e);
    PairArgs a;
    int max_tiles;
    if (sel_row) {
        FPCC_REQUIRE(sel_out && sel_offsets && n_groups > 0 && n_sel > 0, "linear_i8: incomplete selection");
        a = {A, W, sel_row, sel_out, sel_offsets, n_groups, n_sel, n, k, 0, nullptr, 0, 1};
        max_tiles = ceil_div(n_sel, 64) + n_groups;
    } else {
        a = {A, W, nullptr, nullptr, nullptr, 1, m, n, k, 0, nullptr, 0, 0};
        max_tiles = ceil_div(m, 64);
    }
    if (tc_enabled()) {
        rc = launch_pairs_tc(a, ep, out, max_tiles, (cudaStream_t)stream);
        if (rc != FPCC_ERR_UNSUPPORTED) return rc;
    }
    return launch_pairs_simt(a, ep, out, max_tiles, (cudaStream_t)stream);
}
