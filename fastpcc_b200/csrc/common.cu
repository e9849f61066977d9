#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace fpcc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

int check_epilogue(const fpcc_epilogue *e, bool allow_residual) {
    FPCC_REQUIRE(e != nullptr, "epilogue is NULL");
    FPCC_REQUIRE(e->requant_mul != nullptr && e->zero_point != nullptr, "epilogue: requant_mul / zero_point missing");
    FPCC_REQUIRE(e->shift >= 0 && e->shift < 63, "epilogue: shift %d out of range (must be >= 0)", e->shift);
    FPCC_REQUIRE(e->out_type >= FPCC_OUT_I8 && e->out_type <= FPCC_OUT_I32, "epilogue: bad out_type %d", e->out_type);
    if (e->residual) {
        FPCC_REQUIRE(allow_residual, "epilogue: residual is only supported by the fused kernels");
        FPCC_REQUIRE(e->out_type == FPCC_OUT_I32, "epilogue: residual needs int32 output");
    } else {
        FPCC_REQUIRE(e->post_slope == nullptr, "epilogue: post_slope without residual");
    }
    FPCC_REQUIRE((e->row_bias == nullptr) == (e->row_idx == nullptr), "epilogue: row_bias and row_idx go together");
    FPCC_REQUIRE(e->row_bias == nullptr || allow_residual, "epilogue: row_bias is only supported by the fused kernels");
    if (e->post_requant_mul) {
        FPCC_REQUIRE(allow_residual, "epilogue: the fused second stage is only supported by the fused kernels");
        FPCC_REQUIRE(e->out_type == FPCC_OUT_I32 && e->post_zero_point != nullptr, "epilogue: second stage needs an int32 first stage and a zero point");
        FPCC_REQUIRE(e->post_shift >= 0 && e->post_shift < 63, "epilogue: post_shift %d out of range", e->post_shift);
    } else {
        FPCC_REQUIRE(e->post_requant_slope == nullptr, "epilogue: post_requant_slope without post_requant_mul");
        FPCC_REQUIRE(e->aux_out == nullptr, "epilogue: aux_out without post_requant_mul");
    }
    if (e->out_ld != 0) {
        FPCC_REQUIRE(allow_residual && e->out_ld > 0 && e->out_ld % 16 == 0, "epilogue: out_ld must be a positive multiple of 16 (fused kernels only)");
        FPCC_REQUIRE(e->out_type == FPCC_OUT_I8 || (e->post_requant_mul && !e->aux_out), "epilogue: out_ld applies to int8 output rows");
    }
    return FPCC_OK;
}

}  // namespace fpcc

extern "C" const char *fpcc_last_error(void) { return fpcc::g_err; }
extern "C" int fpcc_version(void) { return 100; }

// Host threads that wait for the device (stream / event synchronisation, blocking copies) sleep instead of spinning.
// One process per GPU with several launching threads each (coding groups): under the default policy every waiting
// thread burns a core (measured: 2.6 cores per rank at 3 groups), which starves the ranks of an 8-GPU job on a box with
// fewer cores than that.  Must be called before the device's context is created to take effect everywhere.
extern "C" int fpcc_set_blocking_sync(int device) {
    FPCC_CUDA(cudaSetDevice(device));
    FPCC_CUDA(cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync));
    return FPCC_OK;
}

extern "C" int fpcc_device_check(int *cc_major, int *cc_minor, int *sm) {
    int dev = 0;
    FPCC_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0, n = 0;
    FPCC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FPCC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    FPCC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (sm) *sm = n;
    FPCC_REQUIRE(major == 10, "fastpcc_b200 is built for sm_100a only; found compute capability %d.%d", major, minor);
    return FPCC_OK;
}
