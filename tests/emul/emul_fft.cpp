// TEST INFRASTRUCTURE ONLY -- runs the product's FFT kernel source (heffte_b200/csrc/fft_device.cuh) on the CPU
// through tests/emul/cuda_emul.h so the CPU-only test-suite can check index arithmetic.  Not part of the product.
#ifndef B200_HOST_EMULATION
#define B200_HOST_EMULATION
#endif
#include "fft_host_plan.h"

namespace {
struct emul_launcher {
    template<typename kernel_t, typename args_t>
    int launch(kernel_t kernel, long long blocks, int threads, size_t smem, args_t const &args){
        emul::launch(kernel, dim3((unsigned) blocks), dim3((unsigned) threads), smem, args);
        return 0;
    }
};
}

extern "C" __attribute__((visibility("default"))) int emul_fft1d(const b200_fft1d_desc *desc, int direction, const void *in, void *out, double scale, int *family){
    b200::host_plan plan;
    const char *why = "";
    int rc = b200::make_host_plan(*desc, plan, &why);
    if (rc) return rc;
    *family = (int) plan.family;
    emul_launcher L;
    if (desc->precision == B200_PREC_FLOAT){
        auto table = b200::make_twiddle_table<float>(plan);
        return b200::run_host_plan(plan, table.data(), direction, in, out, scale, L);
    }
    auto table = b200::make_twiddle_table<double>(plan);
    return b200::run_host_plan(plan, table.data(), direction, in, out, scale, L);
}
