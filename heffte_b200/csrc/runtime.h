// Host-side runtime helpers shared by the translation units of libheffte_b200.so:
// error reporting for the C ABI, the launch counter and the CUDA launcher policy.
#pragma once

#include "cuda_compat.h"

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_set>

#include "../../include/heffte_b200_kernels.h"
#include "fft_device.cuh"

namespace b200 {

void set_error(std::string const &message);
int fail(int code, std::string const &message);
int check_cuda(cudaError_t status, const char *what);
extern std::atomic<long long> launch_counter;

// raises the dynamic shared-memory limit of a kernel once (needed above 48 KB)
void allow_smem(const void *kernel, size_t bytes);

struct cuda_launcher {
    cudaStream_t stream;
    template<typename kernel_t, typename args_t>
    int launch(kernel_t kernel, long long blocks, int threads, size_t smem, args_t const &args){
        if (blocks <= 0) return B200_SUCCESS;
        if (blocks > 2147483647LL) return fail(B200_ERR_UNSUPPORTED, "grid too large");
#ifdef B200_HOST_EMULATION
        emul::launch(kernel, dim3(static_cast<unsigned>(blocks)), dim3(static_cast<unsigned>(threads)), smem, args);   // tests/emul only
#else
        if (smem > 48 * 1024) allow_smem(reinterpret_cast<const void*>(kernel), smem);
        kernel<<<static_cast<unsigned>(blocks), threads, smem, stream>>>(args);
#endif
        launch_counter.fetch_add(1, std::memory_order_relaxed);
        return check_cuda(cudaPeekAtLastError(), "kernel launch");
    }
    // kernel selection of the batched FFT (fft_host_plan.h): the instantiations live in the fft_inst_*.cu slices
    int run_pow2(bool strided, bool is_float, bool scatter, int n, fft_args const &a);
    int run_generic(bool is_float, long long blocks, int threads, size_t smem, generic_args const &g);
    int run_real(bool strided, bool is_float, bool scatter, int kind, int m, fft_args const &a);
    template<typename kernel_t, typename args_t>
    int launch3(kernel_t kernel, long long gx, long long gy, long long gz, int threads, size_t smem, args_t const &args){
        if (gx <= 0 or gy <= 0 or gz <= 0) return B200_SUCCESS;
        if (gx > 2147483647LL or gy > 65535 or gz > 65535) return fail(B200_ERR_UNSUPPORTED, "grid too large");
#ifdef B200_HOST_EMULATION
        emul::launch(kernel, dim3((unsigned) gx, (unsigned) gy, (unsigned) gz), dim3(static_cast<unsigned>(threads)), smem, args);
#else
        kernel<<<dim3((unsigned) gx, (unsigned) gy, (unsigned) gz), threads, smem, stream>>>(args);
#endif
        launch_counter.fetch_add(1, std::memory_order_relaxed);
        return check_cuda(cudaPeekAtLastError(), "kernel launch");
    }
};

} // namespace b200
