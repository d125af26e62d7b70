// Host-side runtime helpers shared by the translation units of libheffte_b200.so:
// error reporting for the C ABI, the launch counter and the CUDA launcher policy.
#pragma once

#include "cuda_compat.h"

#include <algorithm>
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

#ifndef B200_HOST_EMULATION
// The tensor map of a strided stage for the TMA-loaded kernel: (line axis, transform axis, slower line axis, batch entry) over
// elements of `real_bytes` (4 or 8), a box of `lpb` lines x 256 rows.  False when the driver offers no encoder or the box of
// the stage does not fit the rules of a tensor map: the caller then keeps the cp.async kernel.
bool encode_tile_map(tma_tile_map &map, const void *base, int real_bytes, long long count_a, long long n, long long stride, long long count_b, long long stride_b,
                     int batch, long long step_bytes, int lpb);
#endif
// raises the dynamic shared-memory limit of a kernel once (needed above 48 KB)
void allow_smem(const void *kernel, size_t bytes);

struct cuda_launcher {
    cudaStream_t stream;
    int batch = 1;          // grid.y: entries of a batched transform (the kernels read blockIdx.y)
    long long max_blocks = 0;   // > 0: cap of grid.x (set only for kernels that walk their tiles grid-stride)
    template<typename kernel_t, typename args_t>
    int launch(kernel_t kernel, long long blocks, int threads, size_t smem, args_t const &args){
        if (blocks <= 0 or batch <= 0) return B200_SUCCESS;
        if (max_blocks > 0 and blocks > max_blocks) blocks = max_blocks;
        if (blocks > 2147483647LL or batch > 65535) return fail(B200_ERR_UNSUPPORTED, "grid too large");
#ifdef B200_HOST_EMULATION
        emul::launch(kernel, dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(batch)), dim3(static_cast<unsigned>(threads)), smem, args);   // tests/emul only
#else
        if (smem > 48 * 1024) allow_smem(reinterpret_cast<const void*>(kernel), smem);
        kernel<<<dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(batch)), threads, smem, stream>>>(args);
#endif
        launch_counter.fetch_add(1, std::memory_order_relaxed);
        return check_cuda(cudaPeekAtLastError(), "kernel launch");
    }
#ifndef B200_HOST_EMULATION
    // a kernel whose tile arrives by TMA: the tensor map is its second parameter
    template<typename kernel_t>
    int launch_tma(kernel_t kernel, long long blocks, int threads, size_t smem, fft_args const &args, tma_tile_map const &map){
        if (blocks <= 0 or batch <= 0) return B200_SUCCESS;
        if (max_blocks > 0 and blocks > max_blocks) blocks = max_blocks;
        if (blocks > 2147483647LL or batch > 65535) return fail(B200_ERR_UNSUPPORTED, "grid too large");
        if (smem > 48 * 1024) allow_smem(reinterpret_cast<const void*>(kernel), smem);
        kernel<<<dim3(static_cast<unsigned>(blocks), static_cast<unsigned>(batch)), threads, smem, stream>>>(args, map);
        launch_counter.fetch_add(1, std::memory_order_relaxed);
        return check_cuda(cudaPeekAtLastError(), "kernel launch");
    }
#endif
    // persistent launch of the paired kernel: the whole grid must be resident at once (its CTAs wait for one another), so it is
    // sized by the occupancy of the kernel; grid.y = batch entries
    template<typename kernel_t>
    int launch_pair(kernel_t kernel, int threads, size_t smem, pair_args const &args){
        if (batch <= 0) return B200_SUCCESS;
#ifdef B200_HOST_EMULATION
        emul::launch(kernel, dim3(1u, static_cast<unsigned>(batch)), dim3(static_cast<unsigned>(threads)), smem, args);   // blocks run one after the other
#else
        if (smem > 48 * 1024) allow_smem(reinterpret_cast<const void*>(kernel), smem);
        int per_sm = 0, sms = 0, device = 0;
        if (cudaGetDevice(&device) != cudaSuccess or cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess or
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess or per_sm < 1)
            return check_cuda(cudaGetLastError(), "occupancy query of the paired kernel");
        int const grid = std::max(1, (per_sm * sms) / batch);      // all entries share the GPU
        kernel<<<dim3(static_cast<unsigned>(grid), static_cast<unsigned>(batch)), threads, smem, stream>>>(args);
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
