// Batched strided 1-D FFT plans: CUDA side of the executors (device twiddle table, launches on a stream).
// Replaces heffte::plan_cufft / plan_cufft_r2c and the cufftExec* calls
// (reference: include/heffte_backend_cuda.h:346-422, 494-524, 580-621, 694-727).
#include "fft_host_plan.h"
#include "runtime.h"

#include <cstring>

namespace b200 {

thread_local std::string last_error_text;
std::atomic<long long> launch_counter{0};

void set_error(std::string const &message){ last_error_text = message; }
int fail(int code, std::string const &message){ set_error(message); return code; }
int check_cuda(cudaError_t status, const char *what){
    if (status == cudaSuccess) return B200_SUCCESS;
    set_error(std::string(what) + ": " + cudaGetErrorString(status));
    return B200_ERR_CUDA;
}
void allow_smem(const void *kernel, size_t){
    static std::mutex guard;
    static std::unordered_set<const void*> done;
    std::lock_guard<std::mutex> lock(guard);
    if (done.count(kernel)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    done.insert(kernel);
}


#define B200_DECLARE_SLICE(name) int name(int n, fft_args const &a, cuda_launcher &L)
B200_DECLARE_SLICE(run_strided_f32_direct); B200_DECLARE_SLICE(run_strided_f32_scatter);
B200_DECLARE_SLICE(run_strided_f64_direct); B200_DECLARE_SLICE(run_strided_f64_scatter);
B200_DECLARE_SLICE(run_contig_f32_direct);  B200_DECLARE_SLICE(run_contig_f32_scatter);
B200_DECLARE_SLICE(run_contig_f64_direct);  B200_DECLARE_SLICE(run_contig_f64_scatter);

int cuda_launcher::run_pow2(bool strided, bool is_float, bool scatter, int n, fft_args const &a){
    if (strided){
        if (is_float) return scatter ? run_strided_f32_scatter(n, a, *this) : run_strided_f32_direct(n, a, *this);
        return scatter ? run_strided_f64_scatter(n, a, *this) : run_strided_f64_direct(n, a, *this);
    }
    if (is_float) return scatter ? run_contig_f32_scatter(n, a, *this) : run_contig_f32_direct(n, a, *this);
    return scatter ? run_contig_f64_scatter(n, a, *this) : run_contig_f64_direct(n, a, *this);
}
int run_real_f32_direct(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_real_f32_scatter(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_real_f64_direct(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_real_f64_scatter(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_sreal_f32_direct(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_sreal_f32_scatter(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_sreal_f64_direct(int kind, int m, fft_args const &a, cuda_launcher &L);
int run_sreal_f64_scatter(int kind, int m, fft_args const &a, cuda_launcher &L);
int cuda_launcher::run_real(bool strided, bool is_float, bool scatter, int kind, int m, fft_args const &a){
    if (strided){
        if (is_float) return scatter ? run_sreal_f32_scatter(kind, m, a, *this) : run_sreal_f32_direct(kind, m, a, *this);
        return scatter ? run_sreal_f64_scatter(kind, m, a, *this) : run_sreal_f64_direct(kind, m, a, *this);
    }
    if (is_float) return scatter ? run_real_f32_scatter(kind, m, a, *this) : run_real_f32_direct(kind, m, a, *this);
    return scatter ? run_real_f64_scatter(kind, m, a, *this) : run_real_f64_direct(kind, m, a, *this);
}
int cuda_launcher::run_generic(bool is_float, long long blocks, int threads, size_t smem, generic_args const &g){
    if (is_float) return launch(fft_generic_kernel<float>, blocks, threads, smem, g);
    return launch(fft_generic_kernel<double>, blocks, threads, smem, g);
}

} // namespace b200

using namespace b200;

struct b200_fft1d_plan_s {
    host_plan host;
    void *twiddle = nullptr;   // device table
};

extern "C" {

const char* b200_last_error(void){ return last_error_text.c_str(); }
long long b200_launch_count(void){ return launch_counter.load(); }
int b200_device_count(void){
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}

int b200_device_alloc(size_t bytes, void **device_pointer){
    if (device_pointer == nullptr) return fail(B200_ERR_INVALID, "null argument");
    *device_pointer = nullptr;
    if (bytes == 0) return B200_SUCCESS;
    return check_cuda(cudaMalloc(device_pointer, bytes), "cudaMalloc");
}
int b200_device_free(void *device_pointer){ return (device_pointer == nullptr) ? B200_SUCCESS : check_cuda(cudaFree(device_pointer), "cudaFree"); }
int b200_copy_to_device(const void *host, void *device, size_t bytes, void *stream){
    if (bytes == 0) return B200_SUCCESS;
    return check_cuda(cudaMemcpyAsync(device, host, bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)), "copy to device");
}
int b200_copy_to_host(const void *device, void *host, size_t bytes, void *stream){
    if (bytes == 0) return B200_SUCCESS;
    return check_cuda(cudaMemcpyAsync(host, device, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)), "copy to host");
}
int b200_copy_on_device(const void *source, void *destination, size_t bytes, void *stream){
    if (bytes == 0) return B200_SUCCESS;
    return check_cuda(cudaMemcpyAsync(destination, source, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)), "copy on device");
}
int b200_stream_synchronize(void *stream){ return check_cuda(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)), "stream synchronize"); }
int b200_stream_create(void **stream){
    if (stream == nullptr) return fail(B200_ERR_INVALID, "null argument");
    cudaStream_t s = nullptr;
    int rc = check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreateWithFlags");
    *stream = s;
    return rc;
}
int b200_stream_destroy(void *stream){ return (stream == nullptr) ? B200_SUCCESS : check_cuda(cudaStreamDestroy(static_cast<cudaStream_t>(stream)), "cudaStreamDestroy"); }
int b200_device_set(int device){ return check_cuda(cudaSetDevice(device), "cudaSetDevice"); }

int b200_fft1d_create(const b200_fft1d_desc *desc, b200_fft1d_plan *out){
    if (desc == nullptr or out == nullptr) return fail(B200_ERR_INVALID, "null argument");
    auto *plan = new b200_fft1d_plan_s();
    const char *why = "";
    int rc = make_host_plan(*desc, plan->host, &why);
    if (rc){ delete plan; return fail(rc, why); }
    if (b200_device_count() < 1){ delete plan; return fail(B200_ERR_NO_DEVICE, "no CUDA device: the b200 backend has no CPU fallback"); }
    size_t bytes = 0;
    if (desc->precision == B200_PREC_FLOAT){
        auto table = make_twiddle_table<float>(plan->host);
        bytes = table.size() * sizeof(float);
        rc = check_cuda(cudaMalloc(&plan->twiddle, bytes), "cudaMalloc(twiddle)");
        if (!rc) rc = check_cuda(cudaMemcpy(plan->twiddle, table.data(), bytes, cudaMemcpyHostToDevice), "cudaMemcpy(twiddle)");
    }else{
        auto table = make_twiddle_table<double>(plan->host);
        bytes = table.size() * sizeof(double);
        rc = check_cuda(cudaMalloc(&plan->twiddle, bytes), "cudaMalloc(twiddle)");
        if (!rc) rc = check_cuda(cudaMemcpy(plan->twiddle, table.data(), bytes, cudaMemcpyHostToDevice), "cudaMemcpy(twiddle)");
    }
    if (rc){ if (plan->twiddle) cudaFree(plan->twiddle); delete plan; return rc; }
    *out = plan;
    return B200_SUCCESS;
}

int b200_fft1d_destroy(b200_fft1d_plan plan){
    if (plan == nullptr) return B200_SUCCESS;
    if (plan->twiddle) cudaFree(plan->twiddle);
    delete plan;
    return B200_SUCCESS;
}

const char* b200_fft1d_kernel_name(b200_fft1d_plan plan){
    if (plan == nullptr) return "null";
    switch(plan->host.family){
        case family_strided: return "strided";
        case family_contig: return "contig";
        case family_contig_real: return "contig_real";
        case family_strided_real: return "strided_real";
        default: return "generic";
    }
}

int b200_fft1d_execute(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream){
    if (plan == nullptr) return fail(B200_ERR_INVALID, "null plan");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, out, scale, L);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

int b200_fft1d_execute_range(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream, long long b_begin, long long b_count){
    if (plan == nullptr) return fail(B200_ERR_INVALID, "null plan");
    if (b_count < 0) return fail(B200_ERR_INVALID, "negative line count");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, out, scale, L, nullptr, b_begin, b_count);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    if (rc == B200_ERR_INVALID) return fail(rc, "line range outside the plan");
    return rc;
}

int b200_fft1d_execute_scatter(b200_fft1d_plan plan, int direction, const void *in, const void *device_scatter_map, double scale, void *stream){
    if (plan == nullptr or device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "null plan or scatter map");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, nullptr, scale, L, device_scatter_map);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

} // extern "C"
