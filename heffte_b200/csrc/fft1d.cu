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
int run_conv_f32_direct(int n, fft_args const &a, cuda_launcher &L);
int run_conv_f32_scatter(int n, fft_args const &a, cuda_launcher &L);
int run_conv_f64_direct(int n, fft_args const &a, cuda_launcher &L);
int run_conv_f64_scatter(int n, fft_args const &a, cuda_launcher &L);
int run_pair_f32_direct(int n, bool contig_first, pair_args const &p, cuda_launcher &L);
int run_pair_f32_scatter(int n, bool contig_first, pair_args const &p, cuda_launcher &L);
int run_pair_f64_direct(int n, bool contig_first, pair_args const &p, cuda_launcher &L);
int run_pair_f64_scatter(int n, bool contig_first, pair_args const &p, cuda_launcher &L);
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

// can the two plans run as one paired launch (fft_pair_kernel)?  One transform along the contiguous axis, one along the middle
// axis of the same box, same power-of-two length with a pair shape, complex data.
static bool pairable(host_plan const &x, host_plan const &y){
    host_plan const &c = (x.family == family_contig) ? x : y, &s = (x.family == family_contig) ? y : x;
    if (c.family != family_contig or s.family != family_strided) return false;
    b200_fft1d_desc const &dc = c.desc, &ds = s.desc;
    if (dc.kind != B200_C2C or ds.kind != B200_C2C or dc.precision != ds.precision) return false;
    if (dc.n != ds.n or not is_pair_length(dc.n)) return false;
    if (dc.in.stride != 1 or dc.out.stride != 1 or ds.in.stride_a != 1 or ds.out.stride_a != 1) return false;
    if (dc.in.stride_a != dc.n or ds.in.stride != dc.n or dc.count_a != ds.n or ds.count_a != dc.n) return false;
    if (dc.count_b != ds.count_b or dc.in.stride_b != ds.in.stride_b or dc.in.stride_b != dc.n * ds.n) return false;
    auto same = [](b200_line_geom const &p, b200_line_geom const &q){ return p.stride == q.stride and p.stride_a == q.stride_a and p.stride_b == q.stride_b; };
    return same(dc.in, dc.out) and same(ds.in, ds.out);
}

int b200_fft1d_pairable(b200_fft1d_plan first, b200_fft1d_plan second){
    return (first != nullptr and second != nullptr and pairable(first->host, second->host)) ? 1 : 0;
}

int b200_fft1d_execute_pair(b200_fft1d_plan first, b200_fft1d_plan second, int direction, const void *in, void *mid,
                            const void *device_scatter_map, double scale, void *counters, int lag, void *stream,
                            int batch, long long in_step, long long mid_step, long long scatter_step, long long local_shift, long long local_step){
    if (first == nullptr or second == nullptr or mid == nullptr or counters == nullptr) return fail(B200_ERR_INVALID, "null argument");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    if (not pairable(first->host, second->host)) return B200_ERR_UNSUPPORTED;
    bool const contig_first = (first->host.family == family_contig);
    bool const backward = (direction == B200_BACKWARD);
    auto fill = [&](b200_fft1d_plan plan, const void *src, void *dst, long long src_step, long long dst_step, double factor){
        b200_fft1d_desc const &d = plan->host.desc;
        fft_args a{};
        a.in = src; a.out = dst; a.twiddle = plan->twiddle; a.twiddle2 = nullptr;
        a.ig = to_geom(d.in); a.og = to_geom(d.out);
        a.nlines = d.count_a * d.count_b;
        a.count_a = static_cast<int>(d.count_a);
        a.backward = backward ? 1 : 0;
        a.scale = factor;
        a.smap = nullptr;
        a.in_step = src_step; a.out_step = dst_step;
        return a;
    };
    fft_args one = fill(first, in, mid, in_step, mid_step, 1.0);
    fft_args two = fill(second, mid, mid, mid_step, mid_step, scale);
    if (device_scatter_map != nullptr){
        two.out = nullptr; two.out_step = 0;
        two.smap = static_cast<const scatter_map*>(device_scatter_map);
        two.scatter_step = scatter_step; two.local_shift = local_shift; two.local_step = local_step;
    }
    pair_args p{};
    p.a = contig_first ? one : two;
    p.b = contig_first ? two : one;
    p.planes = static_cast<unsigned>(first->host.desc.count_b);
    {   // the planes between the two fronts hold about 16 MB: far below the 126 MB of the L2 cache, far above what is in flight
        // (tools/kbench_pair.cu: 4 planes of 512 x 512 complex doubles, 32 planes of 256 x 256 complex floats)
        double const plane_bytes = static_cast<double>(first->host.desc.n) * static_cast<double>(second->host.desc.n) *
                                   ((first->host.desc.precision == B200_PREC_FLOAT) ? 8.0 : 16.0);
        long long automatic = static_cast<long long>(16.0 * 1024 * 1024 / plane_bytes);
        automatic = std::max<long long>(1, std::min<long long>(automatic, 64));
        p.lag = static_cast<unsigned>(lag > 0 ? lag : automatic);
    }
    p.done = static_cast<unsigned*>(counters);
    cudaStream_t const s = static_cast<cudaStream_t>(stream);
    int rc = check_cuda(cudaMemsetAsync(counters, 0, sizeof(unsigned) * static_cast<size_t>(p.planes) * batch, s), "cudaMemsetAsync(pair counters)");
    if (rc) return rc;
    cuda_launcher L{s};
    L.batch = batch;
    int const n = static_cast<int>(first->host.desc.n);
    bool const is_float = (first->host.desc.precision == B200_PREC_FLOAT);
    if (device_scatter_map != nullptr) rc = is_float ? run_pair_f32_scatter(n, contig_first, p, L) : run_pair_f64_scatter(n, contig_first, p, L);
    else rc = is_float ? run_pair_f32_direct(n, contig_first, p, L) : run_pair_f64_direct(n, contig_first, p, L);
    return (rc == -1) ? B200_ERR_UNSUPPORTED : rc;
}

// Two transforms of the same box, neither along its slowest axis, both on the complex fast-path kernels: the first can feed
// the second plane by plane.
static bool overlappable(host_plan const &x, host_plan const &y){
    auto fast = [](host_plan const &h){ return (h.family == family_strided or h.family == family_contig) and h.desc.kind == B200_C2C; };
    if (not fast(x) or not fast(y) or x.desc.precision != y.desc.precision) return false;
    b200_fft1d_desc const &p = x.desc, &q = y.desc;
    if (p.count_b < 2 or p.count_b != q.count_b) return false;                       // planes of the slowest axis
    if (p.out.stride_b != q.in.stride_b or p.in.stride_b != p.out.stride_b or q.in.stride_b != q.out.stride_b) return false;
    if (p.n * p.count_a != q.n * q.count_a) return false;                            // the same plane
    return p.count_a < 2147483647LL and q.count_a < 2147483647LL;
}
int b200_fft1d_overlappable(b200_fft1d_plan first, b200_fft1d_plan second){
    return (first != nullptr and second != nullptr and overlappable(first->host, second->host)) ? 1 : 0;
}

// first: in -> mid on `side_stream`, reporting plane by plane; second: mid -> scatter map on `stream` with a thin grid, waiting
// plane by plane.  stream: zero the counters, fork; side_stream: first transform, join; stream: second transform, wait for the join.
int b200_fft1d_execute_overlapped(b200_fft1d_plan first, b200_fft1d_plan second, int direction, const void *in, void *mid,
                                  const void *device_scatter_map, int map_nb, double scale, void *counters, void *stream, void *side_stream,
                                  void *fork_event, void *join_event,
                                  int batch, long long in_step, long long mid_step, long long scatter_step, long long local_shift, long long local_step,
                                  int thin_blocks){
    if (first == nullptr or second == nullptr or mid == nullptr or counters == nullptr or device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "null argument");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    if (not overlappable(first->host, second->host)) return B200_ERR_UNSUPPORTED;
    long long const planes = first->host.desc.count_b;
    cudaStream_t const main_stream = static_cast<cudaStream_t>(stream), side = static_cast<cudaStream_t>(side_stream);
    int rc = check_cuda(cudaMemsetAsync(counters, 0, sizeof(unsigned) * static_cast<size_t>(planes) * batch, main_stream), "cudaMemsetAsync(plane counters)");
    if (rc == 0) rc = check_cuda(cudaEventRecord(static_cast<cudaEvent_t>(fork_event), main_stream), "cudaEventRecord(fork)");
    if (rc == 0) rc = check_cuda(cudaStreamWaitEvent(side, static_cast<cudaEvent_t>(fork_event), 0), "cudaStreamWaitEvent(fork)");
    if (rc) return rc;
    {
        cuda_launcher L{side};
        batch_steps steps; steps.batch = batch; steps.in_step = in_step; steps.out_step = mid_step;
        steps.done = static_cast<unsigned*>(counters); steps.done_mode = 1; steps.order_nb = map_nb;
        rc = run_host_plan(first->host, first->twiddle, direction, in, mid, 1.0, L, nullptr, 0, -1, steps);
        if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
        if (rc == 0) rc = check_cuda(cudaEventRecord(static_cast<cudaEvent_t>(join_event), side), "cudaEventRecord(join)");
        if (rc) return rc;
    }
    {
        cuda_launcher L{main_stream};
        batch_steps steps; steps.batch = batch; steps.in_step = mid_step;
        steps.scatter_step = scatter_step; steps.local_shift = local_shift; steps.local_step = local_step;
        steps.done = static_cast<unsigned*>(counters); steps.done_mode = 2; steps.done_need = static_cast<unsigned>(first->host.desc.count_a);
        steps.max_blocks = (thin_blocks > 0) ? thin_blocks : 0;
        rc = run_host_plan(second->host, second->twiddle, direction, mid, nullptr, scale, L, device_scatter_map, 0, -1, steps);
        if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
        if (rc == 0) rc = check_cuda(cudaStreamWaitEvent(main_stream, static_cast<cudaEvent_t>(join_event), 0), "cudaStreamWaitEvent(join)");
    }
    return rc;
}

int b200_fft1d_convolvable(b200_fft1d_plan plan){
    if (plan == nullptr) return 0;
    b200_fft1d_desc const &d = plan->host.desc;
    auto same = [](b200_line_geom const &p, b200_line_geom const &q){ return p.stride == q.stride and p.stride_a == q.stride_a and p.stride_b == q.stride_b; };
    return (plan->host.family == family_strided and d.kind == B200_C2C and is_conv_length(d.n) and same(d.in, d.out)) ? 1 : 0;
}

int b200_fft1d_execute_convolve(b200_fft1d_plan plan, const void *in, void *out, const void *device_scatter_map, const void *multiplier,
                                double scale, void *stream, int batch, long long in_step, long long out_step,
                                long long scatter_step, long long local_shift, long long local_step){
    if (plan == nullptr) return fail(B200_ERR_INVALID, "null plan");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    if (not b200_fft1d_convolvable(plan)) return B200_ERR_UNSUPPORTED;
    if (out == nullptr and device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "no destination");
    b200_fft1d_desc const &d = plan->host.desc;
    fft_args a{};
    a.in = in; a.out = out; a.twiddle = plan->twiddle; a.twiddle2 = nullptr;
    a.ig = to_geom(d.in); a.og = to_geom(d.out);
    a.nlines = d.count_a * d.count_b;
    a.count_a = static_cast<int>(d.count_a);
    a.backward = 0;
    a.scale = scale;
    a.smap = static_cast<const scatter_map*>(device_scatter_map);
    a.in_step = in_step; a.out_step = out_step; a.scatter_step = scatter_step; a.local_shift = local_shift; a.local_step = local_step;
    a.multiplier = multiplier;
    if (a.nlines == 0) return B200_SUCCESS;
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    L.batch = batch;
    bool const is_float = (d.precision == B200_PREC_FLOAT);
    int rc;
    if (device_scatter_map != nullptr) rc = is_float ? run_conv_f32_scatter(static_cast<int>(d.n), a, L) : run_conv_f64_scatter(static_cast<int>(d.n), a, L);
    else rc = is_float ? run_conv_f32_direct(static_cast<int>(d.n), a, L) : run_conv_f64_direct(static_cast<int>(d.n), a, L);
    return (rc == -1) ? B200_ERR_UNSUPPORTED : rc;
}

int b200_fft1d_execute_scatter(b200_fft1d_plan plan, int direction, const void *in, const void *device_scatter_map, double scale, void *stream){
    return b200_fft1d_execute_scatter_batch(plan, direction, in, device_scatter_map, scale, stream, 1, 0, 0, 0, 0);
}

int b200_fft1d_execute_batch(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream,
                             int batch, long long in_step, long long out_step){
    if (plan == nullptr) return fail(B200_ERR_INVALID, "null plan");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    batch_steps steps; steps.batch = batch; steps.in_step = in_step; steps.out_step = out_step;
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, out, scale, L, nullptr, 0, -1, steps);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

int b200_fft1d_execute_scatter_batch(b200_fft1d_plan plan, int direction, const void *in, const void *device_scatter_map, double scale, void *stream,
                                     int batch, long long in_step, long long scatter_step, long long local_shift, long long local_step){
    if (plan == nullptr or device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "null plan or scatter map");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    batch_steps steps; steps.batch = batch; steps.in_step = in_step; steps.scatter_step = scatter_step; steps.local_shift = local_shift; steps.local_step = local_step;
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, nullptr, scale, L, device_scatter_map, 0, -1, steps);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

} // extern "C"
