// Batched strided 1-D FFT plans: CUDA side of the executors (device twiddle table, launches on a stream).
// Replaces heffte::plan_cufft / plan_cufft_r2c and the cufftExec* calls
// (reference: include/heffte_backend_cuda.h:346-422, 494-524, 580-621, 694-727).
#include "fft_host_plan.h"
#include "composite.cuh"
#include "runtime.h"
#ifndef B200_HOST_EMULATION
#include <cuda.h>      // the types of the tensor-map encoder; the function itself comes from cudaGetDriverEntryPoint (no libcuda at link time)
#endif

#include <complex>
#include <memory>

#include <cstdlib>
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

#ifndef B200_HOST_EMULATION
bool encode_tile_map(tma_tile_map &map, const void *base, int real_bytes, long long count_a, long long n, long long stride, long long count_b, long long stride_b,
                     int batch, long long step_bytes, int lpb){
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn const encode = []() -> encode_fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult status;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &status) != cudaSuccess or status != cudaDriverEntryPointSuccess){
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<encode_fn>(f);
    }();
    static_assert(sizeof(tma_tile_map) == sizeof(CUtensorMap) and alignof(tma_tile_map) >= alignof(CUtensorMap), "tma_tile_map stands for a CUtensorMap");
    if (encode == nullptr or count_a <= 0 or n <= 0 or count_b <= 0 or batch <= 0) return false;
    long long const cb = 2LL * real_bytes;                       // a complex element
    // (2 count_a reals | n rows | count_b | batch); the strides of axes of extent one only have to be legal
    cuuint64_t const s1 = static_cast<cuuint64_t>(stride * cb);
    cuuint64_t const s2 = (count_b > 1) ? static_cast<cuuint64_t>(stride_b * cb) : s1 * static_cast<cuuint64_t>(n);
    cuuint64_t const s3 = (batch > 1) ? static_cast<cuuint64_t>(step_bytes) : s2 * static_cast<cuuint64_t>(count_b);
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(2 * count_a), static_cast<cuuint64_t>(n), static_cast<cuuint64_t>(count_b), static_cast<cuuint64_t>(batch)};
    cuuint64_t gstride[3] = {s1, s2, s3};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(2 * lpb), static_cast<cuuint32_t>(std::min<long long>(n, 256)), 1, 1};
    cuuint32_t estride[4] = {1, 1, 1, 1};
    for(cuuint64_t v : gstride) if (v == 0 or v % 16 != 0 or v >= (1ULL << 40)) return false;
    for(cuuint64_t v : gdim) if (v == 0 or v > 0xffffffffULL) return false;
    if (reinterpret_cast<uintptr_t>(base) % 16 != 0) return false;
    CUresult const rc = encode(reinterpret_cast<CUtensorMap*>(&map), (real_bytes == 4) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4,
                               const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS;
}
#endif


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
int run_sreal2_f32(int kind, int n, fft_args const &a, cuda_launcher &L);
int run_sreal2_f64(int kind, int n, fft_args const &a, cuda_launcher &L);
int run_creal2_f32(int kind, int n, fft_args const &a, cuda_launcher &L);
int run_creal2_f64(int kind, int n, fft_args const &a, cuda_launcher &L);
int cuda_launcher::run_real(bool strided, bool is_float, bool scatter, int kind, int m, fft_args const &a){
    if (not strided and not scatter and contig_real2_applies(kind, m, a)){
        int const rc = is_float ? run_creal2_f32(kind, 2 * m, a, *this) : run_creal2_f64(kind, 2 * m, a, *this);
        if (rc != -1) return rc;
    }
    if (strided and not scatter and real2_applies(is_float, kind, m, a)){
        int const rc = is_float ? run_sreal2_f32(kind, 2 * m, a, *this) : run_sreal2_f64(kind, 2 * m, a, *this);
        if (rc != -1) return rc;
    }
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

// the composite engine (composite.cuh): any length through sub-plans over a workspace
struct composite_plan {
    long long engine = 0;        // length of the complex sequence built from a line: n, or 2 (n - 1) for the type-I cosine transform
    long long m_fft = 0, n1 = 0, n2 = 1;
    long long L = 0;             // lines per chunk
    bool bluestein = false;
    b200_fft1d_plan sub_a = nullptr, sub_c = nullptr;
    void *work = nullptr, *twiddle = nullptr, *chirp = nullptr, *bhat = nullptr, *w4n = nullptr;
    ~composite_plan(){
        if (sub_a) b200_fft1d_destroy(sub_a);
        if (sub_c) b200_fft1d_destroy(sub_c);
        for(void *p : {work, twiddle, chirp, bhat, w4n}) if (p) cudaFree(p);
    }
};

struct b200_fft1d_plan_s {
    host_plan host;
    void *twiddle = nullptr;   // device table
    std::unique_ptr<composite_plan> composite;
};

namespace {

int largest_prime_factor(long long m){
    long long best = 1;
    for(long long p = 2; p * p <= m; p++) while(m % p == 0){ best = p; m /= p; }
    return static_cast<int>(std::max(best, m));
}
// lengths a sub-plan serves well: the register / shared-memory kernels, or the generic kernel with small prime factors
bool good_sub_length(long long n){ return n >= 2 and n <= 4096 and (is_fast_length(n) or largest_prime_factor(n) <= 61); }

// m = n1 * n2 with both factors good sub-plan lengths, as balanced as possible; false when there is no such split
bool split_length(long long m, long long &n1, long long &n2){
    long long root = 1;
    while((root + 1) * (root + 1) <= m) root++;
    for(long long d = root; d >= 2; d--){
        if (m % d != 0) continue;
        if (good_sub_length(d) and good_sub_length(m / d)){ n1 = m / d; n2 = d; return true; }
    }
    return false;
}

template<typename T>
int upload(std::vector<std::complex<long double>> const &host, void **device){
    std::vector<T> flat(2 * host.size());
    for(size_t i=0; i<host.size(); i++){ flat[2*i] = static_cast<T>(host[i].real()); flat[2*i+1] = static_cast<T>(host[i].imag()); }
    int rc = check_cuda(cudaMalloc(device, flat.size() * sizeof(T)), "cudaMalloc(composite table)");
    if (rc == 0) rc = check_cuda(cudaMemcpy(*device, flat.data(), flat.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy(composite table)");
    return rc;
}

// in-place radix-2 transform of a power-of-two sequence on the host (plan time only: the chirp of Bluestein's algorithm)
void host_fft(std::vector<std::complex<long double>> &x){
    size_t const n = x.size();
    for(size_t i=1, j=0; i<n; i++){
        size_t bit = n >> 1;
        for(; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(x[i], x[j]);
    }
    const long double pi = 3.141592653589793238462643383279502884L;
    for(size_t len = 2; len <= n; len <<= 1){
        std::vector<std::complex<long double>> w(len / 2);
        for(size_t k=0; k<len/2; k++) w[k] = std::complex<long double>(cosl(2 * pi * k / len), -sinl(2 * pi * k / len));
        for(size_t i=0; i<n; i+=len)
            for(size_t k=0; k<len/2; k++){
                std::complex<long double> const u = x[i+k], v = x[i+k+len/2] * w[k];
                x[i+k] = u + v; x[i+k+len/2] = u - v;
            }
    }
}

template<typename T>
int build_composite(b200_fft1d_plan_s &plan, long long engine){
    std::unique_ptr<composite_plan> c(new composite_plan());
    b200_fft1d_desc const &d = plan.host.desc;
    c->engine = engine;
    const long double pi = 3.141592653589793238462643383279502884L;
    if (good_sub_length(engine)){ c->m_fft = engine; c->n1 = engine; c->n2 = 1; }
    else if (split_length(engine, c->n1, c->n2)) c->m_fft = engine;
    else{
        c->bluestein = true;
        c->m_fft = 16;
        while(c->m_fft < 2 * engine - 1) c->m_fft *= 2;
        if (c->m_fft <= 4096){ c->n1 = c->m_fft; c->n2 = 1; }
        else{
            c->n1 = 1;
            while(c->n1 * c->n1 < c->m_fft) c->n1 *= 2;
            c->n2 = c->m_fft / c->n1;
            if (c->n1 > 4096 or c->n2 > 4096) return fail(B200_ERR_UNSUPPORTED, "transform length beyond the composite engine (about 8 million points)");
        }
    }
    size_t const csize = sizeof(T) * 2;
    long long const nlines = std::max<long long>(1, d.count_a * d.count_b);
    // a chunk of lines whose workspace stays below 256 MB
    long long L = std::max<long long>(1, (256LL << 20) / static_cast<long long>(csize * c->m_fft));
    L = std::min<long long>(L, nlines);
    if (L > 8) L -= L % 8;
    c->L = L;
    int rc = check_cuda(cudaMalloc(&c->work, csize * static_cast<size_t>(c->m_fft) * static_cast<size_t>(L)), "cudaMalloc(composite workspace)");
    if (rc) return rc;
    // sub-plans over the workspace [position][line]
    {
        b200_fft1d_desc a{};
        a.precision = d.precision; a.kind = B200_C2C;
        a.n = c->n1;
        a.in = b200_line_geom{c->n2 * L, 1, 0}; a.out = a.in;
        a.count_a = c->n2 * L; a.count_b = 1;
        rc = b200_fft1d_create(&a, &c->sub_a);
        if (rc) return rc;
        if (c->n2 > 1){
            b200_fft1d_desc s{};
            s.precision = d.precision; s.kind = B200_C2C;
            s.n = c->n2;
            s.in = b200_line_geom{L, 1, c->n2 * L}; s.out = s.in;
            s.count_a = L; s.count_b = c->n1;
            rc = b200_fft1d_create(&s, &c->sub_c);
            if (rc) return rc;
        }
    }
    if (c->n2 > 1){
        std::vector<std::complex<long double>> tw(static_cast<size_t>(c->m_fft));
        for(long long t=0; t<c->m_fft; t++){
            long double const angle = 2 * pi * static_cast<long double>(t) / static_cast<long double>(c->m_fft);
            tw[t] = std::complex<long double>(cosl(angle), -sinl(angle));
        }
        rc = upload<T>(tw, &c->twiddle);
        if (rc) return rc;
    }
    if (c->bluestein){
        // w_j = exp(-i pi j^2 / E), the angle reduced through j^2 mod 2E
        std::vector<std::complex<long double>> w(static_cast<size_t>(engine));
        for(long long j=0; j<engine; j++){
            long long const r = static_cast<long long>((static_cast<unsigned __int128>(j) * static_cast<unsigned __int128>(j)) % static_cast<unsigned __int128>(2 * engine));
            long double const angle = pi * static_cast<long double>(r) / static_cast<long double>(engine);
            w[j] = std::complex<long double>(cosl(angle), -sinl(angle));
        }
        rc = upload<T>(w, &c->chirp);
        if (rc) return rc;
        std::vector<std::complex<long double>> b(static_cast<size_t>(c->m_fft), std::complex<long double>(0, 0));
        b[0] = std::conj(w[0]);
        for(long long j=1; j<engine; j++){ b[j] = std::conj(w[j]); b[c->m_fft - j] = std::conj(w[j]); }
        host_fft(b);
        std::vector<std::complex<long double>> placed(b.size());
        for(long long k=0; k<c->m_fft; k++){
            long long const pos = (c->n2 == 1) ? k : (k % c->n1) * c->n2 + k / c->n1;
            placed[pos] = b[k] / static_cast<long double>(c->m_fft);
        }
        rc = upload<T>(placed, &c->bhat);
        if (rc) return rc;
    }
    if (d.kind == B200_COS or d.kind == B200_SIN){
        std::vector<std::complex<long double>> q(static_cast<size_t>(d.n + 1));
        for(long long k=0; k<=d.n; k++){
            long double const angle = 2 * pi * static_cast<long double>(k) / static_cast<long double>(4 * d.n);
            q[k] = std::complex<long double>(cosl(angle), -sinl(angle));
        }
        rc = upload<T>(q, &c->w4n);
        if (rc) return rc;
    }
    plan.composite = std::move(c);
    return B200_SUCCESS;
}

// one chunk after the other: gather, transform the workspace, scatter back
int run_composite(b200_fft1d_plan_s const &plan, int direction, const void *in, void *out, double scale, cudaStream_t stream,
                  const void *scatter, batch_shift shift){
    composite_plan const &c = *plan.composite;
    b200_fft1d_desc const &d = plan.host.desc;
    bool const backward = (direction == B200_BACKWARD);
    bool const is_float = (d.precision == B200_PREC_FLOAT);
    composite_args a{};
    generic_args &g = a.g;
    g.in = in; g.out = out; g.twiddle = nullptr;
    g.ig = to_geom(backward ? d.out : d.in);
    g.og = to_geom(backward ? d.in : d.out);
    g.nlines = d.count_a * d.count_b;
    g.count_a = static_cast<int>(d.count_a);
    g.backward = backward ? 1 : 0;
    g.scale = scale;
    g.n = static_cast<int>(d.n);
    g.m = static_cast<int>(c.engine);
    g.smap = static_cast<const scatter_map*>(scatter);
    switch(d.kind){
        case B200_C2C:  g.mode = mode_c2c; break;
        case B200_R2C:  g.mode = backward ? mode_c2r : mode_r2c; break;
        case B200_COS:  g.mode = backward ? mode_dct3 : mode_dct2; break;
        case B200_SIN:  g.mode = backward ? mode_dst3 : mode_dst2; break;
        default:        g.mode = mode_dct1; break;
    }
    a.work = c.work; a.L = c.L; a.m_fft = c.m_fft; a.n1 = c.n1; a.n2 = c.n2;
    a.twiddle = c.twiddle; a.chirp = c.chirp; a.bhat = c.bhat; a.w4n = c.w4n;
    a.shift = shift;
    cuda_launcher L{stream};
    long long const cells = c.m_fft * c.L;
    long long const blocks = std::max<long long>(1, std::min<long long>((cells + 255) / 256, 148LL * 16));
    auto launch = [&](int which){
        switch(which){
            case 0: return is_float ? L.launch(composite_load_kernel<float>, blocks, 256, 0, a) : L.launch(composite_load_kernel<double>, blocks, 256, 0, a);
            case 1: return is_float ? L.launch(composite_twiddle_kernel<float>, blocks, 256, 0, a) : L.launch(composite_twiddle_kernel<double>, blocks, 256, 0, a);
            case 2: return is_float ? L.launch(composite_pointwise_kernel<float>, blocks, 256, 0, a) : L.launch(composite_pointwise_kernel<double>, blocks, 256, 0, a);
            default: return is_float ? L.launch(composite_store_kernel<float>, blocks, 256, 0, a) : L.launch(composite_store_kernel<double>, blocks, 256, 0, a);
        }
    };
    for(long long line0 = 0; line0 < g.nlines; line0 += c.L){
        a.line0 = line0;
        a.lines = std::min<long long>(c.L, g.nlines - line0);
        int rc = launch(0);
        // forward four-step: sub-transforms of length n1, twiddles, sub-transforms of length n2
        if (rc == 0) rc = b200_fft1d_execute(c.sub_a, B200_FORWARD, c.work, c.work, 1.0, stream);
        if (rc == 0 and c.n2 > 1){
            a.conjugate = 0;
            rc = launch(1);
            if (rc == 0) rc = b200_fft1d_execute(c.sub_c, B200_FORWARD, c.work, c.work, 1.0, stream);
        }
        if (rc == 0 and c.bluestein){
            rc = launch(2);
            // back: the steps undone in reverse order (permuted order in, natural order out)
            if (rc == 0 and c.n2 > 1){
                rc = b200_fft1d_execute(c.sub_c, B200_BACKWARD, c.work, c.work, 1.0, stream);
                a.conjugate = 1;
                if (rc == 0) rc = launch(1);
            }
            if (rc == 0) rc = b200_fft1d_execute(c.sub_a, B200_BACKWARD, c.work, c.work, 1.0, stream);
        }
        if (rc == 0) rc = launch(3);
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

} // namespace

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
// any mix of host and device pointers (unified addressing), synchronous -- what a GPU-aware message layer does with the buffers it is handed
int b200_copy_any(void *destination, const void *source, size_t bytes){
    if (bytes == 0) return B200_SUCCESS;
    return check_cuda(cudaMemcpy(destination, source, bytes, cudaMemcpyDefault), "copy");
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
    // lines beyond the shared-memory engine, and lines with a large prime factor (O(N p) in the generic kernel), take the
    // composite engine: four-step over sub-plans, Bluestein for lengths that do not split (composite.cuh)
    bool const too_long = (rc == B200_ERR_UNSUPPORTED and std::strstr(why, "shared-memory engine") != nullptr);
    bool const awkward = (rc == B200_SUCCESS and plan->host.family == family_generic and largest_prime_factor(plan->host.m) >= 128 and
                          std::getenv("HEFFTE_B200_NO_COMPOSITE") == nullptr);
    if (too_long or awkward){
        if (b200_device_count() < 1){ delete plan; return fail(B200_ERR_NO_DEVICE, "no CUDA device: the b200 backend has no CPU fallback"); }
        plan->host.desc = *desc;
        long long const engine = (desc->kind == B200_COS1) ? 2 * (desc->n - 1) : desc->n;
        rc = (desc->precision == B200_PREC_FLOAT) ? build_composite<float>(*plan, engine) : build_composite<double>(*plan, engine);
        if (rc){ delete plan; return rc; }
        *out = plan;
        return B200_SUCCESS;
    }
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
    if (plan->composite) return plan->composite->bluestein ? "composite (Bluestein)" : "composite (four-step)";
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
    if (plan->composite) return run_composite(*plan, direction, in, out, scale, static_cast<cudaStream_t>(stream), nullptr, batch_shift{0, 0});
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, out, scale, L);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

int b200_fft1d_execute_range(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream, long long b_begin, long long b_count){
    if (plan == nullptr) return fail(B200_ERR_INVALID, "null plan");
    if (plan->composite) return fail(B200_ERR_UNSUPPORTED, "line ranges are not available for composite plans");
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
    if (plan->composite){
        for(int e=0; e<batch; e++){
            int rc = run_composite(*plan, direction, static_cast<const char*>(in) + e * in_step, static_cast<char*>(out) + e * out_step, scale,
                                   static_cast<cudaStream_t>(stream), nullptr, batch_shift{0, 0});
            if (rc) return rc;
        }
        return B200_SUCCESS;
    }
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
    if (plan->composite){
        for(int e=0; e<batch; e++){
            int rc = run_composite(*plan, direction, static_cast<const char*>(in) + e * in_step, nullptr, scale, static_cast<cudaStream_t>(stream),
                                   device_scatter_map, batch_shift{e * scatter_step, local_shift + e * local_step});
            if (rc) return rc;
        }
        return B200_SUCCESS;
    }
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    batch_steps steps; steps.batch = batch; steps.in_step = in_step; steps.scatter_step = scatter_step; steps.local_shift = local_shift; steps.local_step = local_step;
    int rc = run_host_plan(plan->host, plan->twiddle, direction, in, nullptr, scale, L, device_scatter_map, 0, -1, steps);
    if (rc == -1) return fail(B200_ERR_UNSUPPORTED, "no kernel for this length");
    return rc;
}

} // extern "C"
