// Developer micro-benchmark (not part of the product): the paired kernel (two local transforms of a box in one persistent launch,
// the second reading the output of the first from the L2 cache) against the two separate launches, 512^3 fp64 and 256^3 fp32.
#include "../heffte_b200/csrc/fft_host_plan.h"
#include <cstdio>
#include <vector>
using namespace b200;
#define CK(x) do{ cudaError_t e = (x); if (e != cudaSuccess){ printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

struct L { cudaStream_t s = 0;
  template<typename K, typename A> int launch(K k, long long blocks, int threads, size_t smem, A const &a){
    if (smem > 48*1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
    k<<<(unsigned)blocks, threads, smem, s>>>(a); return 0; } };

template<typename F> float timeit(F f, int reps = 10){
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for(int i=0;i<3;i++) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a); for(int i=0;i<reps;i++) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / reps;
}

template<typename T, typename RLA, int LPBA, int TPLA, typename RLB, int TPLB, int LPBB, int MINB, bool CONTIG_FIRST>
float run_pair(pair_args p, unsigned lag, int ctas_per_sm_limit = 0){
    auto kernel = fft_pair_kernel<T, RLA, LPBA, TPLA, RLB, TPLB, LPBB, MINB, false, false, CONTIG_FIRST>;
    size_t smem = pair_smem_bytes<T, RLA, LPBA, RLB, LPBB, false>();
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPLB * LPBB, smem));
    if (ctas_per_sm_limit > 0 && per_sm > ctas_per_sm_limit) per_sm = ctas_per_sm_limit;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    p.lag = lag;
    p.tiles_a = (unsigned)(p.a.count_a / LPBA); p.tiles_b = (unsigned)(p.b.count_a / LPBB);
    return timeit([&]{
        CK(cudaMemsetAsync(p.done, 0, sizeof(unsigned) * p.planes, 0));
        kernel<<<per_sm * sms, TPLB * LPBB, smem, 0>>>(p);
    });
}

template<typename T>
void bench(int n){
    const long long elems = (long long)n * n * n;
    using C = cplx<T>;
    C *x; CK(cudaMalloc(&x, elems * sizeof(C))); CK(cudaMemset(x, 0, elems * sizeof(C)));
    // a recognisable input: check the pair against the two separate launches
    std::vector<C> h(elems);
    for(long long i=0; i<elems; i++){ h[i].x = (T)((i * 2654435761u % 1000) * 1e-3); h[i].y = (T)((i * 40503u % 777) * 1e-3); }
    host_plan hp; const char *why;
    b200_fft1d_desc d{}; d.precision = sizeof(T) == 4 ? 0 : 1; d.kind = B200_C2C; d.n = n; d.count_a = n; d.count_b = n; d.in = {1, n, (long long)n*n}; d.out = d.in;
    make_host_plan(d, hp, &why);
    auto table = make_twiddle_table<T>(hp); char *tw; CK(cudaMalloc(&tw, table.size()*sizeof(T))); CK(cudaMemcpy(tw, table.data(), table.size()*sizeof(T), cudaMemcpyHostToDevice));
    unsigned *counters; CK(cudaMalloc(&counters, sizeof(unsigned) * (n + 1)));
    L l;
    fft_args A; A.in = x; A.out = x; A.twiddle = tw; A.twiddle2 = nullptr; A.ig = A.og = line_geom{1, n, (long long)n*n}; A.nlines = (long long)n*n; A.count_a = n; A.backward = 0; A.scale = 1.0; A.smap = nullptr;
    fft_args B = A; B.ig = B.og = line_geom{n, 1, (long long)n*n};
    pair_args p; p.a = A; p.b = B; p.planes = n; p.done = counters;
    const double gb = 4.0 * elems * sizeof(C) * 1e-9;     // algorithmic bytes of the two passes
    auto report = [&](const char *name, float ms){ printf("%-72s %8.4f ms  %7.1f GB/s algorithmic\n", name, ms, gb / ms * 1e3); fflush(stdout); };
    std::vector<C> ref(elems), got(elems);
    if (n == 512){
        using RA = radix_list<4,8,16,1>; using RA8 = radix_list<8,8,8,1>; using RB = radix_list<8,8,8,1>;
        CK(cudaMemcpy(x, h.data(), elems*sizeof(C), cudaMemcpyHostToDevice));
        launch_contig<T, RA, 2, 8, false>(A, l); launch_strided<T, RB, 32, 8, 3, false>(B, l);
        CK(cudaMemcpy(ref.data(), x, elems*sizeof(C), cudaMemcpyDeviceToHost));
        report("separate: contig <4,8,16> LPB2 + strided <8,8,8> TPL32 LPB8", timeit([&]{ launch_contig<T, RA, 2, 8, false>(A, l); launch_strided<T, RB, 32, 8, 3, false>(B, l); }));
        CK(cudaMemcpy(x, h.data(), elems*sizeof(C), cudaMemcpyHostToDevice));
        { pair_args q = p; q.lag = 4; q.tiles_a = n / 8; q.tiles_b = n / 8;
          auto kernel = fft_pair_kernel<T, RA, 8, 32, RB, 32, 8, 3, false, false, true>;
          size_t smem = pair_smem_bytes<T, RA, 8, RB, 8, false>();
          CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          CK(cudaMemset(counters, 0, sizeof(unsigned) * (n + 1)));
          kernel<<<148 * 3, 256, smem>>>(q); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(got.data(), x, elems*sizeof(C), cudaMemcpyDeviceToHost));
        double err = 0, nrm = 0; for(long long i=0; i<elems; i++){ double dx = got[i].x - ref[i].x, dy = got[i].y - ref[i].y; err += dx*dx + dy*dy; nrm += (double)ref[i].x*ref[i].x + (double)ref[i].y*ref[i].y; }
        printf("pair vs separate launches: rel l2 %.3e\n", sqrt(err / nrm));
        for(unsigned lag : {1u, 2u, 4u, 8u, 16u}){
            char name[128]; snprintf(name, sizeof(name), "pair contig-first A<4,8,16> LPB8 (74KB) minb3 lag %u", lag);
            report(name, run_pair<T, RA, 8, 32, RB, 32, 8, 3, true>(p, lag));
        }
        for(unsigned lag : {2u, 4u, 8u}){
            char name[128]; snprintf(name, sizeof(name), "pair contig-first A<8,8,8> TPL64 LPB4 (64KB) minb3 lag %u", lag);
            report(name, run_pair<T, RA8, 4, 64, RB, 32, 8, 3, true>(p, lag));
        }
        for(int lim : {1, 2}){
            char name[128]; snprintf(name, sizeof(name), "pair contig-first A<4,8,16> LPB8 lag 4, %d CTA/SM", lim);
            report(name, run_pair<T, RA, 8, 32, RB, 32, 8, 3, true>(p, 4, lim));
        }
        for(unsigned lag : {2u, 4u, 8u}){
            char name[128]; snprintf(name, sizeof(name), "pair strided-first (backward order) A<4,8,16> LPB8 lag %u", lag);
            report(name, run_pair<T, RA, 8, 32, RB, 32, 8, 3, false>(p, lag));
        }
    }else if (n == 256){
        using RA = radix_list<16,16,1,1>; using RB = radix_list<8,8,4,1>;
        report("separate: contig <16,16> LPB4 + strided <8,8,4> TPL16 LPB16", timeit([&]{ launch_contig<T, RA, 4, 6, false>(A, l); launch_strided<T, RB, 16, 16, 2, false>(B, l); }));
        for(unsigned lag : {2u, 4u, 8u, 16u, 32u}){
            char name[128]; snprintf(name, sizeof(name), "pair contig-first A<16,16> LPB16 B<8,8,4> TPL16 LPB16 minb4 lag %u", lag);
            report(name, run_pair<T, RA, 16, 16, RB, 16, 16, 4, true>(p, lag));
        }
        for(unsigned lag : {4u, 16u}){
            char name[128]; snprintf(name, sizeof(name), "pair strided-first lag %u", lag);
            report(name, run_pair<T, RA, 16, 16, RB, 16, 16, 4, false>(p, lag));
        }
    }
    CK(cudaFree(x)); CK(cudaFree(tw)); CK(cudaFree(counters));
}

int main(){
    printf("-- 512^3 fp64: dims 0 + 1 (4 x 2.147 GB algorithmic)\n"); bench<double>(512);
    printf("-- 256^3 fp32: dims 0 + 1\n"); bench<float>(256);
    return 0;
}
