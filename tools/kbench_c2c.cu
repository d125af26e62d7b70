// Developer micro-benchmark (not part of the product): radix schedule / tile shape / occupancy variants of the c2c kernels
// on the two single-GPU BASELINE problems: 512^3 fp64 (all three axes) and 256^3 fp32.  Complements tools/kbench.cu.
#include "../heffte_b200/csrc/fft_host_plan.h"
#include <cstdio>
#include <vector>
using namespace b200;

#define CK(x) do{ cudaError_t e = (x); if (e != cudaSuccess){ printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

struct L { cudaStream_t s = 0;
  template<typename K, typename A> int launch(K k, long long blocks, int threads, size_t smem, A const &a){
    if (smem > 48*1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
    k<<<(unsigned)blocks, threads, smem, s>>>(a); return 0; } };

template<typename F> float timeit(F f, int reps = 20){
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for(int i=0;i<3;i++) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a); for(int i=0;i<reps;i++) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / reps;
}

template<typename T> void* twiddles(int n){
    host_plan hp; const char *why;
    b200_fft1d_desc d{}; d.precision = sizeof(T) == 4 ? 0 : 1; d.kind = 0; d.n = n; d.count_a = n; d.count_b = n; d.in = {n, 1, (long long)n*n}; d.out = d.in;
    make_host_plan(d, hp, &why);
    auto table = make_twiddle_table<T>(hp); void *tw; CK(cudaMalloc(&tw, table.size()*sizeof(T))); CK(cudaMemcpy(tw, table.data(), table.size()*sizeof(T), cudaMemcpyHostToDevice));
    return tw;
}

int main(){
    L l;
    {   // ---- 512^3 fp64 -----------------------------------------------------------------------------------------------------
        const int n = 512; const long long elems = (long long)n*n*n; const double gb = 2.0 * elems * 16 * 1e-9;
        double2 *x; CK(cudaMalloc(&x, elems * 16)); CK(cudaMemset(x, 0, elems*16));
        auto report = [&](const char *name, float ms){ printf("%-60s %8.3f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
        fft_args a; a.in = x; a.out = x; a.twiddle = twiddles<double>(n); a.twiddle2 = nullptr; a.nlines = elems / n; a.backward = 0; a.scale = 1.0; a.smap = nullptr;
        using R888 = radix_list<8,8,8,1>; using R1684 = radix_list<16,8,4,1>; using R16162 = radix_list<16,16,2,1>; using R4816 = radix_list<4,8,16,1>; using R8164 = radix_list<8,16,4,1>;
        for(int dim=1; dim<=2; dim++){
            if (dim == 1){ a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n; }
            else { a.ig = a.og = line_geom{(long long)n*n, 1, 0}; a.count_a = n*n; }
            printf("-- fp64 512 strided, dim %d\n", dim);
            report("strided <8,8,8> TPL32 LPB8 minb3 (current)", timeit([&]{ launch_strided<double, R888, 32, 8, 3, false>(a, l); }));
            report("strided <16,8,4> TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R1684, 32, 8, 3, false>(a, l); }));
            report("strided <8,16,4> TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R8164, 32, 8, 3, false>(a, l); }));
            report("strided <16,16,2> TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R16162, 32, 8, 3, false>(a, l); }));
            report("strided <4,8,16> TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R4816, 32, 8, 3, false>(a, l); }));
            report("strided <16,8,4> TPL16 LPB8 minb3 (128thr)", timeit([&]{ launch_strided<double, R1684, 16, 8, 3, false>(a, l); }));
            report("strided <8,8,8> TPL64 LPB4 minb3 (256thr 32KB)", timeit([&]{ launch_strided<double, R888, 64, 4, 3, false>(a, l); }));
            report("strided <8,8,8> TPL64 LPB4 minb6 (256thr 32KB)", timeit([&]{ launch_strided<double, R888, 64, 4, 6, false>(a, l); }));
        }
        a.ig = a.og = line_geom{1, n, 0}; a.count_a = n*n;
        printf("-- fp64 512 contiguous\n");
        report("contig <8,8,8> LPB1 minb12 (current)", timeit([&]{ launch_contig<double, R888, 1, 12, false>(a, l); }));
        report("contig <16,8,4> LPB1 minb12 (32thr)", timeit([&]{ launch_contig<double, R1684, 1, 12, false>(a, l); }));
        report("contig <16,8,4> LPB2 minb8 (64thr)", timeit([&]{ launch_contig<double, R1684, 2, 8, false>(a, l); }));
        report("contig <16,16,2> LPB2 minb8 (64thr)", timeit([&]{ launch_contig<double, R16162, 2, 8, false>(a, l); }));
        report("contig <16,16,2> LPB4 minb4 (128thr)", timeit([&]{ launch_contig<double, R16162, 4, 4, false>(a, l); }));
        report("contig <4,8,16> LPB2 minb8 (64thr)", timeit([&]{ launch_contig<double, R4816, 2, 8, false>(a, l); }));
        CK(cudaFree(x));
    }
    {   // ---- 256^3 fp32 -----------------------------------------------------------------------------------------------------
        const int n = 256; const long long elems = (long long)n*n*n; const double gb = 2.0 * elems * 8 * 1e-9;
        float2 *x; CK(cudaMalloc(&x, elems * 8)); CK(cudaMemset(x, 0, elems*8));
        auto report = [&](const char *name, float ms){ printf("%-60s %8.4f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
        fft_args a; a.in = x; a.out = x; a.twiddle = twiddles<float>(n); a.twiddle2 = nullptr; a.nlines = elems / n; a.backward = 0; a.scale = 1.0; a.smap = nullptr;
        using R884 = radix_list<8,8,4,1>; using R1616 = radix_list<16,16,1,1>; using R488 = radix_list<4,8,8,1>;
        for(int dim=1; dim<=2; dim++){
            if (dim == 1){ a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n; }
            else { a.ig = a.og = line_geom{(long long)n*n, 1, 0}; a.count_a = n*n; }
            printf("-- fp32 256 strided, dim %d\n", dim);
            report("strided <8,8,4> TPL16 LPB16 minb2 (current)", timeit([&]{ launch_strided<float, R884, 16, 16, 2, false>(a, l); }));
            report("strided <8,8,4> TPL16 LPB16 minb4", timeit([&]{ launch_strided<float, R884, 16, 16, 4, false>(a, l); }));
            report("strided <16,16> TPL16 LPB16 minb2", timeit([&]{ launch_strided<float, R1616, 16, 16, 2, false>(a, l); }));
            report("strided <16,16> TPL16 LPB16 minb4", timeit([&]{ launch_strided<float, R1616, 16, 16, 4, false>(a, l); }));
            report("strided <16,16> TPL8 LPB16 minb4 (128thr)", timeit([&]{ launch_strided<float, R1616, 8, 16, 4, false>(a, l); }));
            report("strided <8,8,4> TPL32 LPB16 minb2 (512thr)", timeit([&]{ launch_strided<float, R884, 32, 16, 2, false>(a, l); }));
            report("strided <8,8,4> TPL8 LPB32 minb2 (256thr, 256B rows)", timeit([&]{ launch_strided<float, R884, 8, 32, 2, false>(a, l); }));
        }
        a.ig = a.og = line_geom{1, n, 0}; a.count_a = n*n;
        printf("-- fp32 256 contiguous\n");
        report("contig <8,8,4> LPB4 minb6 (current)", timeit([&]{ launch_contig<float, R884, 4, 6, false>(a, l); }));
        report("contig <16,16> LPB4 minb6 (64thr)", timeit([&]{ launch_contig<float, R1616, 4, 6, false>(a, l); }));
        report("contig <16,16> LPB8 minb4 (128thr)", timeit([&]{ launch_contig<float, R1616, 8, 4, false>(a, l); }));
        report("contig <4,8,8> LPB4 minb6", timeit([&]{ launch_contig<float, R488, 4, 6, false>(a, l); }));
        report("contig <8,8,4> LPB8 minb4 (256thr)", timeit([&]{ launch_contig<float, R884, 8, 4, false>(a, l); }));
        CK(cudaFree(x));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
