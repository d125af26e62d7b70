// Developer micro-benchmark (not part of the product): times FFT kernel variants and pure-copy kernels with the
// same access patterns on the 512^3 fp64 problem, to separate "access pattern ceiling" from "kernel inefficiency".
#include "../heffte_b200/csrc/fft_host_plan.h"
#include <cstdio>
#include <vector>
using namespace b200;

#define CK(x) do{ cudaError_t e = (x); if (e != cudaSuccess){ printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

struct L { cudaStream_t s = 0;
  template<typename K, typename A> int launch(K k, long long blocks, int threads, size_t smem, A const &a){
    if (smem > 48*1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
    k<<<(unsigned)blocks, threads, smem, s>>>(a); return 0; } };

// copy kernel with the strided-tile pattern: tile = LPB adjacent lines (LPB*16 bytes per row), N rows at `stride`
template<int LPB, int ROWS_PER_THREAD>
__global__ void copy_tile(const double2 *in, double2 *out, long long stride, int n, int count_a, long long stride_b, long long nlines){
    int t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    int tpl = blockDim.x / LPB;
    long long line = (long long)blockIdx.x * LPB + t;
    if (line >= nlines) return;
    long long b = line / count_a, a = line - b * count_a;
    long long off = a + b * stride_b;
    double2 v[ROWS_PER_THREAD];
    for(int base = 0; base < n; base += tpl * ROWS_PER_THREAD){
        #pragma unroll
        for(int r=0; r<ROWS_PER_THREAD; r++) v[r] = in[off + (long long)(base + j + r * tpl) * stride];
        #pragma unroll
        for(int r=0; r<ROWS_PER_THREAD; r++) out[off + (long long)(base + j + r * tpl) * stride] = v[r];
    }
}
__global__ void copy_linear(const double2 *in, double2 *out, long long n){
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long step = (long long)gridDim.x * blockDim.x;
    for(; i < n; i += step) out[i] = in[i];
}

template<typename F> float timeit(F f, int reps = 10){
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for(int i=0;i<3;i++) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a); for(int i=0;i<reps;i++) f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main(){
    const int n = 512; const long long elems = (long long)n*n*n; const double gb = 2.0 * elems * 16 * 1e-9;
    double2 *x, *y; CK(cudaMalloc(&x, elems * 16)); CK(cudaMalloc(&y, elems * 16));
    CK(cudaMemset(x, 0, elems*16)); CK(cudaMemset(y, 0, elems*16));
    host_plan hp; const char *why;
    b200_fft1d_desc d{}; d.precision = 1; d.kind = 0; d.n = n; d.count_a = n; d.count_b = n; d.in = {n, 1, (long long)n*n}; d.out = d.in;
    make_host_plan(d, hp, &why);
    auto table = make_twiddle_table<double>(hp); void *tw; CK(cudaMalloc(&tw, table.size()*8)); CK(cudaMemcpy(tw, table.data(), table.size()*8, cudaMemcpyHostToDevice));
    L l;
    auto report = [&](const char *name, float ms){ printf("%-52s %8.3f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); };

    report("copy_linear in->out", timeit([&]{ copy_linear<<<148*8, 256>>>(x, y, elems); }));
    report("copy_linear in-place", timeit([&]{ copy_linear<<<148*8, 256>>>(x, x, elems); }));
    // strided-tile copies, dim1 (stride n) and dim2 (stride n*n)
    report("copy_tile<8,8> dim1 in-place 256thr", timeit([&]{ copy_tile<8,8><<<elems/n/8, 256>>>(x, x, n, n, n, (long long)n*n, elems/n); }));
    report("copy_tile<8,8> dim2 in-place 256thr", timeit([&]{ copy_tile<8,8><<<elems/n/8, 256>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));
    report("copy_tile<16,8> dim1 in-place 256thr", timeit([&]{ copy_tile<16,8><<<elems/n/16, 256>>>(x, x, n, n, n, (long long)n*n, elems/n); }));
    report("copy_tile<16,8> dim2 in-place 256thr", timeit([&]{ copy_tile<16,8><<<elems/n/16, 256>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));
    report("copy_tile<32,8> dim2 in-place 256thr", timeit([&]{ copy_tile<32,8><<<elems/n/32, 256>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));
    report("copy_tile<8,16> dim2 in-place 256thr", timeit([&]{ copy_tile<8,16><<<elems/n/8, 256>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));
    report("copy_tile<8,8> dim2 in-place 512thr", timeit([&]{ copy_tile<8,8><<<elems/n/8, 512>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));
    report("copy_tile<4,8> dim2 in-place 256thr", timeit([&]{ copy_tile<4,8><<<elems/n/4, 256>>>(x, x, (long long)n*n, n, n*n, 0, elems/n); }));

    fft_args a; a.in = x; a.out = x; a.twiddle = tw; a.nlines = elems / n; a.backward = 0; a.scale = 1.0;
    using R512 = radix_list<8,8,8,1>;
    for(int dim=1; dim<=2; dim++){
        if (dim == 1){ a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n; }
        else { a.ig = a.og = line_geom{(long long)n*n, 1, 0}; a.count_a = n*n; }
        printf("-- dim %d\n", dim);
        report("strided TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R512, 32, 8, 3>(a, l); }));
        report("strided TPL32 LPB8 minb2", timeit([&]{ launch_strided<double, R512, 32, 8, 2>(a, l); }));
        report("strided TPL64 LPB8 minb1 (512thr)", timeit([&]{ launch_strided<double, R512, 64, 8, 1>(a, l); }));
        report("strided TPL64 LPB8 minb2 (512thr)", timeit([&]{ launch_strided<double, R512, 64, 8, 2>(a, l); }));
        report("strided TPL16 LPB8 minb3 (128thr)", timeit([&]{ launch_strided<double, R512, 16, 8, 3>(a, l); }));
        report("strided TPL32 LPB16 minb1 (512thr 128KB)", timeit([&]{ launch_strided<double, R512, 32, 16, 1>(a, l); }));
        report("strided TPL16 LPB16 minb1 (256thr 128KB)", timeit([&]{ launch_strided<double, R512, 16, 16, 1>(a, l); }));
    }
    a.ig = a.og = line_geom{1, n, 0}; a.count_a = n*n;
    printf("-- dim 0\n");
    report("contig 512 LPB4 minb3", timeit([&]{ launch_contig<double, R512, 4, 3>(a, l); }));
    report("contig 512 LPB2 minb4", timeit([&]{ launch_contig<double, R512, 2, 4>(a, l); }));
    report("contig 512 LPB2 minb6", timeit([&]{ launch_contig<double, R512, 2, 6>(a, l); }));
    report("contig 512 LPB1 minb8", timeit([&]{ launch_contig<double, R512, 1, 8>(a, l); }));
    report("contig 512 LPB1 minb12", timeit([&]{ launch_contig<double, R512, 1, 12>(a, l); }));
    CK(cudaDeviceSynchronize());
    return 0;
}
