// Developer micro-benchmark (not part of the product): times variants (radix schedule, lines per CTA, occupancy) of the
// real-data FFT kernels on the 512^3 fp64 problem: r2c / DCT-II forward / DCT-III backward along the contiguous axis
// (fft_contig_real_kernel) and along the middle axis (fft_strided_real_kernel).
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

int main(){
    const int n = 512; const long long elems = (long long)n*n*n;
    double *x; double2 *y; CK(cudaMalloc(&x, elems * 8)); CK(cudaMalloc(&y, (long long)(n/2+1)*n*n * 16));
    CK(cudaMemset(x, 0, elems*8)); CK(cudaMemset(y, 0, (long long)(n/2+1)*n*n*16));
    host_plan hp; const char *why;
    b200_fft1d_desc d{}; d.precision = 1; d.kind = B200_COS; d.n = n; d.count_a = (long long)n*n; d.count_b = 1; d.in = {1, n, 0}; d.out = d.in;
    make_host_plan(d, hp, &why);
    auto table = make_twiddle_table<double>(hp); char *tw; CK(cudaMalloc(&tw, table.size()*8)); CK(cudaMemcpy(tw, table.data(), table.size()*8, cudaMemcpyHostToDevice));
    L l;
    const double gb_r2r = 2.0 * elems * 8 * 1e-9, gb_r2c = (elems * 8.0 + (double)(n/2+1)*n*n*16) * 1e-9;
    auto report = [&](const char *name, double gb, float ms){ printf("%-60s %8.3f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };

    fft_args a{}; a.twiddle = tw + 16 * (hp.table_main + hp.table_extra); a.twiddle2 = tw + 16 * hp.table_main;
    a.nlines = elems / n; a.scale = 1.0; a.smap = nullptr;
    using R884 = radix_list<8,8,4,1>; using R1616 = radix_list<16,16,1,1>; using R488 = radix_list<4,8,8,1>;

    // ---- contiguous axis --------------------------------------------------------------------------------------------
    a.count_a = n * n;
    for(int kind : {real_r2c, real_cos}){
        for(int backward=0; backward<2; backward++){
            a.backward = backward;
            if (kind == real_r2c){
                line_geom rg{1, n, 0}, cg{1, n/2+1, 0};
                a.in = backward ? (void*)y : (void*)x; a.out = backward ? (void*)x : (void*)y;
                a.ig = backward ? cg : rg; a.og = backward ? rg : cg;
            }else{ a.in = x; a.out = x; a.ig = a.og = line_geom{1, n, 0}; }
            double gb = (kind == real_r2c) ? gb_r2c : gb_r2r;
            printf("-- contig kind %s %s\n", kind == real_r2c ? "r2c" : "cos", backward ? "backward" : "forward");
#define CV(RL, LPB, MINB, label) if (kind == real_r2c) report(label, gb, timeit([&]{ launch_contig_real<double, RL, LPB, MINB, real_r2c, false>(a, l); })); \
                                 else report(label, gb, timeit([&]{ launch_contig_real<double, RL, LPB, MINB, real_cos, false>(a, l); }));
            CV(R884, 4, 6, "contig_real <8,8,4> LPB4 minb6 (current)")
            CV(R884, 2, 12, "contig_real <8,8,4> LPB2 minb12")
            CV(R884, 8, 3, "contig_real <8,8,4> LPB8 minb3")
            CV(R884, 4, 4, "contig_real <8,8,4> LPB4 minb4")
            CV(R488, 4, 6, "contig_real <4,8,8> LPB4 minb6")
            CV(R1616, 4, 6, "contig_real <16,16> LPB4 minb6 (64thr)")
            CV(R1616, 8, 3, "contig_real <16,16> LPB8 minb3 (128thr)")
            CV(R1616, 8, 4, "contig_real <16,16> LPB8 minb4 (128thr)")
            CV(R1616, 2, 12, "contig_real <16,16> LPB2 minb12 (32thr)")
        }
    }
    // ---- second-generation contiguous DCT kernel (register-direct Makhoul permutation) ---------------------------------------
    {
        a.count_a = n * n; a.in = x; a.out = x; a.ig = a.og = line_geom{1, n, 0};
        using R488 = radix_list<4,8,8,1>; using R884b = radix_list<8,8,4,1>; using R448 = radix_list<4,4,16,1>; using R1644 = radix_list<16,4,4,1>;
        auto run_dct = [&](auto kernel, int threads, size_t smem, int lpb){
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
            long long blocks = (a.nlines + lpb - 1) / lpb;
            kernel<<<(unsigned)blocks, threads, smem>>>(a);
        };
        constexpr size_t pitch = (pad_index(256) + 1) * 16;
        // correctness against the first-generation kernel, forward then backward
        {
            std::vector<double> h(elems), r1(elems), r2(elems);
            for(long long i=0; i<elems; i++) h[i] = (double)((i * 2654435761u) % 1000) * 1e-3;
            for(int backward=0; backward<2; backward++){
                a.backward = backward;
                CK(cudaMemcpy(x, h.data(), elems*8, cudaMemcpyHostToDevice));
                launch_contig_real<double, R1616, 4, 6, real_cos, false>(a, l); CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(r1.data(), x, elems*8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(x, h.data(), elems*8, cudaMemcpyHostToDevice));
                if (backward) run_dct(fft_contig_dct_kernel<double, R884b, 4, 4, real_cos, true, 32>, 128, pitch * 4, 4);
                else run_dct(fft_contig_dct_kernel<double, R488, 4, 4, real_cos, false, 32>, 128, pitch * 4, 4);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(r2.data(), x, elems*8, cudaMemcpyDeviceToHost));
                double err = 0, nrm = 0; for(long long i=0; i<elems; i++){ err += (r1[i]-r2[i])*(r1[i]-r2[i]); nrm += r1[i]*r1[i]; }
                printf("dct2 kernel vs first generation (%s): rel l2 %.3e\n", backward ? "backward" : "forward", sqrt(err / nrm));
            }
            CK(cudaMemset(x, 0, elems*8));
        }
        a.backward = 0;
    }
    // ---- second-generation contiguous kernel: two neighbouring lines per complex line, pairs (k, n-k) in the registers of one thread ----
    {
        a.count_a = n * n; a.twiddle0 = tw; a.in_step = a.out_step = 0;
        using C888 = radix_list<8,8,8,1>;
        for(int kind : {real_cos, real_r2c}){
            for(int backward=0; backward<2; backward++){
                a.backward = backward;
                if (kind == real_r2c){
                    line_geom rg{1, n, 0}, cg{1, n/2+1, 0};
                    a.in = backward ? (void*)y : (void*)x; a.out = backward ? (void*)x : (void*)y;
                    a.ig = backward ? cg : rg; a.og = backward ? rg : cg;
                }else{ a.in = x; a.out = x; a.ig = a.og = line_geom{1, n, 0}; }
                double gb = (kind == real_r2c) ? gb_r2c : gb_r2r;
                printf("-- contig real2 kind %s %s\n", kind == real_r2c ? "r2c" : "cos", backward ? "backward" : "forward");
#define C2(RL, LPB, MINB, label) if (kind == real_r2c) report(label, gb, timeit([&]{ launch_contig_real2<double, RL, LPB, MINB, real_r2c>(a, l); })); \
                                 else report(label, gb, timeit([&]{ launch_contig_real2<double, RL, LPB, MINB, real_cos>(a, l); }));
                C2(C888, 2, 8, "contig_real2 <8,8,8> LPB2 minb8 (64thr)")
                C2(C888, 2, 6, "contig_real2 <8,8,8> LPB2 minb6 (64thr)")
                C2(C888, 2, 12, "contig_real2 <8,8,8> LPB2 minb12 (64thr)")
                C2(C888, 4, 4, "contig_real2 <8,8,8> LPB4 minb4 (128thr)")
                C2(C888, 1, 12, "contig_real2 <8,8,8> LPB1 minb12 (32thr)")
                C2(C888, 1, 16, "contig_real2 <8,8,8> LPB1 minb16 (32thr)")
            }
        }
        a.backward = 0;
    }
    // ---- middle axis (stride n, neighbours adjacent) ----------------------------------------------------------------------
    a.in = x; a.out = x; a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n;
    for(int backward=0; backward<2; backward++){
        a.backward = backward;
        printf("-- strided kind cos %s\n", backward ? "backward" : "forward");
#define SV(RL, TPL, LPB, MINB, label) report(label, gb_r2r, timeit([&]{ launch_strided_real<double, RL, TPL, LPB, MINB, real_cos, false>(a, l); }));
        SV(R884, 16, 16, 3, "strided_real <8,8,4> TPL16 LPB16 minb3 (current)")
        SV(R884, 16, 16, 2, "strided_real <8,8,4> TPL16 LPB16 minb2")
        SV(R884, 32, 16, 1, "strided_real <8,8,4> TPL32 LPB16 minb1 (512thr)")
        SV(R884, 32, 8, 3, "strided_real <8,8,4> TPL32 LPB8 minb3 (256thr 32KB)")
        SV(R884, 32, 8, 6, "strided_real <8,8,4> TPL32 LPB8 minb6 (256thr 32KB)")
        SV(R884, 16, 8, 6, "strided_real <8,8,4> TPL16 LPB8 minb6 (128thr 32KB)")
        SV(R884, 8, 16, 3, "strided_real <8,8,4> TPL8 LPB16 minb3 (128thr)")
        SV(R1616, 16, 16, 3, "strided_real <16,16> TPL16 LPB16 minb3")
        SV(R1616, 16, 16, 2, "strided_real <16,16> TPL16 LPB16 minb2")
        SV(R1616, 16, 8, 4, "strided_real <16,16> TPL16 LPB8 minb4 (128thr 32KB)")
        SV(R488, 16, 16, 3, "strided_real <4,8,8> TPL16 LPB16 minb3")
    }
    // ---- second-generation strided real kernel: two adjacent real lines per complex line, full-length complex passes ---------
    {
        a.in = x; a.out = x; a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n; a.twiddle0 = tw; a.in_step = a.out_step = 0;
        using R888 = radix_list<8,8,8,1>; using R8164 = radix_list<8,16,4,1>; using R4816 = radix_list<4,8,16,1>;
        for(int backward=0; backward<2; backward++){
            a.backward = backward;
            printf("-- strided real2 kind cos %s\n", backward ? "backward" : "forward");
#define S2(RL, TPL, LPB, MINB, label) report(label, gb_r2r, timeit([&]{ launch_strided_real2<double, RL, TPL, LPB, MINB, real_cos>(a, l); }));
            S2(R888, 32, 8, 3, "strided_real2 <8,8,8> TPL32 LPB8 minb3")
            S2(R888, 32, 8, 2, "strided_real2 <8,8,8> TPL32 LPB8 minb2")
            S2(R888, 64, 8, 1, "strided_real2 <8,8,8> TPL64 LPB8 minb1 (512thr)")
            S2(R8164, 32, 8, 3, "strided_real2 <8,16,4> TPL32 LPB8 minb3")
            S2(R4816, 32, 8, 3, "strided_real2 <4,8,16> TPL32 LPB8 minb3")
        }
        line_geom rg{n, 1, (long long)n*n}, cg{n, 1, (long long)n*(n/2+1)};
        a.backward = 0; a.in = x; a.out = y; a.ig = rg; a.og = cg;
        printf("-- strided real2 kind r2c forward / backward\n");
        report("strided_real2 r2c fwd <8,8,8> TPL32 LPB8 minb3", gb_r2c, timeit([&]{ launch_strided_real2<double, R888, 32, 8, 3, real_r2c>(a, l); }));
        a.backward = 1; a.in = y; a.out = x; a.ig = cg; a.og = rg;
        report("strided_real2 c2r bwd <8,8,8> TPL32 LPB8 minb3", gb_r2c, timeit([&]{ launch_strided_real2<double, R888, 32, 8, 3, real_r2c>(a, l); }));
        a.backward = 0;
    }
    // r2c along the middle axis
    {
        line_geom rg{n, 1, (long long)n*n}, cg{n, 1, (long long)n*(n/2+1)};
        a.backward = 0; a.in = x; a.out = y; a.ig = rg; a.og = cg;
        printf("-- strided kind r2c forward\n");
        report("strided_real r2c <8,8,4> TPL16 LPB16 minb3", gb_r2c, timeit([&]{ launch_strided_real<double, R884, 16, 16, 3, real_r2c, false>(a, l); }));
        report("strided_real r2c <16,16> TPL16 LPB16 minb3", gb_r2c, timeit([&]{ launch_strided_real<double, R1616, 16, 16, 3, real_r2c, false>(a, l); }));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
