// TEST INFRASTRUCTURE ONLY -- runs the product's FFT kernel source (heffte_b200/csrc/fft_device.cuh) on the CPU
// through tests/emul/cuda_emul.h so the CPU-only test-suite can check index arithmetic.  Not part of the product.
#ifndef B200_HOST_EMULATION
#define B200_HOST_EMULATION
#endif
#include "fft_host_plan.h"

namespace {
struct emul_launcher {
    int batch = 1;
    long long max_blocks = 0;
    template<typename kernel_t, typename args_t>
    int launch(kernel_t kernel, long long blocks, int threads, size_t smem, args_t const &args){
        emul::launch(kernel, dim3((unsigned) blocks, (unsigned) batch), dim3((unsigned) threads), smem, args);
        return 0;
    }
    int run_pow2(bool strided, bool is_float, bool scatter, int n, b200::fft_args const &a){
        using namespace b200;
        if (strided){
            if (is_float) return scatter ? dispatch_strided<float, true>(n, a, *this) : dispatch_strided<float, false>(n, a, *this);
            return scatter ? dispatch_strided<double, true>(n, a, *this) : dispatch_strided<double, false>(n, a, *this);
        }
        if (is_float) return scatter ? dispatch_contig<float, true>(n, a, *this) : dispatch_contig<float, false>(n, a, *this);
        return scatter ? dispatch_contig<double, true>(n, a, *this) : dispatch_contig<double, false>(n, a, *this);
    }
    int run_real(bool strided, bool is_float, bool scatter, int kind, int m, b200::fft_args const &a){
        using namespace b200;
        if (not strided and not scatter and contig_real2_applies(kind, m, a)){
            int const rc = is_float ? dispatch_contig_real2<float>(kind, 2 * m, a, *this) : dispatch_contig_real2<double>(kind, 2 * m, a, *this);
            if (rc != -1) return rc;
        }
        if (strided and not scatter and real2_applies(is_float, kind, m, a)){
            int const rc = is_float ? dispatch_strided_real2<float>(kind, 2 * m, a, *this) : dispatch_strided_real2<double>(kind, 2 * m, a, *this);
            if (rc != -1) return rc;
        }
        if (strided){
            if (is_float) return scatter ? dispatch_strided_real<float, true>(kind, m, a, *this) : dispatch_strided_real<float, false>(kind, m, a, *this);
            return scatter ? dispatch_strided_real<double, true>(kind, m, a, *this) : dispatch_strided_real<double, false>(kind, m, a, *this);
        }
        if (is_float) return scatter ? dispatch_contig_real<float, true>(kind, m, a, *this) : dispatch_contig_real<float, false>(kind, m, a, *this);
        return scatter ? dispatch_contig_real<double, true>(kind, m, a, *this) : dispatch_contig_real<double, false>(kind, m, a, *this);
    }
    int run_generic(bool is_float, long long blocks, int threads, size_t smem, b200::generic_args const &g){
        if (is_float) return launch(b200::fft_generic_kernel<float>, blocks, threads, smem, g);
        return launch(b200::fft_generic_kernel<double>, blocks, threads, smem, g);
    }
};
}

// `scatter` = host pointer to a b200::scatter_map (or null): the fused-reshape store
extern "C" __attribute__((visibility("default"))) int emul_fft1d_scatter(const b200_fft1d_desc *desc, int direction, const void *in, const void *scatter, double scale, int *family){
    b200::host_plan plan;
    const char *why = "";
    int rc = b200::make_host_plan(*desc, plan, &why);
    if (rc) return rc;
    *family = (int) plan.family;
    emul_launcher L;
    if (desc->precision == B200_PREC_FLOAT){
        auto table = b200::make_twiddle_table<float>(plan);
        return b200::run_host_plan(plan, table.data(), direction, in, nullptr, scale, L, scatter);
    }
    auto table = b200::make_twiddle_table<double>(plan);
    return b200::run_host_plan(plan, table.data(), direction, in, nullptr, scale, L, scatter);
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

// ---- fused reshape (scatter) through the product's host builder and kernels ---------------------------------------------
#include "scatter_build.h"
#include "pack_host.h"

namespace {
b200::box3 box_from9(const int *nine){
    return b200::box3({{nine[0], nine[1], nine[2]}}, {{nine[3], nine[4], nine[5]}}, {{nine[6], nine[7], nine[8]}});
}
bool make_map(const int *mine9, int k_dim, int nranks, const int *dest9, void *const *dest_base, int elem_bytes, b200::scatter_map &map){
    b200::box3 mine = box_from9(mine9);
    std::vector<b200::box3> dest;
    std::vector<void*> bases;
    for(int r=0; r<nranks; r++){ dest.push_back(box_from9(dest9 + 9 * r)); bases.push_back(dest_base[r]); }
    std::string why;
    return b200::build_scatter_map(mine, (k_dim < 0) ? 0 : mine.position_of(k_dim), dest, bases, elem_bytes, map, why);
}
}

// transform along `k_dim` of the box mine9 (desc describes its lines, un-lumped) with the store scattered into the boxes dest9
extern "C" __attribute__((visibility("default")))
int emul_fft1d_reshape(const b200_fft1d_desc *desc, int direction, const void *in, const int *mine9, int k_dim, int nranks, const int *dest9,
                       void *const *dest_base, int elem_bytes_out, double scale){
    b200::scatter_map map;
    if (not make_map(mine9, k_dim, nranks, dest9, dest_base, elem_bytes_out, map)) return 100;
    int family = 0;
    return emul_fft1d_scatter(desc, direction, in, &map, scale, &family);
}

extern "C" __attribute__((visibility("default")))
int emul_scatter_copy(int elem_bytes, const void *src, const int *mine9, int nranks, const int *dest9, void *const *dest_base){
    b200::scatter_map map;
    if (not make_map(mine9, -1, nranks, dest9, dest_base, elem_bytes, map)) return 100;
    b200::box3 mine = box_from9(mine9);
    if (mine.empty()) return 0;
    b200::scatter_copy_args a{src, &map, (int) mine.osize(0), (int) mine.osize(1), (int) mine.osize(2), mine.osize(0), mine.osize(0) * mine.osize(1), 1};
    emul_launcher L;
    return b200::launch_scatter_copy(elem_bytes, a, L);
}
