// Pure host-side planning of a batched 1-D transform (no CUDA calls): kernel family, engine length,
// factorisation, tile shape and the twiddle table.  Shared by fft1d.cu and by the CPU emulation in tests/emul.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/heffte_b200_kernels.h"
#include "fft_dispatch.cuh"

namespace b200 {

enum kernel_family { family_strided = 0, family_contig = 1, family_generic = 2, family_contig_real = 3, family_strided_real = 4 };

struct host_plan {
    b200_fft1d_desc desc;
    kernel_family family = family_generic;
    int m = 0;              // engine (complex FFT) length
    int lpb = 1;            // generic: lines per block
    int lines_fast = 0;     // generic: neighbouring lines are adjacent in memory
    int nfactors = 0;
    int factors[24];
    size_t smem = 0;        // generic: dynamic shared memory
    int threads = 256;      // generic: block size
    long long table_main = 0, table_extra_mod = 1, table_extra = 0;  // twiddle table layout
    long long table_third = 0;   // family_contig_real: W_{n/2}^k for the complex engine, behind the two other segments
    int real_kind = 0;           // family_contig_real: real_r2c / real_cos / real_sin
};

inline std::vector<int> factorize(int m){
    std::vector<int> f;
    while(m % 4 == 0){ f.push_back(4); m /= 4; }
    while(m % 2 == 0){ f.push_back(2); m /= 2; }
    for(int p = 3; (long long)p * p <= m; p += 2)
        while(m % p == 0){ f.push_back(p); m /= p; }
    if (m > 1) f.push_back(m);
    return f;
}

// W_m^k = exp(-2 pi i k / m) evaluated in long double, rounded once to the working precision
template<typename T>
void fill_twiddles(std::vector<T> &table, size_t offset, long long m, long long count){
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for(long long k=0; k<count; k++){
        long double angle = two_pi * static_cast<long double>(k % m) / static_cast<long double>(m);
        table[2 * (offset + k)]     = static_cast<T>(cosl(angle));
        table[2 * (offset + k) + 1] = static_cast<T>(-sinl(angle));
    }
}
template<typename T>
std::vector<T> make_twiddle_table(host_plan const &plan){
    std::vector<T> table(2 * (plan.table_main + plan.table_extra + plan.table_third));
    fill_twiddles<T>(table, 0, plan.table_main, plan.table_main);
    if (plan.table_extra > 0) fill_twiddles<T>(table, plan.table_main, plan.table_extra_mod, plan.table_extra);
    if (plan.table_third > 0) fill_twiddles<T>(table, plan.table_main + plan.table_extra, plan.table_third, plan.table_third);
    return table;
}

// returns 0 or a B200_ERR_* code; `why` receives the reason on failure
inline int make_host_plan(b200_fft1d_desc const &desc, host_plan &plan, const char **why){
    *why = "";
    if (desc.n < 1 or desc.count_a < 0 or desc.count_b < 0){ *why = "bad sizes"; return B200_ERR_INVALID; }
    if (desc.precision != B200_PREC_FLOAT and desc.precision != B200_PREC_DOUBLE){ *why = "bad precision"; return B200_ERR_INVALID; }
    if (desc.kind < B200_C2C or desc.kind > B200_COS1){ *why = "bad kind"; return B200_ERR_INVALID; }
    if (desc.n > 2147483647LL / 8){ *why = "length too large"; return B200_ERR_UNSUPPORTED; }
    if (desc.kind == B200_COS1 and desc.n < 2){ *why = "DCT-I needs at least 2 points"; return B200_ERR_INVALID; }
    plan.desc = desc;
    const int n = static_cast<int>(desc.n);
    const size_t csize = (desc.precision == B200_PREC_FLOAT) ? 8 : 16;

    bool const fast_path = (desc.kind == B200_C2C) and is_fast_length(n);
    if (fast_path){
        bool const contiguous = (desc.in.stride == 1 or desc.out.stride == 1);
        plan.family = contiguous ? family_contig : family_strided;
        plan.m = n;
        plan.table_main = n;
        return B200_SUCCESS;
    }
    plan.family = family_generic;
    plan.m = (desc.kind == B200_COS1) ? 2 * (n - 1) : n;
    auto f = factorize(plan.m);
    if (f.size() > 24){ *why = "too many prime factors"; return B200_ERR_UNSUPPORTED; }
    plan.nfactors = static_cast<int>(f.size());
    for(size_t i=0; i<f.size(); i++) plan.factors[i] = f[i];
    // lines per block: aim at ~4096 points per CTA; the two Stockham buffers must fit shared memory
    size_t const per_line = 2 * csize * static_cast<size_t>(plan.m);
    size_t const budget = 200 * 1024;
    if (per_line > budget){ *why = "transform length exceeds the shared-memory engine"; return B200_ERR_UNSUPPORTED; }
    long long lpb = std::max<long long>(1, 4096 / plan.m);
    lpb = std::min<long long>(lpb, static_cast<long long>(budget / per_line));
    lpb = std::min<long long>(lpb, 64);
    long long const nlines = desc.count_a * desc.count_b;
    lpb = std::max<long long>(1, std::min<long long>(lpb, nlines));
    plan.lpb = static_cast<int>(lpb);
    plan.smem = per_line * lpb;
    plan.lines_fast = (desc.in.stride != 1 and desc.in.stride_a == 1) ? 1 : 0;
    long long const work = lpb * plan.m;
    plan.threads = (work >= 256) ? 256 : ((work >= 128) ? 128 : ((work >= 64) ? 64 : 32));
    plan.table_main = plan.m;
    if (desc.kind == B200_COS or desc.kind == B200_SIN){ plan.table_extra_mod = 4LL * n; plan.table_extra = n; }
    // power-of-two real transforms along contiguous lines: the half-length complex engine (fft_contig_real_kernel); the
    // generic plan above stays as the path for pointers that are not aligned to a complex number
    bool const real_kind_ok = (desc.kind == B200_R2C or desc.kind == B200_COS or desc.kind == B200_SIN);
    bool const real_contig = (desc.in.stride == 1 and desc.out.stride == 1);
    bool const real_strided = (desc.in.stride != 1 and desc.out.stride != 1 and desc.in.stride_a == 1 and desc.out.stride_a == 1);
    if (real_kind_ok and is_fast_real_length(n) and (real_contig or real_strided)){
        plan.family = real_contig ? family_contig_real : family_strided_real;
        plan.real_kind = (desc.kind == B200_R2C) ? real_r2c : ((desc.kind == B200_COS) ? real_cos : real_sin);
        plan.table_extra_mod = 4LL * n; plan.table_extra = n + 1;      // W_{4n}^j, j = 0..n (the generic path reads j < n)
        plan.table_third = n / 2;
    }
    return B200_SUCCESS;
}

inline line_geom to_geom(b200_line_geom const &g){ return line_geom{g.stride, g.stride_a, g.stride_b}; }

// the second-generation strided real kernel pairs adjacent lines: an even number of them per row, rows aligned to a complex number
// contiguous lines: pairs of neighbouring lines (real lines at any alignment: those loads are scalar)
inline bool contig_real2_applies(int kind, int m, fft_args const &a){
    static bool const off = (std::getenv("HEFFTE_B200_REAL_KERNELS_V1") != nullptr);
    return not off and (kind == real_cos or kind == real_sin or kind == real_r2c) and is_real2_length(2LL * m) and a.nlines % 2 == 0 and a.count_a % 2 == 0 and
           a.ig.stride == 1 and a.og.stride == 1;
}
inline bool real2_applies(bool is_float, int kind, int m, fft_args const &a){
    static bool const off = (std::getenv("HEFFTE_B200_REAL_KERNELS_V1") != nullptr);
    if (off or not is_real2_length(2LL * m) or a.count_a % 2 != 0 or a.nlines % 2 != 0) return false;
    size_t const csize = is_float ? 8 : 16;
    bool const real_in = not (kind == real_r2c and a.backward), real_out = not (kind == real_r2c and not a.backward);
    auto even = [](line_geom const &g){ return g.stride % 2 == 0 and g.stride_b % 2 == 0 and g.stride_a == 1; };
    if (reinterpret_cast<uintptr_t>(a.in) % csize or reinterpret_cast<uintptr_t>(a.out) % csize or (a.in_step % static_cast<long long>(csize)) or (a.out_step % static_cast<long long>(csize))) return false;
    if (real_in and not even(a.ig)) return false;
    if (real_out and not even(a.og)) return false;
    if (not real_in and a.ig.stride_a != 1) return false;
    if (not real_out and a.og.stride_a != 1) return false;
    return true;
}

// batched execution: `batch` entries, entry e works on in + e * in_step / out + e * out_step (bytes); fused stores: see fft_args
struct batch_steps {
    int batch = 1;
    long long in_step = 0, out_step = 0, scatter_step = 0, local_shift = 0, local_step = 0;
    // plane-wise overlap of two launches (fft_args::done ...): complex fast-path kernels only
    unsigned *done = nullptr;
    int done_mode = 0;
    unsigned done_need = 0;
    int order_nb = 0;
    long long max_blocks = 0;       // > 0: a thin grid (the kernels walk their tiles grid-stride)
};
template<typename args_t> inline void set_steps(args_t &a, batch_steps const &s){
    a.in_step = s.in_step; a.out_step = s.out_step; a.scatter_step = s.scatter_step; a.local_shift = s.local_shift; a.local_step = s.local_step;
}
inline void set_hooks(fft_args &a, batch_steps const &s){ a.done = s.done; a.done_mode = s.done_mode; a.done_need = s.done_need; a.order_nb = s.order_nb; }

// runs the plan through a Launcher (CUDA stream launcher in the product, thread emulation in tests/emul)
// `scatter` (device pointer to a scatter_map, or null) fuses the following reshape into the store of the transform
// b_begin / b_count (b_count >= 0) restrict the launch to the lines with b in [b_begin, b_begin + b_count): a slab of the box
template<typename Launcher>
int run_host_plan(host_plan const &plan, const void *twiddle, int direction, const void *in, void *out, double scale, Launcher &L,
                  const void *scatter = nullptr, long long b_begin = 0, long long b_count = -1, batch_steps const &steps = batch_steps()){
    L.batch = steps.batch;
    L.max_blocks = 0;
    b200_fft1d_desc const &d = plan.desc;
    bool const backward = (direction == B200_BACKWARD);
    bool const is_float = (d.precision == B200_PREC_FLOAT);
    long long nlines = d.count_a * d.count_b;
    if (b_count >= 0){
        if (b_begin < 0 or b_begin + b_count > d.count_b or scatter != nullptr) return B200_ERR_INVALID;
        size_t const rsize = is_float ? 4 : 8;
        bool const real_in = (d.kind == B200_R2C) ? not backward : (d.kind != B200_C2C);
        bool const real_out = (d.kind == B200_R2C) ? backward : (d.kind != B200_C2C);
        b200_line_geom const &gi = backward ? d.out : d.in, &go = backward ? d.in : d.out;
        in = static_cast<const char*>(in) + b_begin * gi.stride_b * static_cast<long long>(real_in ? rsize : 2 * rsize);
        out = static_cast<char*>(out) + b_begin * go.stride_b * static_cast<long long>(real_out ? rsize : 2 * rsize);
        nlines = d.count_a * b_count;
    }
    if (nlines == 0) return B200_SUCCESS;

    if (plan.family == family_contig_real or plan.family == family_strided_real){
        // contiguous lines: the real line doubles as a line of complex numbers on the r2c load and the c2r store, so the
        // pairs must be aligned (the strided kernel moves the reals one by one)
        size_t const csize = is_float ? 8 : 16;
        b200_line_geom const &rg = d.in;           // geometry of the real side of an r2c plan
        bool aligned = true;
        if (d.kind == B200_R2C and plan.family == family_contig_real){
            const void *real_side = backward ? static_cast<const void*>(out) : in;
            if (backward and scatter != nullptr) real_side = nullptr;      // scattered reals are stored one by one
            aligned = (reinterpret_cast<uintptr_t>(real_side) % csize == 0) and (rg.stride_a % 2 == 0) and (rg.stride_b % 2 == 0);
        }
        if (aligned){
            fft_args a;
            a.in = in; a.out = out;
            a.twiddle = static_cast<const char*>(twiddle) + csize * static_cast<size_t>(plan.table_main + plan.table_extra);
            a.twiddle2 = static_cast<const char*>(twiddle) + csize * static_cast<size_t>(plan.table_main);
            a.ig = to_geom(backward ? d.out : d.in);
            a.og = to_geom(backward ? d.in : d.out);
            a.nlines = nlines;
            a.count_a = static_cast<int>(d.count_a);
            a.backward = backward ? 1 : 0;
            a.scale = scale;
            a.smap = static_cast<const scatter_map*>(scatter);
            set_steps(a, steps);
            a.done = nullptr; a.done_mode = 0; a.done_need = 0; a.order_nb = 0; a.multiplier = nullptr;
            a.twiddle0 = twiddle;         // W_n^k, k < n: the full-length engine of the second-generation kernels
            return L.run_real(plan.family == family_strided_real, is_float, scatter != nullptr, plan.real_kind, static_cast<int>(d.n / 2), a);
        }
    }else if (plan.family != family_generic){
        fft_args a{};
        a.in = in; a.out = out; a.twiddle = twiddle; a.twiddle2 = nullptr;
        a.ig = to_geom(backward ? d.out : d.in);   // backward swaps the roles of the two geometries
        a.og = to_geom(backward ? d.in : d.out);
        a.nlines = nlines;
        a.count_a = static_cast<int>(d.count_a);
        a.backward = backward ? 1 : 0;
        a.scale = scale;
        a.smap = static_cast<const scatter_map*>(scatter);
        set_steps(a, steps);
        set_hooks(a, steps);
        L.max_blocks = steps.max_blocks;
        return L.run_pow2(plan.family == family_strided, is_float, scatter != nullptr, static_cast<int>(d.n), a);
    }

    generic_args g;
    g.in = in; g.out = out; g.twiddle = twiddle;
    g.ig = to_geom(backward ? d.out : d.in);
    g.og = to_geom(backward ? d.in : d.out);
    g.nlines = nlines;
    g.count_a = static_cast<int>(d.count_a);
    g.backward = backward ? 1 : 0;
    g.scale = scale;
    g.n = static_cast<int>(d.n);
    g.m = plan.m;
    g.lpb = plan.lpb;
    g.lines_fast = plan.lines_fast;
    g.nfactors = plan.nfactors;
    g.smap = static_cast<const scatter_map*>(scatter);
    set_steps(g, steps);
    for(int i=0; i<24; i++) g.factors[i] = (i < plan.nfactors) ? plan.factors[i] : 1;
    switch(d.kind){
        case B200_C2C:  g.mode = mode_c2c; break;
        case B200_R2C:  g.mode = backward ? mode_c2r : mode_r2c; break;
        case B200_COS:  g.mode = backward ? mode_dct3 : mode_dct2; break;
        case B200_SIN:  g.mode = backward ? mode_dst3 : mode_dst2; break;
        default:        g.mode = mode_dct1; break;
    }
    long long const blocks = (nlines + plan.lpb - 1) / plan.lpb;
    return L.run_generic(is_float, blocks, plan.threads, plan.smem, g);
}

} // namespace b200
