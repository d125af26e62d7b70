// Kernel selection for the batched 1-D FFT executors: maps (precision, length, access pattern) to one
// instantiation of the kernels in fft_device.cuh.  The Launcher policy is the only thing that differs between the
// product (CUDA launch on a stream) and the CPU emulation used by tests/emul.
#pragma once

#include "fft_device.cuh"
#include <cstdlib>

namespace b200 {

// tile shape of the strided kernel: a row of the tile is LPB adjacent lines = 128 bytes when possible
template<typename T> struct row_lines { static constexpr int value = 128 / (2 * sizeof(T)); }; // 8 (fp64) / 16 (fp32)

constexpr bool is_pow2(long long n){ return n > 0 && (n & (n - 1)) == 0; }

// largest power-of-two length handled by the register/shared-memory kernels
constexpr int pow2_max = 4096;
constexpr int pow2_min = 16;

// lengths with odd factors served by the register/shared-memory kernels (radices 3, 5, 6, 7, 10, 12 next to 2, 4, 8, 16)
constexpr bool is_mixed_fast_length(long long n){
    return n == 48 || n == 96 || n == 192 || n == 384 || n == 768 || n == 1536 ||
           n == 80 || n == 160 || n == 320 || n == 640 || n == 1280 ||
           n == 100 || n == 200 || n == 400 || n == 500 || n == 1000 || n == 2000 ||
           n == 112 || n == 224 || n == 448 || n == 896 || n == 1792 || n == 3584;
}
// real lengths 2 m whose half m has a schedule in the real-data tables (48 and 100 have none there)
// (nor for 3584: a real line of 7168 points is beyond the generic kernel that backs up unaligned lines)
constexpr bool is_mixed_fast_half(long long m){ return is_mixed_fast_length(m) && m != 48 && m != 100 && m != 3584; }
constexpr bool is_fast_length(long long n){ return (is_pow2(n) && n >= pow2_min && n <= pow2_max) || is_mixed_fast_length(n); }

inline int strided_512_variant(){
    static int const v = []{ const char *e = std::getenv("HEFFTE_B200_STRIDED_512"); return (e != nullptr) ? std::atoi(e) : 0; }();
    return v;
}
inline int strided_big_variant(bool scatter){
    static int const v = []{ const char *e = std::getenv("HEFFTE_B200_STRIDED_BIG"); return (e != nullptr) ? std::atoi(e) : -1; }();
    return (v >= 0) ? v : (scatter ? 1 : 0);
}
template<typename T, typename RL, int TPL, int LPB, int MINB, bool SCATTER, typename Launcher>
int launch_strided(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    size_t smem = sizeof(cplx<T>) * (size_t)RL::N * LPB + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_strided_kernel<T, RL, TPL, LPB, MINB, true, SCATTER>, blocks, TPL * LPB, smem, a);
    return L.launch(fft_strided_kernel<T, RL, TPL, LPB, MINB, false, SCATTER>, blocks, TPL * LPB, smem, a);
}
template<typename T, typename RL, int LPB, int MINB, bool SCATTER, typename Launcher>
int launch_contig(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    constexpr int PITCH = pad_index(RL::N) + 1;
    size_t smem = ((sizeof(cplx<T>) * (size_t)PITCH * LPB + 15) / 16) * 16 + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_contig_kernel<T, RL, LPB, MINB, true, SCATTER>, blocks, (RL::N / RL::rmax) * LPB, smem, a);
    return L.launch(fft_contig_kernel<T, RL, LPB, MINB, false, SCATTER>, blocks, (RL::N / RL::rmax) * LPB, smem, a);
}

// contiguous kernel with an explicit number of threads per line (schedules whose radices do not all divide the largest one)
template<typename T, typename RL, int TPL, int LPB, int MINB, bool SCATTER, typename Launcher>
int launch_contig_tpl(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    constexpr int PITCH = pad_index(RL::N) + 1;
    size_t smem = ((sizeof(cplx<T>) * (size_t)PITCH * LPB + 15) / 16) * 16 + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_contig_kernel<T, RL, LPB, MINB, true, SCATTER, TPL>, blocks, TPL * LPB, smem, a);
    return L.launch(fft_contig_kernel<T, RL, LPB, MINB, false, SCATTER, TPL>, blocks, TPL * LPB, smem, a);
}

#ifndef B200_HOST_EMULATION
// When the TMA-loaded tile is taken: rows at least 1 MiB apart in double precision (measured: +3.5 % on the slow axis of 512^3,
// -2 % on the middle axis, tools/kbench_tma.cu).  HEFFTE_B200_TMA = 0: never; = force: whenever the box allows (tests).
template<typename T>
bool tma_tile_wanted(fft_args const &a){
    static int const mode = []{ const char *e = std::getenv("HEFFTE_B200_TMA"); return (e == nullptr) ? 1 : ((e[0] == '0') ? 0 : ((e[0] == 'f') ? 2 : 1)); }();
    if (mode == 0 or a.done_mode != 0 or a.order_nb > 1) return false;
    if (mode == 2) return true;
    return sizeof(T) == 8 and a.ig.stride * static_cast<long long>(sizeof(cplx<T>)) >= (1LL << 20);
}
// -1: the box of the stage does not suit a tensor map (the caller keeps the cp.async kernel)
template<typename T, typename RL, int TPL, int LPB, int MINB, typename Launcher>
int launch_strided_tma(fft_args const &a, Launcher &L){
    if (a.count_a <= 0 or a.nlines <= 0 or a.count_a % LPB != 0 or a.nlines % a.count_a != 0) return -1;
    if (a.ig.stride_a != 1 or a.og.stride_a != 1) return -1;
    long long const count_b = a.nlines / a.count_a;
    tma_tile_map map;
    if (not encode_tile_map(map, a.in, static_cast<int>(sizeof(T)), a.count_a, RL::N, a.ig.stride, count_b, a.ig.stride_b, L.batch, a.in_step, LPB)) return -1;
    long long const blocks = a.nlines / LPB;
    size_t const smem = sizeof(cplx<T>) * (size_t)RL::N * LPB + 16;
    if (a.backward) return L.launch_tma(fft_strided_tma_kernel<T, RL, TPL, LPB, MINB, true>, blocks, TPL * LPB, smem, a, map);
    return L.launch_tma(fft_strided_tma_kernel<T, RL, TPL, LPB, MINB, false>, blocks, TPL * LPB, smem, a, map);
}
#endif

// M = lines-per-row multiplier: 1 for double (8 lines = 128 B), 2 for float (16 lines = 128 B)
template<typename T, bool SCATTER, typename Launcher>
int dispatch_strided(int n, fft_args const &a, Launcher &L){
    constexpr int M = row_lines<T>::value / 8;
    switch(n){
        case 16:   return launch_strided<T, radix_list<4, 4, 1, 1>,   4 / M, 32 * M, 2, SCATTER>(a, L);
        case 32:   return launch_strided<T, radix_list<8, 4, 1, 1>,   4 / M, 32 * M, 2, SCATTER>(a, L);
        case 64:   return launch_strided<T, radix_list<8, 8, 1, 1>,   8 / M, 16 * M, 2, SCATTER>(a, L);
        case 128:  return launch_strided<T, radix_list<8, 4, 4, 1>,  16 / M,  8 * M, 2, SCATTER>(a, L);
        case 256:  return launch_strided<T, radix_list<8, 8, 4, 1>,  32 / M,  8 * M, 2, SCATTER>(a, L);
        case 512:
#ifndef B200_HOST_EMULATION
            // rows far apart (the slow axis of a large box): the tile comes by TMA (fft_strided_tma_kernel)
            if constexpr (!SCATTER){
                if (tma_tile_wanted<T>(a)){
                    int const rc = launch_strided_tma<T, radix_list<8, 8, 8, 1>, 32 / M, 8 * M, 3>(a, L);
                    if (rc != -1) return rc;
                }
            }
#endif
            // developer knob HEFFTE_B200_STRIDED_512 = 1: twice the threads on the tile, two CTAs per SM
            if (SCATTER && strided_512_variant() == 1) return launch_strided<T, radix_list<8, 8, 8, 1>, 64 / M, 8 * M, 2, SCATTER>(a, L);
            return launch_strided<T, radix_list<8, 8, 8, 1>,  32 / M,  8 * M, 3, SCATTER>(a, L);
        // 1024 points and more: the tile takes 128 KB, ONE CTA per SM.  Against NVLink what counts is the number of stores in
        // flight: 1024 (fp32) / 512 (fp64) threads on the tile instead of 256 took the fused stages of 1024^3 fp32 on 2 GPUs from
        // 268 to 706 GB/s (profiles/r02_multi_2gpu/bench_c2c_f32_1024_tile_variant_*.log); half tiles with three CTAs per SM reach
        // 442.  Developer knob HEFFTE_B200_STRIDED_BIG = 0 (few threads) | 1 (many) | 2 (half tiles, 1024 only).
        case 1024:
            switch(strided_big_variant(SCATTER)){
                case 1:  return launch_strided<T, radix_list<16, 8, 8, 1>, 64,      8 * M, 1, SCATTER>(a, L);
                case 2:  return launch_strided<T, radix_list<16, 8, 8, 1>, 64 / M,  4 * M, 3, SCATTER>(a, L);
                default: return launch_strided<T, radix_list<16, 8, 8, 1>, 32 / M,  8 * M, 1, SCATTER>(a, L);
            }
        case 2048:
            if (strided_big_variant(SCATTER) == 1) return launch_strided<T, radix_list<8, 8, 8, 4>, 256 / M, 4 * M, 1, SCATTER>(a, L);
            return launch_strided<T, radix_list<8, 8, 8, 4>, 128 / M,  4 * M, 1, SCATTER>(a, L);
        case 4096:
            if (strided_big_variant(SCATTER) == 1) return launch_strided<T, radix_list<8, 8, 8, 8>, 512 / M, 2 * M, 1, SCATTER>(a, L);
            return launch_strided<T, radix_list<8, 8, 8, 8>, 256 / M,  2 * M, 1, SCATTER>(a, L);
        // lengths with factors 3 and 5: TPL divides N / R for every radix R of the schedule; rows of the tile stay 128 bytes
        case 48:   return launch_strided<T, radix_list<4, 4, 3, 1>,     4, 32 * M, 2, SCATTER>(a, L);
        case 96:   return launch_strided<T, radix_list<12, 8, 1, 1>,    4, 16 * M, 2, SCATTER>(a, L);
        case 192:  return launch_strided<T, radix_list<8, 8, 3, 1>,     8,  8 * M, 2, SCATTER>(a, L);
        case 384:  return launch_strided<T, radix_list<8, 8, 6, 1>,    16,  8 * M, 2, SCATTER>(a, L);
        case 768:  return launch_strided<T, radix_list<8, 8, 12, 1>,   32,  8 * M, 1, SCATTER>(a, L);
        case 1536: return launch_strided<T, radix_list<8, 8, 8, 3>,    64,  4 * M, 1, SCATTER>(a, L);
        case 80:   return launch_strided<T, radix_list<5, 4, 4, 1>,     4, 32 * M, 2, SCATTER>(a, L);
        case 160:  return launch_strided<T, radix_list<10, 4, 4, 1>,    8, 16 * M, 2, SCATTER>(a, L);
        case 320:  return launch_strided<T, radix_list<5, 4, 4, 4>,    16,  8 * M, 2, SCATTER>(a, L);
        case 640:  return launch_strided<T, radix_list<10, 4, 4, 4>,   32,  8 * M, 2, SCATTER>(a, L);
        case 1280: return launch_strided<T, radix_list<10, 8, 4, 4>,   32,  4 * M, 1, SCATTER>(a, L);
        case 100:  return launch_strided<T, radix_list<10, 10, 1, 1>,  10, 16 * M, 2, SCATTER>(a, L);
        case 200:  return launch_strided<T, radix_list<10, 10, 2, 1>,  10,  8 * M, 2, SCATTER>(a, L);
        case 400:  return launch_strided<T, radix_list<10, 10, 4, 1>,  20,  8 * M, 2, SCATTER>(a, L);
        case 500:  return launch_strided<T, radix_list<10, 10, 5, 1>,  25,  8 * M, 2, SCATTER>(a, L);
        case 1000: return launch_strided<T, radix_list<10, 10, 10, 1>, 50,  4 * M, 1, SCATTER>(a, L);
        case 2000: return launch_strided<T, radix_list<10, 10, 10, 2>, 100, 2 * M, 1, SCATTER>(a, L);
        // 7 * 2^k
        case 112:  return launch_strided<T, radix_list<7, 4, 4, 1>,     4, 16 * M, 2, SCATTER>(a, L);
        case 224:  return launch_strided<T, radix_list<7, 8, 4, 1>,     4,  8 * M, 2, SCATTER>(a, L);
        case 448:  return launch_strided<T, radix_list<7, 8, 8, 1>,     8,  8 * M, 2, SCATTER>(a, L);
        case 896:  return launch_strided<T, radix_list<7, 8, 16, 1>,    8,  8 * M, 1, SCATTER>(a, L);
        case 1792: return launch_strided<T, radix_list<7, 16, 16, 1>,  16,  4 * M, 1, SCATTER>(a, L);
        case 3584: return launch_strided<T, radix_list<7, 8, 8, 8>,    64,  2 * M, 1, SCATTER>(a, L);
        default: return -1;
    }
}

// fused spectral operator along a strided axis (fft_strided_conv_kernel): the tile shapes of the strided kernel
template<typename T, typename RL, int TPL, int LPB, int MINB, bool SCATTER, typename Launcher>
int launch_strided_conv(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    size_t smem = sizeof(cplx<T>) * (size_t)RL::N * LPB + (SCATTER ? sizeof(scatter_map) : 0);
    return L.launch(fft_strided_conv_kernel<T, RL, TPL, LPB, MINB, SCATTER>, blocks, TPL * LPB, smem, a);
}
constexpr bool is_conv_length(long long n){ return is_pow2(n) && n >= pow2_min && n <= pow2_max; }
template<typename T, bool SCATTER, typename Launcher>
int dispatch_strided_conv(int n, fft_args const &a, Launcher &L){
    constexpr int M = row_lines<T>::value / 8;
    switch(n){
        case 16:   return launch_strided_conv<T, radix_list<4, 4, 1, 1>,   4 / M, 32 * M, 2, SCATTER>(a, L);
        case 32:   return launch_strided_conv<T, radix_list<8, 4, 1, 1>,   4 / M, 32 * M, 2, SCATTER>(a, L);
        case 64:   return launch_strided_conv<T, radix_list<8, 8, 1, 1>,   8 / M, 16 * M, 2, SCATTER>(a, L);
        case 128:  return launch_strided_conv<T, radix_list<8, 4, 4, 1>,  16 / M,  8 * M, 2, SCATTER>(a, L);
        case 256:  return launch_strided_conv<T, radix_list<8, 8, 4, 1>,  32 / M,  8 * M, 2, SCATTER>(a, L);
        case 512:  return launch_strided_conv<T, radix_list<8, 8, 8, 1>,  32 / M,  8 * M, 3, SCATTER>(a, L);
        case 1024:
            if (strided_big_variant(SCATTER) == 1) return launch_strided_conv<T, radix_list<16, 8, 8, 1>, 64, 8 * M, 1, SCATTER>(a, L);
            return launch_strided_conv<T, radix_list<16, 8, 8, 1>, 32 / M,  8 * M, 1, SCATTER>(a, L);
        case 2048: return launch_strided_conv<T, radix_list<8, 8, 8, 4>, 128 / M,  4 * M, 1, SCATTER>(a, L);
        case 4096: return launch_strided_conv<T, radix_list<8, 8, 8, 8>, 256 / M,  2 * M, 1, SCATTER>(a, L);
        default: return -1;
    }
}

template<typename T, bool SCATTER, typename Launcher>
int dispatch_contig(int n, fft_args const &a, Launcher &L){
    // small CTAs (64-128 threads), many per SM: measured best on B200 (tools/kbench.cu, profiles/)
    switch(n){
        case 16:   return launch_contig<T, radix_list<4, 4, 1, 1>,   32, 8, SCATTER>(a, L);
        case 32:   return launch_contig<T, radix_list<8, 4, 1, 1>,   32, 6, SCATTER>(a, L);
        case 64:   return launch_contig<T, radix_list<8, 8, 1, 1>,   16, 6, SCATTER>(a, L);
        case 128:  return launch_contig<T, radix_list<8, 4, 4, 1>,    8, 6, SCATTER>(a, L);
        // tools/kbench_c2c.cu on B200 (profiles/r02_single/kbench_variants_a.log): 256-point fp32 <16,16> 6.2 vs 5.0 TB/s for <8,8,4>;
        // 512-point fp64 <4,8,16> with two lines per CTA 6.95 vs 6.25 TB/s for <8,8,8> with one
        case 256:  return launch_contig<T, radix_list<16, 16, 1, 1>,  4, 6, SCATTER>(a, L);
        case 512:  return launch_contig<T, radix_list<4, 8, 16, 1>,   2, 8, SCATTER>(a, L);
        // long lines with a fused reshape: twice the CTAs per SM keep more remote stores in flight (see dispatch_strided)
        case 1024:
            if (strided_big_variant(SCATTER) == 1) return launch_contig<T, radix_list<16, 8, 8, 1>, 1, 8, SCATTER>(a, L);
            return launch_contig<T, radix_list<16, 8, 8, 1>,   1, 4, SCATTER>(a, L);
        case 2048:
            if (strided_big_variant(SCATTER) == 1) return launch_contig<T, radix_list<8, 8, 8, 4>, 1, 4, SCATTER>(a, L);
            return launch_contig<T, radix_list<8, 8, 8, 4>,    1, 2, SCATTER>(a, L);
        case 4096:
            if (strided_big_variant(SCATTER) == 1) return launch_contig<T, radix_list<8, 8, 8, 8>, 1, 2, SCATTER>(a, L);
            return launch_contig<T, radix_list<8, 8, 8, 8>,    1, 1, SCATTER>(a, L);
        // lengths with factors 3 and 5 (threads per line given explicitly; a thread holds N / TPL values between two passes)
        case 48:   return launch_contig_tpl<T, radix_list<4, 4, 3, 1>,     4, 16, 4, SCATTER>(a, L);
        case 96:   return launch_contig_tpl<T, radix_list<12, 8, 1, 1>,    4, 16, 4, SCATTER>(a, L);
        case 192:  return launch_contig_tpl<T, radix_list<8, 8, 3, 1>,     8,  8, 4, SCATTER>(a, L);
        case 384:  return launch_contig_tpl<T, radix_list<8, 8, 6, 1>,    16,  4, 4, SCATTER>(a, L);
        case 768:  return launch_contig_tpl<T, radix_list<8, 8, 12, 1>,   32,  2, 4, SCATTER>(a, L);
        case 1536: return launch_contig_tpl<T, radix_list<8, 8, 8, 3>,    64,  1, 4, SCATTER>(a, L);
        case 80:   return launch_contig_tpl<T, radix_list<5, 4, 4, 1>,     4, 16, 4, SCATTER>(a, L);
        case 160:  return launch_contig_tpl<T, radix_list<10, 4, 4, 1>,    8,  8, 4, SCATTER>(a, L);
        case 320:  return launch_contig_tpl<T, radix_list<5, 4, 4, 4>,    16,  4, 4, SCATTER>(a, L);
        case 640:  return launch_contig_tpl<T, radix_list<10, 4, 4, 4>,   32,  2, 4, SCATTER>(a, L);
        case 1280: return launch_contig_tpl<T, radix_list<10, 8, 4, 4>,   32,  2, 2, SCATTER>(a, L);
        case 100:  return launch_contig_tpl<T, radix_list<10, 10, 1, 1>,  10,  8, 4, SCATTER>(a, L);
        case 200:  return launch_contig_tpl<T, radix_list<10, 10, 2, 1>,  10,  8, 4, SCATTER>(a, L);
        case 400:  return launch_contig_tpl<T, radix_list<10, 10, 4, 1>,  20,  4, 4, SCATTER>(a, L);
        case 500:  return launch_contig_tpl<T, radix_list<10, 10, 5, 1>,  25,  4, 4, SCATTER>(a, L);
        case 1000: return launch_contig_tpl<T, radix_list<10, 10, 10, 1>, 50,  2, 2, SCATTER>(a, L);
        case 2000: return launch_contig_tpl<T, radix_list<10, 10, 10, 2>, 100, 1, 2, SCATTER>(a, L);
        case 112:  return launch_contig_tpl<T, radix_list<7, 4, 4, 1>,     4, 16, 4, SCATTER>(a, L);
        case 224:  return launch_contig_tpl<T, radix_list<7, 8, 4, 1>,     4,  8, 4, SCATTER>(a, L);
        case 448:  return launch_contig_tpl<T, radix_list<7, 8, 8, 1>,     8,  8, 4, SCATTER>(a, L);
        case 896:  return launch_contig_tpl<T, radix_list<7, 8, 16, 1>,    8,  4, 4, SCATTER>(a, L);
        case 1792: return launch_contig_tpl<T, radix_list<7, 16, 16, 1>,  16,  2, 2, SCATTER>(a, L);
        case 3584: return launch_contig_tpl<T, radix_list<7, 8, 8, 8>,    64,  1, 2, SCATTER>(a, L);
        default: return -1;
    }
}

// ---- paired kernel (fft_pair_kernel): the contiguous-axis and the middle-axis transform of a box in one persistent launch ------
// Shapes per length: both phases use the same number of threads; the larger tile decides the shared memory.
// tools/kbench_pair.cu on B200: 256^3 fp32 0.097 ms against 0.119 ms for the two launches; 512^3 fp64 1.55 against 1.27 ms --
// the L2 pays for the small planes only, so the plan uses the pair where the second transform is bound by NVLink anyway (the
// local pass in front of a fused reshape hides behind the transfer) and, without a reshape, for single precision.
template<typename T, int N> struct pair_shape;
template<> struct pair_shape<double, 128>  { using A = radix_list<8, 4, 4, 1>;  using B = radix_list<8, 4, 4, 1>;  static constexpr int TPLA = 16, LPBA = 8,  TPLB = 16, LPBB = 8,  MINB = 4; };
template<> struct pair_shape<double, 256>  { using A = radix_list<8, 8, 4, 1>;  using B = radix_list<8, 8, 4, 1>;  static constexpr int TPLA = 32, LPBA = 8,  TPLB = 32, LPBB = 8,  MINB = 3; };
template<> struct pair_shape<double, 512>  { using A = radix_list<8, 8, 8, 1>;  using B = radix_list<8, 8, 8, 1>;  static constexpr int TPLA = 64, LPBA = 4,  TPLB = 32, LPBB = 8,  MINB = 3; };
template<> struct pair_shape<double, 1024> { using A = radix_list<16, 8, 8, 1>; using B = radix_list<16, 8, 8, 1>; static constexpr int TPLA = 64, LPBA = 4,  TPLB = 32, LPBB = 8,  MINB = 1; };
template<> struct pair_shape<float, 128>   { using A = radix_list<8, 4, 4, 1>;  using B = radix_list<8, 4, 4, 1>;  static constexpr int TPLA = 16, LPBA = 8,  TPLB = 8,  LPBB = 16, MINB = 4; };
template<> struct pair_shape<float, 256>   { using A = radix_list<16, 16, 1, 1>; using B = radix_list<8, 8, 4, 1>; static constexpr int TPLA = 16, LPBA = 16, TPLB = 16, LPBB = 16, MINB = 4; };
template<> struct pair_shape<float, 512>   { using A = radix_list<8, 8, 8, 1>;  using B = radix_list<8, 8, 8, 1>;  static constexpr int TPLA = 64, LPBA = 4,  TPLB = 16, LPBB = 16, MINB = 3; };
template<> struct pair_shape<float, 1024>  { using A = radix_list<16, 8, 8, 1>; using B = radix_list<16, 8, 8, 1>; static constexpr int TPLA = 64, LPBA = 4,  TPLB = 16, LPBB = 16, MINB = 1; };

constexpr bool is_pair_length(long long n){ return n == 128 || n == 256 || n == 512 || n == 1024; }

// L.launch_pair(kernel, threads, smem, args) sizes the grid to what is resident at once (one CTA in the emulation)
template<typename T, int N, bool SCATTER, typename Launcher>
int launch_pair_n(bool contig_first, pair_args p, Launcher &L){
    using S = pair_shape<T, N>;
    constexpr int threads = S::TPLB * S::LPBB;
    static_assert(S::TPLA * S::LPBA == threads, "pair shapes must agree on the block size");
    size_t const smem = pair_smem_bytes<T, typename S::A, S::LPBA, typename S::B, S::LPBB, SCATTER>();
    if (p.a.count_a % S::LPBA != 0 or p.b.count_a % S::LPBB != 0) return -1;     // tiles must not straddle planes
    p.tiles_a = static_cast<unsigned>(p.a.count_a / S::LPBA);
    p.tiles_b = static_cast<unsigned>(p.b.count_a / S::LPBB);
    bool const bwd = p.a.backward != 0;
    if (contig_first){
        if (bwd) return L.launch_pair(fft_pair_kernel<T, typename S::A, S::LPBA, S::TPLA, typename S::B, S::TPLB, S::LPBB, S::MINB, true, SCATTER, true>, threads, smem, p);
        return L.launch_pair(fft_pair_kernel<T, typename S::A, S::LPBA, S::TPLA, typename S::B, S::TPLB, S::LPBB, S::MINB, false, SCATTER, true>, threads, smem, p);
    }
    if (bwd) return L.launch_pair(fft_pair_kernel<T, typename S::A, S::LPBA, S::TPLA, typename S::B, S::TPLB, S::LPBB, S::MINB, true, SCATTER, false>, threads, smem, p);
    return L.launch_pair(fft_pair_kernel<T, typename S::A, S::LPBA, S::TPLA, typename S::B, S::TPLB, S::LPBB, S::MINB, false, SCATTER, false>, threads, smem, p);
}
template<typename T, bool SCATTER, typename Launcher>
int dispatch_pair(int n, bool contig_first, pair_args const &p, Launcher &L){
    switch(n){
        case 128:  return launch_pair_n<T, 128, SCATTER>(contig_first, p, L);
        case 256:  return launch_pair_n<T, 256, SCATTER>(contig_first, p, L);
        case 512:  return launch_pair_n<T, 512, SCATTER>(contig_first, p, L);
        case 1024: return launch_pair_n<T, 1024, SCATTER>(contig_first, p, L);
        default: return -1;
    }
}

// real-data variants (fft_contig_real_kernel): m = n/2 is the length of the complex engine
template<typename T, typename RL, int LPB, int MINB, int KIND, bool SCATTER, typename Launcher>
int launch_contig_real(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    constexpr int PITCH = pad_index(RL::N) + 1;
    size_t smem = ((sizeof(cplx<T>) * (size_t)PITCH * LPB + 15) / 16) * 16 + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_contig_real_kernel<T, RL, LPB, MINB, KIND, true, SCATTER>, blocks, (RL::N / RL::rmax) * LPB, smem, a);
    return L.launch(fft_contig_real_kernel<T, RL, LPB, MINB, KIND, false, SCATTER>, blocks, (RL::N / RL::rmax) * LPB, smem, a);
}

constexpr int real_pow2_min = 2 * pow2_min;   // real lengths served by the fast real kernels
constexpr int real_pow2_max = 4096;
// real lengths n = 2m served by the real-data kernels: powers of two 32 ... 4096 and twice the mixed c2c lengths
constexpr bool is_fast_real_length(long long n){
    return (is_pow2(n) && n >= real_pow2_min && n <= real_pow2_max) || (n % 2 == 0 && is_mixed_fast_half(n / 2));
}

template<typename T, typename RL, int TPL, int LPB, int MINB, int KIND, bool SCATTER, typename Launcher>
int launch_contig_real_tpl(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    constexpr int PITCH = pad_index(RL::N) + 1;
    size_t smem = ((sizeof(cplx<T>) * (size_t)PITCH * LPB + 15) / 16) * 16 + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_contig_real_kernel<T, RL, LPB, MINB, KIND, true, SCATTER, TPL>, blocks, TPL * LPB, smem, a);
    return L.launch(fft_contig_real_kernel<T, RL, LPB, MINB, KIND, false, SCATTER, TPL>, blocks, TPL * LPB, smem, a);
}

template<typename T, int KIND, bool SCATTER, typename Launcher>
int dispatch_contig_real_kind(int m, fft_args const &a, Launcher &L){
    switch(m){
        case 16:   return launch_contig_real<T, radix_list<4, 4, 1, 1>,   32, 6, KIND, SCATTER>(a, L);
        case 32:   return launch_contig_real<T, radix_list<8, 4, 1, 1>,   32, 4, KIND, SCATTER>(a, L);
        case 64:   return launch_contig_real<T, radix_list<8, 8, 1, 1>,   16, 4, KIND, SCATTER>(a, L);
        case 128:  return launch_contig_real<T, radix_list<8, 4, 4, 1>,    8, 4, KIND, SCATTER>(a, L);
        // two radix-16 passes: one shared-memory exchange less than <8,8,4> (tools/kbench_real.cu on B200, 512-point fp64 lines:
        // r2c 5.6 vs 4.6 TB/s, c2r 4.7 vs 4.0, DCT-II 4.4 vs 3.8, DCT-III 3.7 vs 3.3)
        case 256:  return launch_contig_real<T, radix_list<16, 16, 1, 1>,  4, 6, KIND, SCATTER>(a, L);
        case 512:  return launch_contig_real<T, radix_list<8, 8, 8, 1>,    1, 12, KIND, SCATTER>(a, L);
        case 1024: return launch_contig_real<T, radix_list<16, 8, 8, 1>,   1, 4, KIND, SCATTER>(a, L);
        case 2048: return launch_contig_real<T, radix_list<8, 8, 8, 4>,    1, 2, KIND, SCATTER>(a, L);
        // m with factors 3 and 5 (real lengths 160 ... 4000); the first radix is even: the sine kinds flip the upper half there
        case 96:   return launch_contig_real_tpl<T, radix_list<12, 8, 1, 1>,    4, 16, 4, KIND, SCATTER>(a, L);
        case 192:  return launch_contig_real_tpl<T, radix_list<8, 8, 3, 1>,     8,  8, 4, KIND, SCATTER>(a, L);
        case 384:  return launch_contig_real_tpl<T, radix_list<8, 8, 6, 1>,    16,  4, 4, KIND, SCATTER>(a, L);
        case 768:  return launch_contig_real_tpl<T, radix_list<8, 8, 12, 1>,   32,  2, 4, KIND, SCATTER>(a, L);
        case 1536: return launch_contig_real_tpl<T, radix_list<8, 8, 8, 3>,    64,  1, 4, KIND, SCATTER>(a, L);
        case 80:   return launch_contig_real_tpl<T, radix_list<4, 4, 5, 1>,     4, 16, 4, KIND, SCATTER>(a, L);
        case 160:  return launch_contig_real_tpl<T, radix_list<10, 4, 4, 1>,    8,  8, 4, KIND, SCATTER>(a, L);
        case 320:  return launch_contig_real_tpl<T, radix_list<4, 4, 4, 5>,    16,  4, 4, KIND, SCATTER>(a, L);
        case 640:  return launch_contig_real_tpl<T, radix_list<10, 4, 4, 4>,   32,  2, 4, KIND, SCATTER>(a, L);
        case 1280: return launch_contig_real_tpl<T, radix_list<10, 8, 4, 4>,   32,  2, 2, KIND, SCATTER>(a, L);
        case 200:  return launch_contig_real_tpl<T, radix_list<10, 10, 2, 1>,  10,  8, 4, KIND, SCATTER>(a, L);
        case 400:  return launch_contig_real_tpl<T, radix_list<10, 10, 4, 1>,  20,  4, 4, KIND, SCATTER>(a, L);
        case 500:  return launch_contig_real_tpl<T, radix_list<10, 10, 5, 1>,  25,  4, 4, KIND, SCATTER>(a, L);
        case 1000: return launch_contig_real_tpl<T, radix_list<10, 10, 10, 1>, 50,  2, 2, KIND, SCATTER>(a, L);
        case 2000: return launch_contig_real_tpl<T, radix_list<10, 10, 10, 2>, 100, 1, 2, KIND, SCATTER>(a, L);
        // m = 7 * 2^k (real lengths 224 ... 7168); the radix 7 comes last: the first radix stays even for the sine kinds
        case 112:  return launch_contig_real_tpl<T, radix_list<4, 4, 7, 1>,     4, 16, 4, KIND, SCATTER>(a, L);
        case 224:  return launch_contig_real_tpl<T, radix_list<8, 4, 7, 1>,     4,  8, 4, KIND, SCATTER>(a, L);
        case 448:  return launch_contig_real_tpl<T, radix_list<8, 8, 7, 1>,     8,  8, 4, KIND, SCATTER>(a, L);
        case 896:  return launch_contig_real_tpl<T, radix_list<16, 8, 7, 1>,    8,  4, 4, KIND, SCATTER>(a, L);
        case 1792: return launch_contig_real_tpl<T, radix_list<16, 16, 7, 1>,  16,  2, 2, KIND, SCATTER>(a, L);
        default: return -1;
    }
}
template<typename T, typename RL, int TPL, int LPB, int MINB, int KIND, bool SCATTER, typename Launcher>
int launch_strided_real(fft_args const &a, Launcher &L){
    long long blocks = (a.nlines + LPB - 1) / LPB;
    size_t smem = sizeof(cplx<T>) * (size_t)RL::N * LPB + (SCATTER ? sizeof(scatter_map) : 0);
    if (a.backward) return L.launch(fft_strided_real_kernel<T, RL, TPL, LPB, MINB, KIND, true, SCATTER>, blocks, TPL * LPB, smem, a);
    return L.launch(fft_strided_real_kernel<T, RL, TPL, LPB, MINB, KIND, false, SCATTER>, blocks, TPL * LPB, smem, a);
}
// tile [m][LPB]: a row of LPB adjacent reals is 128 bytes where shared memory allows (fp64: 16 lines, fp32: 32 lines)
template<typename T, int KIND, bool SCATTER, typename Launcher>
int dispatch_strided_real_kind(int m, fft_args const &a, Launcher &L){
    constexpr int F = row_lines<T>::value / 8;    // 1 (fp64) / 2 (fp32)
    switch(m){
        case 16:   return launch_strided_real<T, radix_list<4, 4, 1, 1>,    4 / F, 32 * F, 2, KIND, SCATTER>(a, L);
        case 32:   return launch_strided_real<T, radix_list<8, 4, 1, 1>,    4 / F, 32 * F, 2, KIND, SCATTER>(a, L);
        case 64:   return launch_strided_real<T, radix_list<8, 8, 1, 1>,    8 / F, 16 * F, 2, KIND, SCATTER>(a, L);
        case 128:  return launch_strided_real<T, radix_list<8, 4, 4, 1>,    8 / F, 16 * F, 2, KIND, SCATTER>(a, L);
        case 256:  // forward: two radix-16 passes (4.8 vs 4.6 TB/s, r2c 5.2 vs 4.9); backward: <8,8,4> stays ahead (4.0 TB/s)
            if (a.backward) return launch_strided_real<T, radix_list<8, 8, 4, 1>, 16 / F, 16 * F, 3, KIND, SCATTER>(a, L);
            return launch_strided_real<T, radix_list<16, 16, 1, 1>, 16 / F, 16 * F, 3, KIND, SCATTER>(a, L);
        case 512:  return launch_strided_real<T, radix_list<8, 8, 8, 1>,   32 / F, 16 * F, 1, KIND, SCATTER>(a, L);
        case 1024: return launch_strided_real<T, radix_list<16, 8, 8, 1>,  32 / F,  8 * F, 1, KIND, SCATTER>(a, L);
        case 2048: return launch_strided_real<T, radix_list<8, 8, 8, 4>,  128 / F,  4 * F, 1, KIND, SCATTER>(a, L);
        // m with factors 3 and 5; even first radix (sine kinds); tiles of 40 ... 96 KB
        case 96:   return launch_strided_real<T, radix_list<12, 8, 1, 1>,    4, 32 * F, 2, KIND, SCATTER>(a, L);
        case 192:  return launch_strided_real<T, radix_list<8, 8, 3, 1>,     8, 16 * F, 2, KIND, SCATTER>(a, L);
        case 384:  return launch_strided_real<T, radix_list<8, 8, 6, 1>,    16, 16 * F, 2, KIND, SCATTER>(a, L);
        case 768:  return launch_strided_real<T, radix_list<8, 8, 12, 1>,   32,  8 * F, 1, KIND, SCATTER>(a, L);
        case 1536: return launch_strided_real<T, radix_list<8, 8, 8, 3>,    64,  4 * F, 1, KIND, SCATTER>(a, L);
        case 80:   return launch_strided_real<T, radix_list<4, 4, 5, 1>,     4, 32 * F, 2, KIND, SCATTER>(a, L);
        case 160:  return launch_strided_real<T, radix_list<10, 4, 4, 1>,    8, 16 * F, 2, KIND, SCATTER>(a, L);
        case 320:  return launch_strided_real<T, radix_list<4, 4, 4, 5>,    16, 16 * F, 2, KIND, SCATTER>(a, L);
        case 640:  return launch_strided_real<T, radix_list<10, 4, 4, 4>,   32,  8 * F, 1, KIND, SCATTER>(a, L);
        case 1280: return launch_strided_real<T, radix_list<10, 8, 4, 4>,   32,  4 * F, 1, KIND, SCATTER>(a, L);
        case 200:  return launch_strided_real<T, radix_list<10, 10, 2, 1>,  10, 16 * F, 2, KIND, SCATTER>(a, L);
        case 400:  return launch_strided_real<T, radix_list<10, 10, 4, 1>,  20,  8 * F, 2, KIND, SCATTER>(a, L);
        case 500:  return launch_strided_real<T, radix_list<10, 10, 5, 1>,  25,  8 * F, 1, KIND, SCATTER>(a, L);
        case 1000: return launch_strided_real<T, radix_list<10, 10, 10, 1>, 50,  4 * F, 1, KIND, SCATTER>(a, L);
        case 2000: return launch_strided_real<T, radix_list<10, 10, 10, 2>, 100, 2 * F, 1, KIND, SCATTER>(a, L);
        case 112:  return launch_strided_real<T, radix_list<4, 4, 7, 1>,     4, 32 * F, 2, KIND, SCATTER>(a, L);
        case 224:  return launch_strided_real<T, radix_list<8, 4, 7, 1>,     4, 16 * F, 2, KIND, SCATTER>(a, L);
        case 448:  return launch_strided_real<T, radix_list<8, 8, 7, 1>,     8, 16 * F, 2, KIND, SCATTER>(a, L);
        case 896:  return launch_strided_real<T, radix_list<16, 8, 7, 1>,    8,  8 * F, 1, KIND, SCATTER>(a, L);
        case 1792: return launch_strided_real<T, radix_list<16, 16, 7, 1>,  16,  4 * F, 1, KIND, SCATTER>(a, L);
        default: return -1;
    }
}
// second generation of the strided real kernels (fft_strided_real2_kernel): two adjacent real lines per complex line, the tile
// shapes of the complex strided kernel for the full real length n
template<typename T, typename RL, int TPL, int LPB, int MINB, int KIND, typename Launcher>
int launch_strided_real2(fft_args const &a, Launcher &L){
    long long const pairs = (a.nlines + 1) / 2;
    long long blocks = (pairs + LPB - 1) / LPB;
    size_t smem = sizeof(cplx<T>) * (size_t)RL::N * LPB;
    if (a.backward) return L.launch(fft_strided_real2_kernel<T, RL, TPL, LPB, MINB, KIND, true>, blocks, TPL * LPB, smem, a);
    return L.launch(fft_strided_real2_kernel<T, RL, TPL, LPB, MINB, KIND, false>, blocks, TPL * LPB, smem, a);
}
constexpr bool is_real2_length(long long n){ return is_pow2(n) && n >= 32 && n <= 4096; }
template<typename T, int KIND, typename Launcher>
int dispatch_strided_real2_kind(int n, fft_args const &a, Launcher &L){
    constexpr int M = row_lines<T>::value / 8;
    switch(n){
        case 32:   return launch_strided_real2<T, radix_list<8, 4, 1, 1>,   4 / M, 32 * M, 2, KIND>(a, L);
        case 64:   return launch_strided_real2<T, radix_list<8, 8, 1, 1>,   8 / M, 16 * M, 2, KIND>(a, L);
        case 128:  return launch_strided_real2<T, radix_list<8, 4, 4, 1>,  16 / M,  8 * M, 2, KIND>(a, L);
        case 256:  return launch_strided_real2<T, radix_list<8, 8, 4, 1>,  32 / M,  8 * M, 2, KIND>(a, L);
        case 512:  return launch_strided_real2<T, radix_list<8, 8, 8, 1>,  32 / M,  8 * M, 3, KIND>(a, L);
        case 1024: return launch_strided_real2<T, radix_list<16, 8, 8, 1>, 32 / M,  8 * M, 1, KIND>(a, L);
        case 2048: return launch_strided_real2<T, radix_list<8, 8, 8, 4>, 128 / M,  4 * M, 1, KIND>(a, L);
        case 4096: return launch_strided_real2<T, radix_list<8, 8, 8, 8>, 256 / M,  2 * M, 1, KIND>(a, L);
        default: return -1;
    }
}
template<typename T, typename Launcher>
int dispatch_strided_real2(int kind, int n, fft_args const &a, Launcher &L){
    switch(kind){
        case real_r2c: return dispatch_strided_real2_kind<T, real_r2c>(n, a, L);
        case real_cos: return dispatch_strided_real2_kind<T, real_cos>(n, a, L);
        case real_sin: return dispatch_strided_real2_kind<T, real_sin>(n, a, L);
        default: return -1;
    }
}

// contiguous-axis real transforms, two lines per complex line (fft_contig_real2_kernel): schedules that start and end with the
// same radix R; a line pair is worked on by n / 2R threads
template<typename T, typename RL, int LPB, int MINB, int KIND, typename Launcher>
int launch_contig_real2(fft_args const &a, Launcher &L){
    long long const pairs = a.nlines / 2;
    long long blocks = (pairs + LPB - 1) / LPB;
    constexpr int PITCH = pad_index(RL::N) + 1;
    constexpr int threads = (RL::N / (2 * RL::radix(0))) * LPB;
    size_t smem = ((sizeof(cplx<T>) * (size_t)PITCH * LPB + 15) / 16) * 16;
    if (a.backward) return L.launch(fft_contig_real2_kernel<T, RL, LPB, MINB, KIND, true>, blocks, threads, smem, a);
    return L.launch(fft_contig_real2_kernel<T, RL, LPB, MINB, KIND, false>, blocks, threads, smem, a);
}
template<typename T, int KIND, typename Launcher>
int dispatch_contig_real2_kind(int n, fft_args const &a, Launcher &L){
    switch(n){
        // one warp per CTA, sixteen CTAs per SM (tools/kbench_real.cu on B200, 512-point fp64: r2c 6.24 TB/s against 5.96 for 64
        // threads x 8; profiles/r02_single/kbench_real_second_generation.log)
        case 32:   return launch_contig_real2<T, radix_list<4, 2, 4, 1>,    8, 16, KIND>(a, L);
        case 64:   return launch_contig_real2<T, radix_list<8, 8, 1, 1>,    8, 16, KIND>(a, L);
        case 128:  return launch_contig_real2<T, radix_list<4, 8, 4, 1>,    2, 16, KIND>(a, L);
        case 256:  return launch_contig_real2<T, radix_list<8, 4, 8, 1>,    2, 16, KIND>(a, L);
        case 512:  return launch_contig_real2<T, radix_list<8, 8, 8, 1>,    1, 16, KIND>(a, L);
        case 1024: return launch_contig_real2<T, radix_list<8, 16, 8, 1>,   1, 8, KIND>(a, L);
        case 2048: return launch_contig_real2<T, radix_list<8, 8, 4, 8>,    1, 4, KIND>(a, L);
        case 4096: return launch_contig_real2<T, radix_list<8, 8, 8, 8>,    1, 2, KIND>(a, L);
        default: return -1;
    }
}
template<typename T, typename Launcher>
int dispatch_contig_real2(int kind, int n, fft_args const &a, Launcher &L){
    switch(kind){
        case real_r2c: return dispatch_contig_real2_kind<T, real_r2c>(n, a, L);
        case real_cos: return dispatch_contig_real2_kind<T, real_cos>(n, a, L);
        case real_sin: return dispatch_contig_real2_kind<T, real_sin>(n, a, L);
        default: return -1;
    }
}

template<typename T, bool SCATTER, typename Launcher>
int dispatch_strided_real(int kind, int m, fft_args const &a, Launcher &L){
    switch(kind){
        case real_r2c: return dispatch_strided_real_kind<T, real_r2c, SCATTER>(m, a, L);
        case real_cos: return dispatch_strided_real_kind<T, real_cos, SCATTER>(m, a, L);
        case real_sin: return dispatch_strided_real_kind<T, real_sin, SCATTER>(m, a, L);
        default: return -1;
    }
}

template<typename T, bool SCATTER, typename Launcher>
int dispatch_contig_real(int kind, int m, fft_args const &a, Launcher &L){
    switch(kind){
        case real_r2c: return dispatch_contig_real_kind<T, real_r2c, SCATTER>(m, a, L);
        case real_cos: return dispatch_contig_real_kind<T, real_cos, SCATTER>(m, a, L);
        case real_sin: return dispatch_contig_real_kind<T, real_sin, SCATTER>(m, a, L);
        default: return -1;
    }
}

} // namespace b200
