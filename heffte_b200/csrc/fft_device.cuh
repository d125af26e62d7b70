// Device-side building blocks of the batched 1-D FFT executors (sm_100a).
//
// Replaces the cuFFT calls of the reference's CUDA backend (reference: include/heffte_backend_cuda.h:356-415,
// 494-524, 592-727 -- cufftMakePlanMany + cufftExec{C2C,Z2Z,R2C,D2Z,C2R,Z2D}) with hand-written kernels:
//   * fft_strided_kernel   : lines whose neighbours are adjacent in memory (FFT along the middle / slow axis of a
//                            box).  A CTA owns a tile of LPB adjacent lines, so every global access is a full
//                            128-byte row; in-place decimation-in-frequency passes in shared memory, the digit
//                            reversal is absorbed into the row index of the final store (free, because coalescing
//                            runs across lines, not along the transform).
//   * fft_contig_kernel    : lines contiguous in memory (FFT along the fast axis).  Stockham auto-sort passes so
//                            that both the first global load and the last global store are unit-stride.
//   * fft_generic_kernel   : any length (mixed radix by prime factors, O(N*p) per factor p), used for
//                            non-power-of-two sizes and as the r2c/c2r/r2r engine for such sizes.
// Butterflies are register resident (radix 2/4/8/16); data is exchanged between passes through shared memory.
// The inverse transform reuses the forward code through the identity  ifft(x) = swap(fft(swap(x)))  where
// swap exchanges real and imaginary parts, so direction costs nothing.
#pragma once

#include "cuda_compat.h"
#include "scatter_map.h"
#include <stdint.h>

namespace b200 {

template<typename T> struct cplx_of;
template<> struct cplx_of<float>  { using type = float2; };
template<> struct cplx_of<double> { using type = double2; };
template<typename T> using cplx = typename cplx_of<T>::type;

template<typename T> __device__ __forceinline__ cplx<T> mk(T x, T y){ cplx<T> r; r.x = x; r.y = y; return r; }
template<typename C> __device__ __forceinline__ C cadd(C a, C b){ C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template<typename C> __device__ __forceinline__ C csub(C a, C b){ C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template<typename C> __device__ __forceinline__ C cmul(C a, C b){
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// multiply by -i  (forward quarter turn)
template<typename C> __device__ __forceinline__ C mul_mi(C a){ C r; r.x = a.y; r.y = -a.x; return r; }
template<typename C> __device__ __forceinline__ C cswap(C a){ C r; r.x = a.y; r.y = a.x; return r; }

// ---------------------------------------------------------------------------------------------------------
// register butterflies: forward DFT of length R, input and output in natural order
// ---------------------------------------------------------------------------------------------------------
template<typename T, int R> struct butterfly;

template<typename T> struct butterfly<T, 2>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[2]){
        cplx<T> a = v[0];
        v[0] = cadd(a, v[1]);
        v[1] = csub(a, v[1]);
    }
};

template<typename T> struct butterfly<T, 4>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[4]){
        cplx<T> s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
        cplx<T> s13 = cadd(v[1], v[3]), d13 = mul_mi(csub(v[1], v[3]));
        v[0] = cadd(s02, s13);
        v[2] = csub(s02, s13);
        v[1] = cadd(d02, d13);
        v[3] = csub(d02, d13);
    }
};

template<typename T> struct butterfly<T, 8>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[8]){
        const T h = static_cast<T>(0.70710678118654752440084436210485);
        // split into even part a (k + k+4 sums) and twiddled odd part b
        cplx<T> a[4], b[4];
        #pragma unroll
        for(int k=0; k<4; k++){
            a[k] = cadd(v[k], v[k+4]);
            b[k] = csub(v[k], v[k+4]);
        }
        // b[k] *= W8^k : W8 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
        b[1] = mk<T>((b[1].x + b[1].y) * h, (b[1].y - b[1].x) * h);
        b[2] = mul_mi(b[2]);
        b[3] = mk<T>((b[3].y - b[3].x) * h, -(b[3].x + b[3].y) * h);
        butterfly<T, 4>::run(a);
        butterfly<T, 4>::run(b);
        #pragma unroll
        for(int k=0; k<4; k++){
            v[2*k]   = a[k];
            v[2*k+1] = b[k];
        }
    }
};

template<typename T> struct butterfly<T, 16>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[16]){
        // 16 = 4 x 4: four radix-4 over stride 4, twiddle by W16^(j*k), four radix-4 over the columns
        const T c1 = static_cast<T>(0.92387953251128675612818318939679); // cos(pi/8)
        const T s1 = static_cast<T>(0.38268343236508977172845998403040); // sin(pi/8)
        const T h  = static_cast<T>(0.70710678118654752440084436210485);
        cplx<T> col[4][4];
        #pragma unroll
        for(int j=0; j<4; j++){
            cplx<T> t[4] = {v[j], v[j+4], v[j+8], v[j+12]};
            butterfly<T, 4>::run(t);
            #pragma unroll
            for(int k=0; k<4; k++) col[j][k] = t[k];
        }
        // twiddles W16^(j*k), forward sign: W16^m = cos(m pi/8) - i sin(m pi/8)
        col[1][1] = cmul(col[1][1], mk<T>( c1, -s1));
        col[1][2] = mk<T>((col[1][2].x + col[1][2].y) * h, (col[1][2].y - col[1][2].x) * h);
        col[1][3] = cmul(col[1][3], mk<T>( s1, -c1));
        col[2][1] = mk<T>((col[2][1].x + col[2][1].y) * h, (col[2][1].y - col[2][1].x) * h);
        col[2][2] = mul_mi(col[2][2]);
        col[2][3] = mk<T>((col[2][3].y - col[2][3].x) * h, -(col[2][3].x + col[2][3].y) * h);
        col[3][1] = cmul(col[3][1], mk<T>( s1, -c1));
        col[3][2] = mk<T>((col[3][2].y - col[3][2].x) * h, -(col[3][2].x + col[3][2].y) * h);
        col[3][3] = cmul(col[3][3], mk<T>(-c1,  s1)); // W16^9 = -W16^1
        #pragma unroll
        for(int k=0; k<4; k++){
            cplx<T> t[4] = {col[0][k], col[1][k], col[2][k], col[3][k]};
            butterfly<T, 4>::run(t);
            #pragma unroll
            for(int j=0; j<4; j++) v[k + 4*j] = t[j];
        }
    }
};

// ---- odd and composite radices (lengths 3 * 2^k, 5 * 2^k, 10^k ...) ----------------------------------------------
template<typename T> struct butterfly<T, 3>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[3]){
        const T s = static_cast<T>(0.86602540378443864676372317075294);   // sin(2 pi / 3)
        const cplx<T> t = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        const cplx<T> m = mk<T>(v[0].x - t.x * T(0.5), v[0].y - t.y * T(0.5));
        const cplx<T> r = mk<T>(d.y * s, -d.x * s);                      // -i s (v1 - v2)
        v[0] = cadd(v[0], t);
        v[1] = cadd(m, r);
        v[2] = csub(m, r);
    }
};

template<typename T> struct butterfly<T, 5>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[5]){
        const T c1 = static_cast<T>( 0.30901699437494742410229341718282);  // cos(2 pi / 5)
        const T c2 = static_cast<T>(-0.80901699437494742410229341718282);  // cos(4 pi / 5)
        const T s1 = static_cast<T>( 0.95105651629515357211643933337938);  // sin(2 pi / 5)
        const T s2 = static_cast<T>( 0.58778525229247312916870595463907);  // sin(4 pi / 5)
        const cplx<T> a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        const cplx<T> r1 = mk<T>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        const cplx<T> r2 = mk<T>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        const cplx<T> i1 = mk<T>(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
        const cplx<T> i2 = mk<T>(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
        v[0] = mk<T>(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = mk<T>(r1.x + i1.y, r1.y - i1.x);      // r1 - i i1
        v[4] = mk<T>(r1.x - i1.y, r1.y + i1.x);      // r1 + i i1
        v[2] = mk<T>(r2.x + i2.y, r2.y - i2.x);
        v[3] = mk<T>(r2.x - i2.y, r2.y + i2.x);
    }
};

template<typename T> struct butterfly<T, 7>{
    __device__ __forceinline__ static void run(cplx<T> (&v)[7]){
        // cos / sin of 2 pi m / 7, m = 1, 2, 3
        const T c1 = static_cast<T>( 0.62348980185873353052500488400424), c2 = static_cast<T>(-0.22252093395631440428890256449679),
                c3 = static_cast<T>(-0.90096886790241912623610231950745);
        const T s1 = static_cast<T>( 0.78183148246802980870844452667406), s2 = static_cast<T>( 0.97492791218182360701813168299393),
                s3 = static_cast<T>( 0.43388373911755812047576833284836);
        const cplx<T> a1 = cadd(v[1], v[6]), a2 = cadd(v[2], v[5]), a3 = cadd(v[3], v[4]);
        const cplx<T> b1 = csub(v[1], v[6]), b2 = csub(v[2], v[5]), b3 = csub(v[3], v[4]);
        // even parts r_m = v0 + sum_k cos(2 pi m k / 7) a_k,  odd parts i_m = sum_k sin(2 pi m k / 7) b_k
        const cplx<T> r1 = mk<T>(v[0].x + c1 * a1.x + c2 * a2.x + c3 * a3.x, v[0].y + c1 * a1.y + c2 * a2.y + c3 * a3.y);
        const cplx<T> r2 = mk<T>(v[0].x + c2 * a1.x + c3 * a2.x + c1 * a3.x, v[0].y + c2 * a1.y + c3 * a2.y + c1 * a3.y);
        const cplx<T> r3 = mk<T>(v[0].x + c3 * a1.x + c1 * a2.x + c2 * a3.x, v[0].y + c3 * a1.y + c1 * a2.y + c2 * a3.y);
        const cplx<T> i1 = mk<T>(s1 * b1.x + s2 * b2.x + s3 * b3.x, s1 * b1.y + s2 * b2.y + s3 * b3.y);
        const cplx<T> i2 = mk<T>(s2 * b1.x - s3 * b2.x - s1 * b3.x, s2 * b1.y - s3 * b2.y - s1 * b3.y);
        const cplx<T> i3 = mk<T>(s3 * b1.x - s1 * b2.x + s2 * b3.x, s3 * b1.y - s1 * b2.y + s2 * b3.y);
        v[0] = mk<T>(v[0].x + a1.x + a2.x + a3.x, v[0].y + a1.y + a2.y + a3.y);
        v[1] = mk<T>(r1.x + i1.y, r1.y - i1.x);   v[6] = mk<T>(r1.x - i1.y, r1.y + i1.x);      // r - i i  /  r + i i
        v[2] = mk<T>(r2.x + i2.y, r2.y - i2.x);   v[5] = mk<T>(r2.x - i2.y, r2.y + i2.x);
        v[3] = mk<T>(r3.x + i3.y, r3.y - i3.x);   v[4] = mk<T>(r3.x - i3.y, r3.y + i3.x);
    }
};

// R = P * Q in registers: P-point transforms over n1 (input n = n1 Q + n2), twiddles W_R^(n2 k1), Q-point transforms over n2,
// output k = k1 + P k2.  COS / SIN hold cos / sin of 2 pi m / R for m = 0 .. R-1.
template<typename T, int P, int Q, typename TABLE>
__device__ __forceinline__ void composite_butterfly(cplx<T> (&v)[P * Q]){
    cplx<T> a[Q][P];
    #pragma unroll
    for(int n2=0; n2<Q; n2++){
        cplx<T> t[P];
        #pragma unroll
        for(int n1=0; n1<P; n1++) t[n1] = v[n1 * Q + n2];
        butterfly<T, P>::run(t);
        #pragma unroll
        for(int k1=0; k1<P; k1++){
            constexpr int R = P * Q;
            const int m = (n2 * k1) % R;
            a[n2][k1] = (m == 0) ? t[k1] : cmul(t[k1], mk<T>(static_cast<T>(TABLE::cosine(m)), static_cast<T>(-TABLE::sine(m))));
        }
    }
    #pragma unroll
    for(int k1=0; k1<P; k1++){
        cplx<T> t[Q];
        #pragma unroll
        for(int n2=0; n2<Q; n2++) t[n2] = a[n2][k1];
        butterfly<T, Q>::run(t);
        #pragma unroll
        for(int k2=0; k2<Q; k2++) v[k1 + P * k2] = t[k2];
    }
}
struct unit_circle_6 {
    __host__ __device__ static constexpr double cosine(int m){ return (m == 0) ? 1.0 : (m == 1 || m == 5) ? 0.5 : (m == 3) ? -1.0 : -0.5; }
    __host__ __device__ static constexpr double sine(int m){ return (m == 0 || m == 3) ? 0.0 : (m < 3) ? 0.86602540378443864676 : -0.86602540378443864676; }
};
struct unit_circle_10 {   // 36 degrees
    __host__ __device__ static constexpr double cosine(int m){
        return (m == 0) ? 1.0 : (m == 5) ? -1.0 : (m == 1 || m == 9) ? 0.80901699437494742410 : (m == 2 || m == 8) ? 0.30901699437494742410 :
               (m == 3 || m == 7) ? -0.30901699437494742410 : -0.80901699437494742410;
    }
    __host__ __device__ static constexpr double sine(int m){
        return (m == 0 || m == 5) ? 0.0 : (m == 1 || m == 4) ? 0.58778525229247312917 : (m == 2 || m == 3) ? 0.95105651629515357212 :
               (m == 6 || m == 9) ? -0.58778525229247312917 : -0.95105651629515357212;
    }
};
struct unit_circle_12 {   // 30 degrees
    __host__ __device__ static constexpr double cosine(int m){
        return (m == 0) ? 1.0 : (m == 6) ? -1.0 : (m == 3 || m == 9) ? 0.0 : (m == 1 || m == 11) ? 0.86602540378443864676 : (m == 2 || m == 10) ? 0.5 :
               (m == 4 || m == 8) ? -0.5 : -0.86602540378443864676;
    }
    __host__ __device__ static constexpr double sine(int m){
        return (m == 0 || m == 6) ? 0.0 : (m == 3) ? 1.0 : (m == 9) ? -1.0 : (m == 1 || m == 5) ? 0.5 : (m == 2 || m == 4) ? 0.86602540378443864676 :
               (m == 7 || m == 11) ? -0.5 : -0.86602540378443864676;
    }
};
template<typename T> struct butterfly<T, 6>{  __device__ __forceinline__ static void run(cplx<T> (&v)[6]){  composite_butterfly<T, 2, 3, unit_circle_6>(v); } };
template<typename T> struct butterfly<T, 10>{ __device__ __forceinline__ static void run(cplx<T> (&v)[10]){ composite_butterfly<T, 2, 5, unit_circle_10>(v); } };
template<typename T> struct butterfly<T, 12>{ __device__ __forceinline__ static void run(cplx<T> (&v)[12]){ composite_butterfly<T, 4, 3, unit_circle_12>(v); } };

// ---------------------------------------------------------------------------------------------------------
// addressing of a batch of lines: line l = (a, b) with a = l % count_a; element i of the line lives at
//   base + a*stride_a + b*stride_b + i*stride      (all in elements of the scalar type of that side)
// ---------------------------------------------------------------------------------------------------------
struct line_geom {
    long long stride;
    long long stride_a;
    long long stride_b;
};

struct fft_args {
    const void *in;
    void *out;
    const void *twiddle;   // W_N^k, k = 0..N-1, forward sign, complex of the working precision
    const void *twiddle2;  // real-data kernels only: W_{4n}^j, j = 0..n (n = real length = 2N)
    line_geom ig, og;
    long long nlines;
    int count_a;
    int backward;          // 0 forward, 1 backward
    double scale;          // applied on the final store
    const scatter_map *smap;   // device pointer; non-null selects the scatter variants (og is ignored)
    // batched transforms (reference include/heffte_fft3d.h:391-414): entry e = blockIdx.y works on in + e * in_step and
    // out + e * out_step (bytes).  Fused stores add e * scatter_step to every destination, and local_shift + e * local_step to
    // the destinations that lie in this rank's own memory (scatter_map::local_mask): that is how the last stage of a plan lands
    // its own part in the caller's array without a per-call copy of the map.
    long long in_step, out_step, scatter_step, local_shift, local_step;
    const void *multiplier;    // fused spectral operator (fft_strided_conv_kernel): null = the spectrum times itself
    // two kernels that overlap on two streams (a local transform feeding the fused transform behind it, plane by plane):
    // done[b] counts the tiles of the producer stored for plane b (b = line / count_a).  done_mode 1: this launch signals,
    // 2: this launch waits until done[b] >= done_need before it touches plane b.  order_nb > 1: the planes are visited
    // round-robin over that many ranges (the order of the fused kernels, scatter_tile_order) so that producer and consumer agree.
    unsigned *done;
    int done_mode;
    unsigned done_need;
    int order_nb;
    const void *twiddle0;      // real-data kernels, second generation: W_n^k for the full real length n
};
struct batch_shift { long long all, local; };
// the arguments of batch entry blockIdx.y
__device__ __forceinline__ fft_args batch_entry(fft_args a, batch_shift &shift){
    const long long e = blockIdx.y;
    a.in = static_cast<const char*>(a.in) + e * a.in_step;
    if (a.out != nullptr) a.out = static_cast<char*>(a.out) + e * a.out_step;
    shift.all = e * a.scatter_step;
    shift.local = a.local_shift + e * a.local_step;
    if (a.done_mode != 0) a.done += e * (a.nlines / a.count_a);      // one set of plane counters per entry
    return a;
}

// ---------------------------------------------------------------------------------------------------------
// scatter map: where the output of a batched transform goes when the following reshape is fused into the store.
// The local box is cut into cells (<= 8 per axis) such that every cell lands in ONE destination box; a cell knows the
// address of its destination (local memory or a peer GPU's memory mapped over NVLink) and the strides of that box.
// Coordinates are local: k = index along the transform, (a, b) = line index split as a = line % count_a, b = line / count_a.
// Replaces, fused: direct_packer::pack + MPI_Alltoallv + direct/transpose_packer::unpack of the reference
// (src/heffte_reshape3d.cpp:365-443, include/heffte_pack3d.h:89-197).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int scatter_find(const int *cut, int n, int x){
    int c = 0;
    for(int i=1; i<n; i++) c += (x >= cut[i]) ? 1 : 0;
    return c;
}
// cell row of a line: constant for a thread that owns one line
__device__ __forceinline__ int scatter_row(const scatter_map *m, int a, int b){
    return scatter_find(m->cut_a, m->na, a) * m->nb + scatter_find(m->cut_b, m->nb, b);
}
template<typename V>
__device__ __forceinline__ V* scatter_address(const scatter_map *m, int row, int k, int a, int b){
    const int ck = scatter_find(m->cut_k, m->nk, k);
    const scatter_cell &c = m->cell[ck * (m->na * m->nb) + row];
    return reinterpret_cast<V*>(c.base) + (k * c.sk + a * c.sa + b * c.sb);
}
// every thread of the CTA takes part; the caller synchronises before the map is used.  The destination addresses are
// re-based on the way: shift.all for every cell, shift.local on top for the cells that stay in this rank's memory.
__device__ __forceinline__ void scatter_stage(scatter_map *dst, const scatter_map *src, batch_shift shift = batch_shift{0, 0}){
    const int words = (scatter_header_bytes + static_cast<int>(sizeof(scatter_cell)) * src->ncells) / 16;
    const unsigned long long mask = src->local_mask;
    const int4 *g = reinterpret_cast<const int4*>(src);
    int4 *s = reinterpret_cast<int4*>(dst);
    constexpr int header_words = scatter_header_bytes / 16;
    for(int i = threadIdx.x; i < words; i += blockDim.x){
        int4 v = g[i];
        const int w = i - header_words;
        if (w >= 0 && (w & 1) == 0){          // first half of a cell: the 64-bit base address sits in (x, y)
            long long base = static_cast<long long>((static_cast<unsigned long long>(static_cast<unsigned>(v.y)) << 32) | static_cast<unsigned>(v.x));
            base += shift.all + (((mask >> (w >> 1)) & 1ULL) ? shift.local : 0);
            v.x = static_cast<int>(static_cast<unsigned long long>(base) & 0xffffffffULL);
            v.y = static_cast<int>(static_cast<unsigned long long>(base) >> 32);
        }
        s[i] = v;
    }
}
// the same resolution for kernels that read the map from global memory (generic kernel)
template<typename V>
__device__ __forceinline__ V* scatter_address_shifted(const scatter_map *m, int row, int k, int a, int b, batch_shift shift){
    const int ck = scatter_find(m->cut_k, m->nk, k);
    const int index = ck * (m->na * m->nb) + row;
    const scatter_cell &c = m->cell[index];
    const long long base = c.base + shift.all + (((m->local_mask >> index) & 1ULL) ? shift.local : 0);
    return reinterpret_cast<V*>(base) + (k * c.sk + a * c.sa + b * c.sb);
}

__device__ __forceinline__ long long line_offset(line_geom const &g, int count_a, long long line){
    long long b = line / count_a;
    long long a = line - b * count_a;
    return a * g.stride_a + b * g.stride_b;
}

template<typename T> __device__ __forceinline__ cplx<T> ldg_c(const cplx<T> *p){ return __ldg(p); }

// asynchronous global -> shared copy of one complex element (LDGSTS): no register staging, the whole tile is in flight
#ifndef B200_HOST_EMULATION
template<int BYTES> __device__ __forceinline__ void async_copy(void *smem_dst, const void *gsrc){
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gsrc));
    else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(s), "l"(gsrc));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void async_wait_all(){ asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
#else
template<int BYTES> inline void async_copy(void *smem_dst, const void *gsrc){ std::memcpy(smem_dst, gsrc, BYTES); }
inline void async_wait_all(){}
#endif

#ifndef B200_HOST_EMULATION
__device__ __forceinline__ unsigned load_acquire(const unsigned *p){
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ticket_take(unsigned *p){ return atomicAdd(p, 1u); }
__device__ __forceinline__ void count_release(unsigned *p){ __threadfence(); atomicAdd(p, 1u); }
__device__ __forceinline__ void count_release_n(unsigned *p, unsigned n){ __threadfence(); atomicAdd(p, n); }
#else
inline unsigned load_acquire(const unsigned *p){ return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline unsigned ticket_take(unsigned *p){ return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
inline void count_release(unsigned *p){ __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
inline void count_release_n(unsigned *p, unsigned n){ __atomic_fetch_add(p, n, __ATOMIC_ACQ_REL); }
#endif

__host__ __device__ constexpr int cmax(int a, int b){ return a > b ? a : b; }
__host__ __device__ constexpr int ilog2(int n){ return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// apply W^(o*r*step) to v[r], r = 1..R-1; table holds W_N^k.  FEW_LOADS: fetch only the power-of-two entries
// from the table and form the others by one or two products (error <= 2 roundings).
template<typename T, int R, bool FEW_LOADS>
__device__ __forceinline__ void apply_twiddles(cplx<T> (&v)[R], const cplx<T> *tw, int base){
    if constexpr (!FEW_LOADS || (R != 8 && R != 16)){
        #pragma unroll
        for(int r=1; r<R; r++) v[r] = cmul(v[r], ldg_c<T>(tw + r * base));
    }else{
        cplx<T> w1 = ldg_c<T>(tw + base), w2 = ldg_c<T>(tw + 2 * base), w4 = ldg_c<T>(tw + 4 * base);
        cplx<T> w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
        v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
        if constexpr (R == 16){
            cplx<T> w8 = ldg_c<T>(tw + 8 * base);
            v[8]  = cmul(v[8],  w8);
            v[9]  = cmul(v[9],  cmul(w8, w1)); v[10] = cmul(v[10], cmul(w8, w2)); v[11] = cmul(v[11], cmul(w8, w3));
            v[12] = cmul(v[12], cmul(w8, w4)); v[13] = cmul(v[13], cmul(w8, w5)); v[14] = cmul(v[14], cmul(w8, w6));
            v[15] = cmul(v[15], cmul(w8, w7));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// radix schedules
// ---------------------------------------------------------------------------------------------------------
template<int R0, int R1, int R2, int R3> struct radix_list {
    static constexpr int N = R0 * R1 * R2 * R3;
    static constexpr int passes = (R1 == 1) ? 1 : ((R2 == 1) ? 2 : ((R3 == 1) ? 3 : 4));
    static constexpr int rmax = cmax(cmax(R0, R1), cmax(R2, R3));
    // only ever called with compile-time arguments
    __host__ __device__ static constexpr int radix(int s){ return s == 0 ? R0 : (s == 1 ? R1 : (s == 2 ? R2 : R3)); }
    // distance between the legs of a butterfly of pass s: N / (R0 * ... * Rs)
    __host__ __device__ static constexpr int stride(int s){ return s == 0 ? N / R0 : (s == 1 ? N / (R0 * R1) : (s == 2 ? N / (R0 * R1 * R2) : 1)); }
};

// natural index k of the value that ends at in-place position p after all DIF passes (all powers of two: shifts/masks)
template<typename RL>
__device__ __forceinline__ unsigned dif_output_index(unsigned p){
    constexpr unsigned r0 = RL::radix(0), s0 = RL::stride(0);
    unsigned k = p / s0;
    if constexpr (RL::passes > 1){ constexpr unsigned r1 = RL::radix(1), s1 = RL::stride(1); k += ((p / s1) % r1) * r0; }
    if constexpr (RL::passes > 2){ constexpr unsigned r2 = RL::radix(2), s2 = RL::stride(2); k += ((p / s2) % r2) * (r0 * RL::radix(1)); }
    if constexpr (RL::passes > 3){ constexpr unsigned r3 = RL::radix(3); k += (p % r3) * (r0 * RL::radix(1) * RL::radix(2)); }
    return k;
}

__device__ __forceinline__ long long tile_line_offset(line_geom const &g, int count_a, unsigned line){
    unsigned b = line / static_cast<unsigned>(count_a);
    unsigned a = line - b * static_cast<unsigned>(count_a);
    return static_cast<long long>(a) * g.stride_a + static_cast<long long>(b) * g.stride_b;
}

// Order in which the tiles of a fused FFT + reshape visit the box: round-robin over the nb destination ranges of the slower
// line axis, so that the CTAs that run side by side write to every destination GPU at once.  The ranks of a plan walk
// their boxes in step: a plain sweep aims every sender at the same receivers at the same time and the inbound NVLink of
// those GPUs becomes the limit (measured on 4 GPUs: 417 GB/s per sender instead of 680).  A bijection on the tile
// indices; the identity when the tiles straddle rows of the box or the ranges do not divide it.
template<int LPB>
__device__ __forceinline__ unsigned round_robin_tile_order(fft_args const &a, unsigned nb, unsigned blk);
template<int LPB>
__device__ __forceinline__ unsigned scatter_tile_order(fft_args const &a, unsigned blk){
    return round_robin_tile_order<LPB>(a, static_cast<unsigned>(a.smap->nb), blk);
}
template<int LPB>
__device__ __forceinline__ unsigned round_robin_tile_order(fft_args const &a, unsigned nb, unsigned blk){
    const unsigned ca = static_cast<unsigned>(a.count_a);
    if (nb <= 1 || ca % LPB != 0) return blk;
    const unsigned tiles_per_b = ca / LPB;
    const unsigned count_b = static_cast<unsigned>(a.nlines / ca);
    if (count_b % nb != 0) return blk;
    const unsigned visit = blk / tiles_per_b, tile = blk - visit * tiles_per_b;
    const unsigned b = (visit % nb) * (count_b / nb) + visit / nb;
    return b * tiles_per_b + tile;
}

// ---------------------------------------------------------------------------------------------------------
// strided kernel: a CTA owns a tile of LPB adjacent lines, shared memory is laid out [position][line]; the tile is
// brought in with asynchronous copies (LDGSTS, all of it in flight at once, no registers held across the load), the
// passes are in-place decimation-in-frequency, and the digit reversal is absorbed into the row index of the store.
// ---------------------------------------------------------------------------------------------------------
struct scatter_ctx { const scatter_map *map; int row, a, b; };   // per-thread view of the map (one line per thread)

template<typename T, typename RL, int S, int TPL, int LPB, bool BWD, bool SCATTER>
__device__ __forceinline__ void strided_pass(cplx<T> *sm, unsigned t, unsigned j, bool valid, cplx<T> *gout, long long ostride,
                                             const cplx<T> *tw, T scale, bool do_scale, scatter_ctx const &sc){
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned ST = RL::stride(S);         // distance between butterfly legs
    constexpr unsigned NB = RL::N / R;             // butterflies per line
    constexpr bool FIRST = (S == 0), LAST = (S == RL::passes - 1);
    #pragma unroll
    for(unsigned u=0; u<NB/TPL; u++){
        const unsigned q = j + u * TPL;
        const unsigned o = q % ST;
        const unsigned p0 = (q / ST) * (ST * R) + o;
        cplx<T> *cell = sm + p0 * LPB + t;
        cplx<T> v[R];
        #pragma unroll
        for(unsigned r=0; r<R; r++){
            cplx<T> x = cell[r * ST * LPB];
            v[r] = (FIRST && BWD) ? cswap(x) : x;
        }
        butterfly<T, R>::run(v);
        if constexpr (!LAST){
            // DIF twiddle after the butterfly: W_{ST*R}^(o*r) = W_N^(o*r*N/(ST*R))
            apply_twiddles<T, R, true>(v, tw, o * (RL::N / (ST * R)));
            #pragma unroll
            for(unsigned r=0; r<R; r++) cell[r * ST * LPB] = v[r];
        }else{
            if (valid){
                // last pass: ST == 1, p = q*R + r, so the natural index is k(q*R) + r * N/R
                const unsigned k0 = dif_output_index<RL>(p0);
                if constexpr (SCATTER){
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        cplx<T> x = BWD ? cswap(v[r]) : v[r];
                        if (do_scale){ x.x *= scale; x.y *= scale; }
                        *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(k0 + r * (RL::N / R)), sc.a, sc.b) = x;
                    }
                }else{
                    cplx<T> *dst = gout + static_cast<long long>(k0) * ostride;
                    const long long hop = static_cast<long long>(RL::N / R) * ostride;
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        cplx<T> x = BWD ? cswap(v[r]) : v[r];
                        if (do_scale){ x.x *= scale; x.y *= scale; }
                        *dst = x;
                        dst += hop;
                    }
                }
            }
        }
    }
}

// One tile of the strided kernel: LPB adjacent lines starting at line tile * LPB (after the destination round-robin of the
// scatter variants).  `smap` is the scatter map ALREADY staged in shared memory by the caller (null for plain stores).
// Every thread of the CTA takes part; the caller separates two tiles that share the buffer by a __syncthreads().
template<typename T, typename RL, int TPL, int LPB, bool BWD, bool SCATTER>
__device__ __forceinline__ void strided_tile(cplx<T> *sm, const scatter_map *smap, fft_args const &a, unsigned tile){
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const unsigned line = (SCATTER ? scatter_tile_order<LPB>(a, tile) : tile) * LPB + t;
    const bool valid = line < a.nlines;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    constexpr int P = RL::passes;

    if (valid){
        const cplx<T> *src = reinterpret_cast<const cplx<T>*>(a.in) + tile_line_offset(a.ig, a.count_a, line) + static_cast<long long>(j) * a.ig.stride;
        const long long hop = static_cast<long long>(TPL) * a.ig.stride;
        cplx<T> *dst = sm + j * LPB + t;
        #pragma unroll 4
        for(unsigned row = j; row < RL::N; row += TPL){
            async_copy<sizeof(cplx<T>)>(dst, src);
            src += hop;
            dst += TPL * LPB;
        }
    }
    cplx<T> *gout = nullptr;
    scatter_ctx sc{nullptr, 0, 0, 0};
    if constexpr (SCATTER){
        sc.map = smap;
        sc.b = static_cast<int>(line / static_cast<unsigned>(a.count_a));
        sc.a = static_cast<int>(line - static_cast<unsigned>(sc.b) * static_cast<unsigned>(a.count_a));
    }else{
        gout = reinterpret_cast<cplx<T>*>(a.out) + (valid ? tile_line_offset(a.og, a.count_a, line) : 0);
    }
    async_wait_all();
    __syncthreads();
    if constexpr (SCATTER) sc.row = valid ? scatter_row(sc.map, sc.a, sc.b) : 0;

    strided_pass<T, RL, 0, TPL, LPB, BWD, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, scale, do_scale, sc);
    if constexpr (P > 1){
        __syncthreads();
        strided_pass<T, RL, 1, TPL, LPB, BWD, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 2){
        __syncthreads();
        strided_pass<T, RL, 2, TPL, LPB, BWD, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 3){
        __syncthreads();
        strided_pass<T, RL, 3, TPL, LPB, BWD, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, scale, do_scale, sc);
    }
}

// number of tiles of a launch
template<int LPB> __host__ __device__ inline unsigned tile_count(fft_args const &a){ return static_cast<unsigned>((a.nlines + LPB - 1) / LPB); }

// Overlap of two launches plane by plane (fft_args::done).  A consumer waits for the planes its tile touches; a producer visits
// the planes in the consumer's order and reports every tile it has stored.  Whole CTA; ends with a barrier when it waited.
template<int LPB>
__device__ __forceinline__ void wait_for_planes(fft_args const &a, unsigned ordered_tile){
    if (a.done_mode != 2) return;
    if (threadIdx.x == 0){
        const unsigned first = (ordered_tile * LPB) / static_cast<unsigned>(a.count_a);
        long long last_line = static_cast<long long>(ordered_tile) * LPB + LPB - 1;
        if (last_line >= a.nlines) last_line = a.nlines - 1;
        const unsigned last = static_cast<unsigned>(last_line / a.count_a);
        for(unsigned b = first; b <= last; b++){ while(load_acquire(a.done + b) < a.done_need){} }
    }
    __syncthreads();
}
template<int LPB>
__device__ __forceinline__ void report_plane(fft_args const &a, unsigned ordered_tile){
    if (a.done_mode != 1) return;
    __syncthreads();                   // every thread has issued its stores
    if (threadIdx.x == 0){
        // the LINES of this tile, plane by plane (a tile may straddle two planes): the consumer waits for count_a lines
        const long long l0 = static_cast<long long>(ordered_tile) * LPB;
        long long l1 = l0 + LPB;
        if (l1 > a.nlines) l1 = a.nlines;
        for(long long l = l0; l < l1; ){
            const long long b = l / a.count_a;
            long long end = (b + 1) * a.count_a;
            if (end > l1) end = l1;
            count_release_n(a.done + b, static_cast<unsigned>(end - l));
            l = end;
        }
    }
}

// The kernel walks the tiles grid-stride: one tile per CTA when the grid covers the box (the usual launch), several when the
// grid is kept thin on purpose so that another kernel can share the SMs (NVLink-bound stages of a multi-GPU plan).
template<typename T, typename RL, int TPL, int LPB, int MINB, bool BWD, bool SCATTER>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    cplx<T> *sm = reinterpret_cast<cplx<T>*>(smem_raw);
    scatter_map *smap = nullptr;
    if constexpr (SCATTER){
        smap = reinterpret_cast<scatter_map*>(sm + static_cast<size_t>(RL::N) * LPB);   // behind the tile
        scatter_stage(smap, a.smap, shift);      // visible after the first barrier inside strided_tile
    }
    const unsigned ntiles = tile_count<LPB>(a);
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        if constexpr (SCATTER){
            if (a.done_mode == 2) wait_for_planes<LPB>(a, scatter_tile_order<LPB>(a, tile));
            strided_tile<T, RL, TPL, LPB, BWD, SCATTER>(sm, smap, a, tile);
        }else{
            const unsigned at = (a.order_nb > 1) ? round_robin_tile_order<LPB>(a, static_cast<unsigned>(a.order_nb), tile) : tile;
            wait_for_planes<LPB>(a, at);
            strided_tile<T, RL, TPL, LPB, BWD, SCATTER>(sm, smap, a, at);
            report_plane<LPB>(a, at);
        }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

#ifndef B200_HOST_EMULATION
// ---------------------------------------------------------------------------------------------------------
// The strided kernel with its tile brought in by TMA: ONE elected thread asks for the [n][LPB] box of the tile with
// cp.async.bulk.tensor (two requests of 256 rows; the tensor map -- built on the host per launch, a kernel parameter --
// describes the box of the stage as (line axis, transform axis, slower line axis, batch entry)), completion arrives on an
// mbarrier, the passes are those of fft_strided_kernel.  Measured on B200 (tools/kbench_tma.cu, 512-point fp64 lines,
// profiles/r02_single/kbench_tma.log): rows 4 MB apart (the slow axis of 512^3) 6.15 instead of 5.94 TB/s -- no per-thread address
// arithmetic and no LSU / MIO slots for 512 scattered rows; rows 8 KB apart (the middle axis) 6.64 instead of 6.77 TB/s, so
// the host side takes this kernel only for widely spaced rows.  Plain stores, no plane hooks (see fft_args::done).
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) tma_tile_map { unsigned long long opaque[16]; };      // a CUtensorMap
// host side (fft1d.cu): the tensor map of a strided stage, false when the driver or the box does not allow one
bool encode_tile_map(tma_tile_map &map, const void *base, int real_bytes, long long count_a, long long n, long long stride, long long count_b, long long stride_b,
                     int batch, long long step_bytes, int lpb);
__device__ __forceinline__ unsigned shared_address(const void *p){ return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

template<typename T, typename RL, int TPL, int LPB, int MINB, bool BWD>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_tma_kernel(fft_args a0, const __grid_constant__ tma_tile_map tmap){
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned N = RL::N;
    constexpr unsigned TILE_BYTES = N * LPB * sizeof(cplx<T>);
    constexpr int P = RL::passes;
    cplx<T> *sm = reinterpret_cast<cplx<T>*>(tma_smem_raw);
    unsigned long long *bar_word = reinterpret_cast<unsigned long long*>(tma_smem_raw + TILE_BYTES);
    const unsigned bar = shared_address(bar_word);
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    const scatter_ctx sc{nullptr, 0, 0, 0};
    if (threadIdx.x == 0){
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const unsigned ntiles = tile_count<LPB>(a);
    unsigned phase = 0;
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        if (threadIdx.x == 0){
            const unsigned line0 = tile * LPB;
            const unsigned b = line0 / static_cast<unsigned>(a.count_a);
            const unsigned first = line0 - b * static_cast<unsigned>(a.count_a);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(TILE_BYTES) : "memory");
            #pragma unroll
            for(unsigned r0 = 0; r0 < N; r0 += 256){
                const unsigned dst = shared_address(sm + static_cast<size_t>(r0) * LPB);
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                             :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(static_cast<int>(2 * first)), "r"(static_cast<int>(r0)),
                                "r"(static_cast<int>(b)), "r"(static_cast<int>(blockIdx.y)), "r"(bar) : "memory");
            }
        }
        unsigned arrived = 0;
        while(!arrived)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(arrived) : "r"(bar), "r"(phase) : "memory");
        phase ^= 1;
        const unsigned line = tile * LPB + t;
        cplx<T> *gout = reinterpret_cast<cplx<T>*>(a.out) + tile_line_offset(a.og, a.count_a, line);
        strided_pass<T, RL, 0, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc);
        if constexpr (P > 1){ __syncthreads(); strided_pass<T, RL, 1, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if constexpr (P > 2){ __syncthreads(); strided_pass<T, RL, 2, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if constexpr (P > 3){ __syncthreads(); strided_pass<T, RL, 3, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if (tile + gridDim.x < ntiles){
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // the threads' accesses of the tile come before the next bulk write
            __syncthreads();
        }
    }
}
#endif

// ---------------------------------------------------------------------------------------------------------
// fused spectral operator along a strided axis: forward transform, pointwise product, backward transform of every line in ONE
// pass over memory (reference benchmarks/convolution.cpp:86-97 runs forward(scale::full), x[i] *= x[i], backward as three
// sweeps plus the reshapes around them).  The forward part is the decimation-in-frequency kernel above, except that the last
// pass leaves the spectrum in the tile (digit-reversed positions).  The product is applied there: times `scale`, then times
// itself or times the caller's multiplier (same layout as the input box).  The backward part undoes the passes in reverse
// order -- decimation in time: twiddle, then butterfly, on re/im-swapped data so that the forward butterflies and the forward
// twiddle table serve (swap(z conj(w)) = swap(z) w, swap(conjDFT(z)) = DFT(swap(z))) -- and ends in natural order, stored
// through the plain or the fused-reshape path like any backward transform.
// ---------------------------------------------------------------------------------------------------------
// forward pass S < P-1: butterfly, twiddle, back into the tile
template<typename T, typename RL, int S, int TPL, int LPB>
__device__ __forceinline__ void conv_forward_pass(cplx<T> *sm, unsigned t, unsigned j, const cplx<T> *tw){
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned ST = RL::stride(S);
    constexpr unsigned NB = RL::N / R;
    static_assert(S < RL::passes - 1, "the last pass is the turn-around");
    #pragma unroll 1
    for(unsigned u=0; u<NB/TPL; u++){
        const unsigned q = j + u * TPL;
        const unsigned o = q % ST;
        const unsigned p0 = (q / ST) * (ST * R) + o;
        cplx<T> *cell = sm + p0 * LPB + t;
        cplx<T> v[R];
        #pragma unroll
        for(unsigned r=0; r<R; r++) v[r] = cell[r * ST * LPB];
        butterfly<T, R>::run(v);
        apply_twiddles<T, R, true>(v, tw, o * (RL::N / (ST * R)));
        #pragma unroll
        for(unsigned r=0; r<R; r++) cell[r * ST * LPB] = v[r];
    }
}

// turn-around, pass P-1 in both directions on the same R adjacent positions: last forward butterfly, the product (natural
// index of leg r: k0 + r N/R), swap, first backward butterfly (no twiddle on either side of the product)
template<typename T, typename RL, int TPL, int LPB, bool SCATTER>
__device__ __forceinline__ void conv_turn_pass(cplx<T> *sm, unsigned t, unsigned j, bool valid, T scale, const cplx<T> *mult, long long mstride,
                                                cplx<T> *gout, long long ostride, scatter_ctx const &sc){
    constexpr unsigned S = RL::passes - 1;
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned NB = RL::N / R;
    #pragma unroll 1
    for(unsigned u=0; u<NB/TPL; u++){
        const unsigned q = j + u * TPL;
        const unsigned p0 = q * R;
        cplx<T> *cell = sm + p0 * LPB + t;
        cplx<T> v[R];
        #pragma unroll
        for(unsigned r=0; r<R; r++) v[r] = cell[r * LPB];
        butterfly<T, R>::run(v);
        const unsigned k0 = dif_output_index<RL>(p0);
        #pragma unroll
        for(unsigned r=0; r<R; r++){
            cplx<T> x = mk<T>(v[r].x * scale, v[r].y * scale);
            cplx<T> m = x;
            if (mult != nullptr) m = valid ? mult[static_cast<long long>(k0 + r * (RL::N / R)) * mstride] : mk<T>(0, 0);
            v[r] = cswap(cmul(x, m));
        }
        butterfly<T, R>::run(v);
        if constexpr (S > 0){
            #pragma unroll
            for(unsigned r=0; r<R; r++) cell[r * LPB] = v[r];
        }else{
            // a single pass: the turn-around is the whole transform
            if (valid){
                #pragma unroll
                for(unsigned r=0; r<R; r++){
                    const cplx<T> x = cswap(v[r]);
                    if constexpr (SCATTER) *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(p0 + r), sc.a, sc.b) = x;
                    else gout[static_cast<long long>(p0 + r) * ostride] = x;
                }
            }
        }
    }
}

// backward pass S < P-1 (decimation in time on swapped data): twiddle, butterfly; pass 0 stores the result
template<typename T, typename RL, int S, int TPL, int LPB, bool SCATTER>
__device__ __forceinline__ void conv_backward_pass(cplx<T> *sm, unsigned t, unsigned j, bool valid, cplx<T> *gout, long long ostride,
                                                    const cplx<T> *tw, scatter_ctx const &sc){
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned ST = RL::stride(S);
    constexpr unsigned NB = RL::N / R;
    static_assert(S < RL::passes - 1, "the last pass is the turn-around");
    #pragma unroll 1
    for(unsigned u=0; u<NB/TPL; u++){
        const unsigned q = j + u * TPL;
        const unsigned o = q % ST;
        const unsigned p0 = (q / ST) * (ST * R) + o;
        cplx<T> *cell = sm + p0 * LPB + t;
        cplx<T> v[R];
        #pragma unroll
        for(unsigned r=0; r<R; r++) v[r] = cell[r * ST * LPB];
        apply_twiddles<T, R, true>(v, tw, o * (RL::N / (ST * R)));
        butterfly<T, R>::run(v);
        if constexpr (S > 0){
            #pragma unroll
            for(unsigned r=0; r<R; r++) cell[r * ST * LPB] = v[r];
        }else{
            // pass 0 comes last: position p0 + r ST is the natural index
            if (valid){
                #pragma unroll
                for(unsigned r=0; r<R; r++){
                    const cplx<T> x = cswap(v[r]);
                    if constexpr (SCATTER) *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(p0 + r * ST), sc.a, sc.b) = x;
                    else gout[static_cast<long long>(p0 + r * ST) * ostride] = x;
                }
            }
        }
    }
}

template<typename T, typename RL, int TPL, int LPB, int MINB, bool SCATTER>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_conv_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    cplx<T> *sm = reinterpret_cast<cplx<T>*>(smem_raw);
    scatter_map *smap = nullptr;
    if constexpr (SCATTER){
        smap = reinterpret_cast<scatter_map*>(sm + static_cast<size_t>(RL::N) * LPB);
        scatter_stage(smap, a.smap, shift);
    }
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const T scale = static_cast<T>(a.scale);
    constexpr int P = RL::passes;
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const unsigned ntiles = tile_count<LPB>(a);
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        const unsigned line = (SCATTER ? scatter_tile_order<LPB>(a, tile) : tile) * LPB + t;
        const bool valid = line < a.nlines;
        const long long ioff = valid ? tile_line_offset(a.ig, a.count_a, line) : 0;
        if (valid){
            const cplx<T> *src = reinterpret_cast<const cplx<T>*>(a.in) + ioff + static_cast<long long>(j) * a.ig.stride;
            const long long hop = static_cast<long long>(TPL) * a.ig.stride;
            cplx<T> *dst = sm + j * LPB + t;
            #pragma unroll 4
            for(unsigned row = j; row < RL::N; row += TPL){
                async_copy<sizeof(cplx<T>)>(dst, src);
                src += hop;
                dst += TPL * LPB;
            }
        }
        const cplx<T> *mult = (a.multiplier != nullptr) ? reinterpret_cast<const cplx<T>*>(a.multiplier) + ioff : nullptr;
        cplx<T> *gout = nullptr;
        scatter_ctx sc{nullptr, 0, 0, 0};
        if constexpr (SCATTER){
            sc.map = smap;
            sc.b = static_cast<int>(line / static_cast<unsigned>(a.count_a));
            sc.a = static_cast<int>(line - static_cast<unsigned>(sc.b) * static_cast<unsigned>(a.count_a));
        }else{
            gout = reinterpret_cast<cplx<T>*>(a.out) + (valid ? tile_line_offset(a.og, a.count_a, line) : 0);
        }
        async_wait_all();
        __syncthreads();
        if constexpr (SCATTER) sc.row = valid ? scatter_row(sc.map, sc.a, sc.b) : 0;

        if constexpr (P > 1){ conv_forward_pass<T, RL, 0, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
        if constexpr (P > 2){ conv_forward_pass<T, RL, 1, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
        if constexpr (P > 3){ conv_forward_pass<T, RL, 2, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
        conv_turn_pass<T, RL, TPL, LPB, SCATTER>(sm, t, j, valid, scale, mult, a.ig.stride, gout, a.og.stride, sc);
        if constexpr (P > 3){ __syncthreads(); conv_backward_pass<T, RL, 2, TPL, LPB, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, sc); }
        if constexpr (P > 2){ __syncthreads(); conv_backward_pass<T, RL, 1, TPL, LPB, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, sc); }
        if constexpr (P > 1){ __syncthreads(); conv_backward_pass<T, RL, 0, TPL, LPB, SCATTER>(sm, t, j, valid, gout, a.og.stride, tw, sc); }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// contiguous kernel: Stockham auto-sort, LPB lines per CTA, TPL = N / rmax threads per line, padded rows.
// The first pass loads straight from global memory into the butterfly registers (unit stride across the threads of
// a line) and the last pass stores straight from registers (unit stride again: that is what auto-sort buys).
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr unsigned pad_index(unsigned i){ return i + (i >> 3); }

// IN_SMEM: the first pass finds its input in the row (a prologue put it there); OUT_SMEM: the last pass leaves its output
// in the row, natural order (an epilogue takes it from there).  Used by the real-data kernels below.
template<typename T, typename RL, int S, int NS, int TPL, bool BWD, bool SCATTER, bool IN_SMEM = false, bool OUT_SMEM = false, bool NEGATE_UPPER = false>
__device__ __forceinline__ void contig_pass(cplx<T> *row, unsigned j, bool valid, const cplx<T> *gin, cplx<T> *gout,
                                            long long istride, long long ostride, const cplx<T> *tw, T scale, bool do_scale,
                                            scatter_ctx const &sc){
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned NB = RL::N / R;
    constexpr unsigned BPT = NB / TPL;              // butterflies per thread in this pass
    constexpr bool FIRST = (S == 0) && !IN_SMEM, LAST = (S == RL::passes - 1) && !OUT_SMEM;
    cplx<T> v[BPT][R];
    #pragma unroll
    for(unsigned u=0; u<BPT; u++){
        const unsigned q = j + u * TPL;
        if constexpr (FIRST){
            if (valid){
                const cplx<T> *src = gin + static_cast<long long>(q) * istride;
                const long long hop = static_cast<long long>(NB) * istride;
                #pragma unroll
                for(unsigned r=0; r<R; r++){
                    cplx<T> x = *src;
                    src += hop;
                    v[u][r] = BWD ? cswap(x) : x;
                }
            }else{
                #pragma unroll
                for(unsigned r=0; r<R; r++) v[u][r] = mk<T>(0, 0);
            }
        }else{
            #pragma unroll
            for(unsigned r=0; r<R; r++){
                cplx<T> x = row[pad_index(q + r * NB)];
                // DST-II (first pass only): the odd samples fill the upper half of the permuted sequence and carry a minus sign
                if (NEGATE_UPPER && S == 0 && r >= R / 2){ x.x = -x.x; x.y = -x.y; }
                v[u][r] = x;
            }
        }
    }
    if constexpr (!FIRST) __syncthreads();   // every leg has been read before anybody overwrites the row
    #pragma unroll
    for(unsigned u=0; u<BPT; u++){
        const unsigned q = j + u * TPL;
        const unsigned k = q % NS;
        if constexpr (NS > 1) apply_twiddles<T, R, true>(v[u], tw, k * (RL::N / (NS * R)));
        butterfly<T, R>::run(v[u]);
        const unsigned o = (q / NS) * (NS * R) + k;
        if constexpr (!LAST){
            #pragma unroll
            for(unsigned r=0; r<R; r++) row[pad_index(o + r * NS)] = v[u][r];
        }else{
            if (valid){
                if constexpr (SCATTER){
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        cplx<T> x = BWD ? cswap(v[u][r]) : v[u][r];
                        if (do_scale){ x.x *= scale; x.y *= scale; }
                        *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(o + r * NS), sc.a, sc.b) = x;
                    }
                }else{
                    cplx<T> *dst = gout + static_cast<long long>(o) * ostride;
                    const long long hop = static_cast<long long>(NS) * ostride;
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        cplx<T> x = BWD ? cswap(v[u][r]) : v[u][r];
                        if (do_scale){ x.x *= scale; x.y *= scale; }
                        *dst = x;
                        dst += hop;
                    }
                }
            }
        }
    }
}

// One tile of the contiguous kernel: LPB lines starting at line tile * LPB.  TPL threads share a line; it must divide N / R
// for every radix R of the schedule (power-of-two schedules: N / rmax).  `smap`: scatter map already staged in shared memory.
template<typename T, typename RL, int LPB, bool BWD, bool SCATTER, int TPL>
__device__ __forceinline__ void contig_tile(unsigned char *smem_raw, const scatter_map *smap, fft_args const &a, unsigned tile){
    constexpr unsigned PITCH = pad_index(RL::N) + 1;
    const unsigned j = threadIdx.x % TPL, t = threadIdx.x / TPL;
    cplx<T> *row = reinterpret_cast<cplx<T>*>(smem_raw) + t * PITCH;
    const unsigned line = (SCATTER ? scatter_tile_order<LPB>(a, tile) : tile) * LPB + t;
    const bool valid = line < a.nlines;
    const cplx<T> *gin = reinterpret_cast<const cplx<T>*>(a.in) + (valid ? tile_line_offset(a.ig, a.count_a, line) : 0);
    cplx<T> *gout = nullptr;
    scatter_ctx sc{nullptr, 0, 0, 0};
    if constexpr (SCATTER){
        sc.map = smap;
        sc.b = static_cast<int>(line / static_cast<unsigned>(a.count_a));
        sc.a = static_cast<int>(line - static_cast<unsigned>(sc.b) * static_cast<unsigned>(a.count_a));
        sc.row = valid ? scatter_row(smap, sc.a, sc.b) : 0;
    }else{
        gout = reinterpret_cast<cplx<T>*>(a.out) + (valid ? tile_line_offset(a.og, a.count_a, line) : 0);
    }
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    constexpr int P = RL::passes;
    constexpr int N1 = RL::radix(0), N2 = N1 * RL::radix(1), N3 = N2 * RL::radix(2);

    contig_pass<T, RL, 0, 1, TPL, BWD, SCATTER>(row, j, valid, gin, gout, a.ig.stride, a.og.stride, tw, scale, do_scale, sc);
    if constexpr (P > 1){
        __syncthreads();
        contig_pass<T, RL, 1, N1, TPL, BWD, SCATTER>(row, j, valid, gin, gout, a.ig.stride, a.og.stride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 2){
        __syncthreads();
        contig_pass<T, RL, 2, N2, TPL, BWD, SCATTER>(row, j, valid, gin, gout, a.ig.stride, a.og.stride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 3){
        __syncthreads();
        contig_pass<T, RL, 3, N3, TPL, BWD, SCATTER>(row, j, valid, gin, gout, a.ig.stride, a.og.stride, tw, scale, do_scale, sc);
    }
}

template<typename T, typename RL, int LPB, int MINB, bool BWD, bool SCATTER, int TPL = RL::N / RL::rmax>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_contig_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned PITCH = pad_index(RL::N) + 1;
    scatter_map *smap = nullptr;
    if constexpr (SCATTER){
        smap = reinterpret_cast<scatter_map*>(smem_raw + ((sizeof(cplx<T>) * PITCH * LPB + 15) / 16) * 16);
        scatter_stage(smap, a.smap, shift);
        __syncthreads();
    }
    const unsigned ntiles = tile_count<LPB>(a);
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        if constexpr (SCATTER){
            if (a.done_mode == 2) wait_for_planes<LPB>(a, scatter_tile_order<LPB>(a, tile));
            contig_tile<T, RL, LPB, BWD, SCATTER, TPL>(smem_raw, smap, a, tile);
        }else{
            const unsigned at = (a.order_nb > 1) ? round_robin_tile_order<LPB>(a, static_cast<unsigned>(a.order_nb), tile) : tile;
            wait_for_planes<LPB>(a, at);
            contig_tile<T, RL, LPB, BWD, SCATTER, TPL>(smem_raw, smap, a, at);
            report_plane<LPB>(a, at);
        }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// paired kernel: TWO consecutive transforms of the same box -- along the contiguous axis (A) and along the middle axis (B)
// -- in ONE persistent launch, plane by plane, so that B finds the output of A in the L2 cache (126 MB on B200) instead of
// HBM and overwrites it in place before it is ever written back: the pair costs one read and one write of HBM instead of
// two of each (SURVEY 8d counts 2 x 2 x D algorithmic bytes for it).
// The CTAs of the grid (as many as fit the GPU at once) walk a list of tickets dealt round-robin; tickets come
// in groups: the A tiles of plane g, then the B tiles of plane g - lag.  A B tile waits (acquire spin on a per-plane
// counter) until every A tile of its plane has been stored; with a lag of a few planes the wait is never taken, and the
// planes in flight (lag x plane bytes) stay far below the L2 capacity.  No deadlock: a B ticket only waits for A tickets
// with smaller numbers, which are held by running CTAs that wait for nothing.
// B may be a SCATTER variant: the first local pass of a multi-GPU stage then hides behind the NVLink-bound fused stage.
// ---------------------------------------------------------------------------------------------------------
struct pair_args {
    fft_args a, b;                 // A: lines along the contiguous axis, B: lines along the middle axis, both over the whole box
    unsigned planes;               // extent of the slowest axis
    unsigned tiles_a, tiles_b;     // tiles per plane of A and of B
    unsigned lag;                  // planes between the A and the B front
    unsigned *done;                // [planes], zeroed before the launch: tiles of the first transform stored, per plane
};


// shared memory of the paired kernel: the larger of the two tiles, then (scatter variants) the staged map, then the ticket
template<typename T, typename RLA, int LPBA, typename RLB, int LPBB>
__host__ __device__ constexpr size_t pair_tile_bytes(){
    constexpr size_t tile_a = ((sizeof(cplx<T>) * (pad_index(RLA::N) + 1) * LPBA + 15) / 16) * 16;
    constexpr size_t tile_b = sizeof(cplx<T>) * static_cast<size_t>(RLB::N) * LPBB;
    return tile_a > tile_b ? tile_a : tile_b;
}
template<typename T, typename RLA, int LPBA, typename RLB, int LPBB, bool SCATTER>
__host__ __device__ constexpr size_t pair_smem_bytes(){
    return pair_tile_bytes<T, RLA, LPBA, RLB, LPBB>() + (SCATTER ? sizeof(scatter_map) : 0);
}

// CONTIG_FIRST: the contiguous-axis transform runs ahead (forward order of a box), else the middle-axis one (backward order).
// SCATTER applies to the transform that runs second.  p.a always describes the contiguous-axis transform, p.b the middle-axis one.
template<typename T, typename RLA, int LPBA, int TPLA, typename RLB, int TPLB, int LPBB, int MINB, bool BWD, bool SCATTER, bool CONTIG_FIRST>
__global__ void __launch_bounds__(TPLB * LPBB, MINB) fft_pair_kernel(pair_args p){
    static_assert(TPLA * LPBA == TPLB * LPBB, "both phases use the whole CTA");
    B200_DYN_SMEM(smem_raw);
    batch_shift shift_a, shift_b;
    p.a = batch_entry(p.a, shift_a);
    p.b = batch_entry(p.b, shift_b);
    p.done += static_cast<size_t>(blockIdx.y) * p.planes;
    constexpr size_t tile_bytes = pair_tile_bytes<T, RLA, LPBA, RLB, LPBB>();
    scatter_map *smap = nullptr;
    if constexpr (SCATTER){
        smap = reinterpret_cast<scatter_map*>(smem_raw + tile_bytes);
        scatter_stage(smap, CONTIG_FIRST ? p.b.smap : p.a.smap, CONTIG_FIRST ? shift_b : shift_a);
    }
    const unsigned tiles_first = CONTIG_FIRST ? p.tiles_a : p.tiles_b, tiles_second = CONTIG_FIRST ? p.tiles_b : p.tiles_a;
    const unsigned group = tiles_first + tiles_second;
    const unsigned total = (p.planes + p.lag) * group;
    // tickets are dealt round-robin: CTA b takes b, b + gridDim.x, ...  (a shared atomic ticket counter serialises at one L2
    // address: tools/kbench_pair.cu).  Still free of deadlock as long as the whole grid is resident (the host sizes it by
    // occupancy): the CTA that holds the smallest unfinished ticket never waits for anything unfinished.
    for(unsigned ticket = blockIdx.x; ticket < total; ticket += gridDim.x){
        if (ticket != blockIdx.x) __syncthreads();      // the previous tile is done with shared memory
        const unsigned g = ticket / group, r = ticket - g * group;
        if (r < tiles_first){
            if (g >= p.planes) continue;       // the first front has left the box: only tiles of the second transform remain
            if constexpr (CONTIG_FIRST) contig_tile<T, RLA, LPBA, BWD, false, TPLA>(smem_raw, nullptr, p.a, g * p.tiles_a + r);
            else strided_tile<T, RLB, TPLB, LPBB, BWD, false>(reinterpret_cast<cplx<T>*>(smem_raw), nullptr, p.b, g * p.tiles_b + r);
            __syncthreads();                   // every thread has issued its stores
            if (threadIdx.x == 0) count_release(p.done + g);
        }else{
            if (g < p.lag) continue;           // the second front has not entered the box yet
            const unsigned plane = g - p.lag;
            if (threadIdx.x == 0){ while(load_acquire(p.done + plane) < tiles_first){} }
            __syncthreads();
            if constexpr (CONTIG_FIRST) strided_tile<T, RLB, TPLB, LPBB, BWD, SCATTER>(reinterpret_cast<cplx<T>*>(smem_raw), smap, p.b, plane * p.tiles_b + (r - tiles_first));
            else contig_tile<T, RLA, LPBA, BWD, SCATTER, TPLA>(smem_raw, smap, p.a, plane * p.tiles_a + (r - tiles_first));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// real-data variants of the contiguous kernel (power-of-two real length n = 2M, lines contiguous on both sides):
//   real_r2c : forward  real n -> complex n/2+1 ;  backward complex n/2+1 -> real n          (cufftExecD2Z / Z2D)
//   real_cos : forward  REDFT10 (DCT-II)        ;  backward 2*REDFT01 (DCT-III)               (reference cufft_cos)
//   real_sin : forward  RODFT10 (DST-II)        ;  backward 2*RODFT01 (DST-III)               (reference cufft_sin)
// All of them run ONE complex FFT of length M = n/2 per line (the even/odd packing z_j = x_2j + i x_2j+1) between a
// prologue and an epilogue that stay in shared memory:
//   rfft:   E = (Z_k + conj Z_{M-k})/2,  O = -i (Z_k - conj Z_{M-k})/2,  X_k = E + W_n^k O,  X_{M-k} = conj(E - W_n^k O)
//   irfft:  A = X_k + conj X_{M-k},  B = X_k - conj X_{M-k},  Z_k = A + i W_n^{-k} B,  Z_{M-k} = conj(A - i W_n^{-k} B)
//   DCT-II (Makhoul): v = (x_0, x_2, ..., x_3, x_1), V = rfft(v), y_k = 2 Re(w_k V_k), y_{n-k} = -2 Im(w_k V_k), w_k = W_{4n}^k
//   DCT-III: V_k = conj(w_k) (y_k - i y_{n-k}), v = irfft(V), x_{2i} = 2 v_i, x_{2i+1} = 2 v_{n-1-i}
//   DST-II(x)_k = DCT-II((-1)^i x_i)_{n-1-k};  DST-III(x)_i = (-1)^i DCT-III(reversed x)_i
// The reference reaches the r2r transforms through a length-4n r2c FFT and pre/post-processing kernels launched line
// by line (include/heffte_r2r_executor.h:43-180, 191-278; src/heffte_backend_cuda.cu:149-327): 16x the FFT work and
// 2 x 32768 launches per stage at 512^3 / 8 ranks; here a stage is one launch that moves every real number once.
// ---------------------------------------------------------------------------------------------------------
enum real_kind : int { real_r2c = 0, real_cos = 1, real_sin = 2 };

// position of real number p of a line inside a padded complex row (re/im interleaved)
__host__ __device__ constexpr unsigned real_pos(unsigned p){ return 2 * pad_index(p >> 1) + (p & 1); }

// (TPL_: threads per line, see fft_contig_kernel; the first radix of the schedule must be even for the sine kinds)
template<typename T, typename RL, int LPB, int MINB, int KIND, bool BWD, bool SCATTER, int TPL_ = RL::N / RL::rmax>
__global__ void __launch_bounds__(TPL_ * LPB, MINB) fft_contig_real_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned M = RL::N, NR = 2 * RL::N;
    constexpr int TPL = TPL_;
    constexpr unsigned PITCH = pad_index(RL::N) + 1;
    constexpr bool R2C = (KIND == real_r2c);
    constexpr bool PROLOGUE = !(R2C && !BWD);               // everything but the real-to-complex forward load
    constexpr bool EPILOGUE = !(R2C && BWD) || SCATTER;     // everything but the plain complex-to-real store
    constexpr bool REAL_IN = !R2C || !BWD, REAL_OUT = !R2C || BWD;
    const unsigned j = threadIdx.x % TPL, t = threadIdx.x / TPL;
    cplx<T> *row = reinterpret_cast<cplx<T>*>(smem_raw) + t * PITCH;
    T *rrow = reinterpret_cast<T*>(row);
    const unsigned line = (SCATTER ? scatter_tile_order<LPB>(a, blockIdx.x) : blockIdx.x) * LPB + t;
    const bool valid = line < a.nlines;
    const long long ioff = valid ? tile_line_offset(a.ig, a.count_a, line) : 0;
    const T *rin = reinterpret_cast<const T*>(a.in) + (REAL_IN ? ioff : 2 * ioff);          // both views of the input line
    const cplx<T> *cin = reinterpret_cast<const cplx<T>*>(rin);
    T *rout = nullptr;
    scatter_ctx sc{nullptr, 0, 0, 0};
    if constexpr (SCATTER){
        scatter_map *smap = reinterpret_cast<scatter_map*>(smem_raw + ((sizeof(cplx<T>) * PITCH * LPB + 15) / 16) * 16);
        scatter_stage(smap, a.smap, shift);
        __syncthreads();
        sc.map = smap;
        sc.b = static_cast<int>(line / static_cast<unsigned>(a.count_a));
        sc.a = static_cast<int>(line - static_cast<unsigned>(sc.b) * static_cast<unsigned>(a.count_a));
        sc.row = valid ? scatter_row(smap, sc.a, sc.b) : 0;
    }else{
        const long long ooff = valid ? tile_line_offset(a.og, a.count_a, line) : 0;
        rout = reinterpret_cast<T*>(a.out) + (REAL_OUT ? ooff : 2 * ooff);
    }
    cplx<T> *cout = reinterpret_cast<cplx<T>*>(rout);
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const cplx<T> *tx = reinterpret_cast<const cplx<T>*>(a.twiddle2);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;

    // ---- prologue: build the (swapped, for the backward engine) complex input of the M-point transform in the row -------
    // Global memory is read with asynchronous copies straight into shared memory (the whole line in flight at once).
    if constexpr (PROLOGUE){
        if constexpr (!BWD){
            // DCT-II / DST-II: Makhoul permutation of the real line; the sign of the sine variant rides on the first pass
            if (valid){
                #pragma unroll 4
                for(unsigned i = j; i < NR; i += TPL){
                    const unsigned p = (i & 1) ? NR - 1 - (i >> 1) : (i >> 1);
                    async_copy<sizeof(T)>(rrow + real_pos(p), rin + i);
                }
            }
            async_wait_all();
            __syncthreads();
        }else{
            // stage the line as it is (M+1 complex numbers, or n reals), then every thread pulls its pairs (k, M-k) into
            // registers, and only after a barrier writes Z_k and Z_{M-k} over them
            if (valid){
                if constexpr (R2C){
                    #pragma unroll 4
                    for(unsigned k = j; k <= M; k += TPL) async_copy<sizeof(cplx<T>)>(row + pad_index(k), cin + k);
                }else{
                    #pragma unroll 4
                    for(unsigned i = j; i < NR; i += TPL) async_copy<sizeof(T)>(rrow + real_pos(i), rin + i);
                }
            }
            async_wait_all();
            __syncthreads();
            constexpr unsigned KPT = (M / 2 + TPL) / TPL;      // pairs per thread: ceil((M/2 + 1) / TPL)
            cplx<T> vk[KPT], vm[KPT];
            #pragma unroll
            for(unsigned u=0; u<KPT; u++){
                const unsigned k = j + u * TPL;
                if (k <= M / 2){
                    if constexpr (R2C){
                        vk[u] = row[pad_index(k)]; vm[u] = row[pad_index(M - k)];
                        if (k == 0){ vk[u].y = 0; vm[u].y = 0; }      // c2r ignores the imaginary part of the self-conjugate entries
                    }else{
                        T yk, ynk, ymk, ypk;   // y_k, y_{n-k}, y_{M-k}, y_{M+k} of the (reversed, for the sine) input
                        if constexpr (KIND == real_cos){
                            yk = rrow[real_pos(k)]; ynk = (k == 0) ? T(0) : rrow[real_pos(NR - k)];
                            ymk = rrow[real_pos(M - k)]; ypk = rrow[real_pos(M + k)];
                        }else{
                            yk = rrow[real_pos(NR - 1 - k)]; ynk = (k == 0) ? T(0) : rrow[real_pos(k - 1)];
                            ymk = rrow[real_pos(M - 1 + k)]; ypk = rrow[real_pos(M - 1 - k)];
                        }
                        const cplx<T> wk = ldg_c<T>(tx + k), wm = ldg_c<T>(tx + (M - k));
                        vk[u] = cmul(mk<T>(yk, -ynk), mk<T>(wk.x, -wk.y));
                        vm[u] = cmul(mk<T>(ymk, -ypk), mk<T>(wm.x, -wm.y));
                    }
                }
            }
            __syncthreads();
            #pragma unroll
            for(unsigned u=0; u<KPT; u++){
                const unsigned k = j + u * TPL;
                if (k <= M / 2){
                    const cplx<T> A = mk<T>(vk[u].x + vm[u].x, vk[u].y - vm[u].y), B = mk<T>(vk[u].x - vm[u].x, vk[u].y + vm[u].y);
                    const cplx<T> w = ldg_c<T>(tx + 4 * k);
                    const cplx<T> wb = cmul(mk<T>(w.x, -w.y), B);         // W_n^{-k} B
                    const cplx<T> C = mk<T>(-wb.y, wb.x);                 // i W_n^{-k} B
                    row[pad_index(k)] = mk<T>(A.y + C.y, A.x + C.x);                      // swap(A + C)
                    if (k > 0) row[pad_index(M - k)] = mk<T>(-(A.y - C.y), A.x - C.x);    // swap(conj(A - C))
                }
            }
            __syncthreads();
        }
    }

    // ---- the M-point complex transform ---------------------------------------------------------------------------
    constexpr int P = RL::passes;
    constexpr int N1 = RL::radix(0), N2 = N1 * RL::radix(1), N3 = N2 * RL::radix(2);
    // swaps of the backward engine live in the prologue / epilogue when those exist
    contig_pass<T, RL, 0, 1, TPL, BWD, false, PROLOGUE, (EPILOGUE && P == 1), (KIND == real_sin && !BWD)>(row, j, valid, cin, cout, 1, 1, tw, scale, do_scale, sc);
    if constexpr (P > 1){
        __syncthreads();
        contig_pass<T, RL, 1, N1, TPL, BWD, false, false, (EPILOGUE && P == 2)>(row, j, valid, cin, cout, 1, 1, tw, scale, do_scale, sc);
    }
    if constexpr (P > 2){
        __syncthreads();
        contig_pass<T, RL, 2, N2, TPL, BWD, false, false, (EPILOGUE && P == 3)>(row, j, valid, cin, cout, 1, 1, tw, scale, do_scale, sc);
    }
    if constexpr (P > 3){
        __syncthreads();
        contig_pass<T, RL, 3, N3, TPL, BWD, false, false, (EPILOGUE && P == 4)>(row, j, valid, cin, cout, 1, 1, tw, scale, do_scale, sc);
    }

    // ---- epilogue ------------------------------------------------------------------------------------------------
    if constexpr (EPILOGUE){
        __syncthreads();
        if (!valid) return;
        auto put_real = [&](unsigned i, T value){
            if (do_scale) value *= scale;
            if constexpr (SCATTER) *scatter_address<T>(sc.map, sc.row, static_cast<int>(i), sc.a, sc.b) = value;
            else rout[i] = value;
        };
        if constexpr (BWD){
            // the engine ran forward on swapped data: row[e] = (Im z_e, Re z_e) with z_e = v_2e + i v_2e+1
            for(unsigned i = j; i < NR; i += TPL){
                if constexpr (R2C){
                    put_real(i, (i & 1) ? row[pad_index(i >> 1)].x : row[pad_index(i >> 1)].y);
                }else{
                    const unsigned p = (i & 1) ? NR - 1 - (i >> 1) : (i >> 1);
                    T v = (p & 1) ? row[pad_index(p >> 1)].x : row[pad_index(p >> 1)].y;
                    v *= T(2);
                    if (KIND == real_sin && (i & 1)) v = -v;
                    put_real(i, v);
                }
            }
        }else{
            const T half = static_cast<T>(0.5);
            for(unsigned k = j; k <= M / 2; k += TPL){
                const cplx<T> zk = row[pad_index(k)], zm = row[pad_index((M - k) % M)];
                const cplx<T> E = mk<T>((zk.x + zm.x) * half, (zk.y - zm.y) * half);
                const cplx<T> O = mk<T>((zk.y + zm.y) * half, (zm.x - zk.x) * half);      // -i (Z_k - conj Z_{M-k}) / 2
                const cplx<T> Pk = cmul(ldg_c<T>(tx + 4 * k), O);
                const cplx<T> vk = mk<T>(E.x + Pk.x, E.y + Pk.y), vm = mk<T>(E.x - Pk.x, -(E.y - Pk.y));
                if constexpr (R2C){
                    cplx<T> xk = vk, xm = vm;
                    if (do_scale){ xk.x *= scale; xk.y *= scale; xm.x *= scale; xm.y *= scale; }
                    if constexpr (SCATTER){
                        *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(k), sc.a, sc.b) = xk;
                        *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(M - k), sc.a, sc.b) = xm;
                    }else{
                        cout[k] = xk;
                        cout[M - k] = xm;
                    }
                }else{
                    const cplx<T> a1 = cmul(ldg_c<T>(tx + k), vk), a2 = cmul(ldg_c<T>(tx + (M - k)), vm);
                    // y_k, y_{n-k}, y_{M-k}, y_{M+k}; the sine transform stores y_p at n-1-p
                    auto put_y = [&](unsigned p, T value){ put_real((KIND == real_sin) ? NR - 1 - p : p, value); };
                    put_y(k, T(2) * a1.x);
                    if (k > 0) put_y(NR - k, T(-2) * a1.y);
                    put_y(M - k, T(2) * a2.x);
                    if (k > 0) put_y(M + k, T(-2) * a2.y);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// DCT / DST along contiguous lines, second generation (plain stores only): the Makhoul permutation never touches shared
// memory.  With z_e = v_2e + i v_2e+1 and v = (x_0, x_2, ..., x_3, x_1), FOUR consecutive reals x_4i .. x_4i+3 are exactly
//      z_i = (x_4i, x_4i+2)          and          z_{M-1-i} = (x_4i+3, x_4i+1)                     (i < M/2)
// and in a Stockham pass whose legs are q + r NB the mirror index M-1-(q + r NB) is leg R-1-r of butterfly NB-1-q.  A thread
// that owns the butterflies q and NB-1-q therefore moves whole 32-byte (fp64) / 16-byte (fp32) pieces of the line between
// global memory and its registers: the forward transform loads its first pass that way, the backward transform stores its
// last pass that way.  No 8-byte asynchronous copies, no bank conflicts, one shared-memory round trip less than
// fft_contig_real_kernel.  Needs an even number of butterflies per thread in that pass and 16-byte aligned lines.
// ---------------------------------------------------------------------------------------------------------
template<typename T> struct quad { T c[4]; };
template<typename T> __device__ __forceinline__ quad<T> load_quad(const T *p);
template<> __device__ __forceinline__ quad<double> load_quad<double>(const double *p){
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    return quad<double>{{a.x, a.y, b.x, b.y}};
}
template<> __device__ __forceinline__ quad<float> load_quad<float>(const float *p){
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    return quad<float>{{a.x, a.y, a.z, a.w}};
}
template<typename T> __device__ __forceinline__ void store_quad(T *p, T c0, T c1, T c2, T c3);
template<> __device__ __forceinline__ void store_quad<double>(double *p, double c0, double c1, double c2, double c3){
    double2 a; a.x = c0; a.y = c1; double2 b; b.x = c2; b.y = c3;
    reinterpret_cast<double2*>(p)[0] = a; reinterpret_cast<double2*>(p)[1] = b;
}
template<> __device__ __forceinline__ void store_quad<float>(float *p, float c0, float c1, float c2, float c3){
    float4 a; a.x = c0; a.y = c1; a.z = c2; a.w = c3;
    reinterpret_cast<float4*>(p)[0] = a;
}

template<typename T, typename RL, int LPB, int MINB, int KIND, bool BWD, int TPL_>
__global__ void __launch_bounds__(TPL_ * LPB, MINB) fft_contig_dct_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    static_assert(KIND == real_cos || KIND == real_sin, "cosine / sine transforms only");
    static_assert(RL::passes >= 2, "needs at least two passes");
    constexpr unsigned M = RL::N, NR = 2 * RL::N;
    constexpr int TPL = TPL_;
    constexpr int P = RL::passes;
    constexpr unsigned PITCH = pad_index(RL::N) + 1;
    constexpr int N1 = RL::radix(0), N2 = N1 * RL::radix(1), N3 = N2 * RL::radix(2);
    const unsigned j = threadIdx.x % TPL, t = threadIdx.x / TPL;
    cplx<T> *row = reinterpret_cast<cplx<T>*>(smem_raw) + t * PITCH;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const cplx<T> *tx = reinterpret_cast<const cplx<T>*>(a.twiddle2);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    const scatter_ctx sc{nullptr, 0, 0, 0};
    const unsigned ntiles = tile_count<LPB>(a);
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        const unsigned line = tile * LPB + t;
        const bool valid = line < a.nlines;
        const T *rin = reinterpret_cast<const T*>(a.in) + (valid ? tile_line_offset(a.ig, a.count_a, line) : 0);
        T *rout = reinterpret_cast<T*>(a.out) + (valid ? tile_line_offset(a.og, a.count_a, line) : 0);

        if constexpr (!BWD){
            // ---- first pass: butterflies q and NB-1-q, inputs straight from global memory in pieces of four reals ------------
            {
                constexpr unsigned R = RL::radix(0), NB = M / R, BPT = NB / TPL, H = BPT / 2;
                static_assert(BPT % 2 == 0 && R % 2 == 0, "the forward transform pairs the butterflies of its first pass");
                cplx<T> v[BPT][R];
                #pragma unroll
                for(unsigned u=0; u<H; u++){
                    const unsigned q = j + u * TPL;
                    #pragma unroll
                    for(unsigned r=0; r<R/2; r++){
                        quad<T> c, d;
                        if (valid){ c = load_quad<T>(rin + 4 * (q + r * NB)); d = load_quad<T>(rin + 4 * ((NB - 1 - q) + r * NB)); }
                        else{ c = quad<T>{{0, 0, 0, 0}}; d = c; }
                        if (KIND == real_sin){ c.c[1] = -c.c[1]; c.c[3] = -c.c[3]; d.c[1] = -d.c[1]; d.c[3] = -d.c[3]; }
                        v[u][r] = mk<T>(c.c[0], c.c[2]);       v[u + H][R - 1 - r] = mk<T>(c.c[3], c.c[1]);
                        v[u + H][r] = mk<T>(d.c[0], d.c[2]);   v[u][R - 1 - r] = mk<T>(d.c[3], d.c[1]);
                    }
                }
                #pragma unroll
                for(unsigned u=0; u<BPT; u++){
                    const unsigned q = (u < H) ? (j + u * TPL) : (NB - 1 - (j + (u - H) * TPL));
                    butterfly<T, R>::run(v[u]);
                    #pragma unroll
                    for(unsigned r=0; r<R; r++) row[pad_index(q * R + r)] = v[u][r];
                }
            }
            __syncthreads();
            contig_pass<T, RL, 1, N1, TPL, false, false, false, (P == 2)>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            if constexpr (P > 2){
                __syncthreads();
                contig_pass<T, RL, 2, N2, TPL, false, false, false, (P == 3)>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            }
            if constexpr (P > 3){
                __syncthreads();
                contig_pass<T, RL, 3, N3, TPL, false, false, false, true>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            }
            __syncthreads();
            // ---- epilogue: spectrum of the real sequence from the pair (k, M-k), then the quarter-wave twiddle ------------------
            if (valid){
                const T half = static_cast<T>(0.5);
                for(unsigned k = j; k <= M / 2; k += TPL){
                    const cplx<T> zk = row[pad_index(k)], zm = row[pad_index((M - k) % M)];
                    const cplx<T> E = mk<T>((zk.x + zm.x) * half, (zk.y - zm.y) * half);
                    const cplx<T> O = mk<T>((zk.y + zm.y) * half, (zm.x - zk.x) * half);
                    const cplx<T> Pk = cmul(ldg_c<T>(tx + 4 * k), O);
                    const cplx<T> vk = mk<T>(E.x + Pk.x, E.y + Pk.y), vm = mk<T>(E.x - Pk.x, -(E.y - Pk.y));
                    const cplx<T> a1 = cmul(ldg_c<T>(tx + k), vk), a2 = cmul(ldg_c<T>(tx + (M - k)), vm);
                    const T s2 = do_scale ? T(2) * scale : T(2);
                    auto put_y = [&](unsigned p, T value){ rout[(KIND == real_sin) ? NR - 1 - p : p] = value; };
                    put_y(k, s2 * a1.x);
                    if (k > 0) put_y(NR - k, -s2 * a1.y);
                    put_y(M - k, s2 * a2.x);
                    if (k > 0) put_y(M + k, -s2 * a2.y);
                }
            }
        }else{
            // ---- prologue: Z_k and Z_{M-k} from four reals fetched straight from global memory, written swapped ------------------
            if (valid){
                for(unsigned k = j; k <= M / 2; k += TPL){
                    T yk, ynk, ymk, ypk;
                    if constexpr (KIND == real_cos){
                        yk = rin[k]; ynk = (k == 0) ? T(0) : rin[NR - k]; ymk = rin[M - k]; ypk = rin[M + k];
                    }else{
                        yk = rin[NR - 1 - k]; ynk = (k == 0) ? T(0) : rin[k - 1]; ymk = rin[M - 1 + k]; ypk = rin[M - 1 - k];
                    }
                    const cplx<T> wk = ldg_c<T>(tx + k), wm = ldg_c<T>(tx + (M - k));
                    const cplx<T> vk = cmul(mk<T>(yk, -ynk), mk<T>(wk.x, -wk.y));
                    const cplx<T> vm = cmul(mk<T>(ymk, -ypk), mk<T>(wm.x, -wm.y));
                    const cplx<T> A = mk<T>(vk.x + vm.x, vk.y - vm.y), B = mk<T>(vk.x - vm.x, vk.y + vm.y);
                    const cplx<T> w = ldg_c<T>(tx + 4 * k);
                    const cplx<T> wb = cmul(mk<T>(w.x, -w.y), B);
                    const cplx<T> C = mk<T>(-wb.y, wb.x);
                    row[pad_index(k)] = mk<T>(A.y + C.y, A.x + C.x);
                    if (k > 0) row[pad_index(M - k)] = mk<T>(-(A.y - C.y), A.x - C.x);
                }
            }
            __syncthreads();
            contig_pass<T, RL, 0, 1, TPL, true, false, true, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            if constexpr (P > 2){
                __syncthreads();
                contig_pass<T, RL, 1, N1, TPL, true, false, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            }
            if constexpr (P > 3){
                __syncthreads();
                contig_pass<T, RL, 2, N2, TPL, true, false, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc);
            }
            __syncthreads();
            // ---- last pass: butterflies q and NB-1-q, four reals per store straight from the registers ------------------------------
            {
                constexpr unsigned S = P - 1;
                constexpr unsigned R = RL::radix(S), NB = M / R, BPT = NB / TPL, H = BPT / 2;
                static_assert(BPT % 2 == 0 && R % 2 == 0, "the backward transform pairs the butterflies of its last pass");
                cplx<T> v[BPT][R];
                #pragma unroll
                for(unsigned u=0; u<BPT; u++){
                    const unsigned q = (u < H) ? (j + u * TPL) : (NB - 1 - (j + (u - H) * TPL));
                    #pragma unroll
                    for(unsigned r=0; r<R; r++) v[u][r] = row[pad_index(q + r * NB)];
                    apply_twiddles<T, R, true>(v[u], tw, q);          // NS = NB in the last pass: W_N^(q r)
                    butterfly<T, R>::run(v[u]);
                }
                if (valid){
                    const T two = do_scale ? T(2) * scale : T(2);
                    const T odd = (KIND == real_sin) ? -two : two;
                    #pragma unroll
                    for(unsigned u=0; u<H; u++){
                        const unsigned q = j + u * TPL;
                        #pragma unroll
                        for(unsigned r=0; r<R/2; r++){
                            // the engine ran forward on swapped data: (.x, .y) = (Im z, Re z);  z_i = v[u][r], z_{M-1-i} = v[u+H][R-1-r]
                            const cplx<T> zi = v[u][r], zm = v[u + H][R - 1 - r];
                            store_quad<T>(rout + 4 * (q + r * NB), two * zi.y, odd * zm.x, two * zi.x, odd * zm.y);
                            const cplx<T> wi = v[u + H][r], wm = v[u][R - 1 - r];
                            store_quad<T>(rout + 4 * ((NB - 1 - q) + r * NB), two * wi.y, odd * wm.x, two * wi.x, odd * wm.y);
                        }
                    }
                }
            }
        }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// real-data variants of the strided kernel: same transforms and same algebra as fft_contig_real_kernel, for lines whose
// neighbours are adjacent in memory (a real transform along the middle / slow axis of a box, i.e. every r2r stage but
// the first of a plan that does not reorder, and r2c along dimensions 1 / 2).  The tile is [M][LPB] complex numbers;
// a row of LPB adjacent reals is one 128-byte access.  Forward: the real rows land (asynchronously) in the re/im slots
// the even/odd packing wants, decimation-in-frequency passes in place, and the rfft / DCT post-processing reads the
// pair (k, M-k) at its digit-reversed positions.  Backward: the pair-wise pre-processing writes the tile, and the last
// pass stores real rows straight from registers.
// ---------------------------------------------------------------------------------------------------------
// in-place position of natural index k after all DIF passes (inverse of dif_output_index)
template<typename RL>
__device__ __forceinline__ unsigned dif_position_of(unsigned k){
    constexpr unsigned r0 = RL::radix(0);
    unsigned p = (k % r0) * RL::stride(0);
    k /= r0;
    if constexpr (RL::passes > 1){ constexpr unsigned r1 = RL::radix(1); p += (k % r1) * RL::stride(1); k /= r1; }
    if constexpr (RL::passes > 2){ constexpr unsigned r2 = RL::radix(2); p += (k % r2) * RL::stride(2); k /= r2; }
    if constexpr (RL::passes > 3){ p += k; }
    return p;
}

// where real number p of the engine output goes, and what it is multiplied by, on the backward store
template<typename T, int KIND>
__device__ __forceinline__ void real_backward_target(unsigned p, unsigned M, unsigned &row, T &factor){
    if constexpr (KIND == real_r2c){ row = p; factor = T(1); }
    else{
        row = (p < M) ? 2 * p : 2 * (2 * M - 1 - p) + 1;
        factor = (KIND == real_sin && (row & 1)) ? T(-2) : T(2);
    }
}

template<typename T, typename RL, int S, int TPL, int LPB, int KIND, bool BWD, bool SCATTER>
__device__ __forceinline__ void strided_real_pass(cplx<T> *sm, unsigned t, unsigned j, bool valid, T *rout, long long ostride,
                                                  const cplx<T> *tw, T scale, bool do_scale, scatter_ctx const &sc){
    constexpr unsigned R = RL::radix(S);
    constexpr unsigned ST = RL::stride(S);
    constexpr unsigned NB = RL::N / R;
    constexpr unsigned M = RL::N;
    constexpr bool FIRST = (S == 0), LAST = (S == RL::passes - 1);
    #pragma unroll
    for(unsigned u=0; u<NB/TPL; u++){
        const unsigned q = j + u * TPL;
        const unsigned o = q % ST;
        const unsigned p0 = (q / ST) * (ST * R) + o;
        cplx<T> *cell = sm + p0 * LPB + t;
        cplx<T> v[R];
        #pragma unroll
        for(unsigned r=0; r<R; r++){
            cplx<T> x = cell[r * ST * LPB];
            // DST-II: the odd samples sit in the second half of the permuted sequence and carry a minus sign
            if (FIRST && !BWD && KIND == real_sin && r >= R / 2){ x.x = -x.x; x.y = -x.y; }
            v[r] = x;
        }
        butterfly<T, R>::run(v);
        if constexpr (!LAST || !BWD){
            if constexpr (!LAST) apply_twiddles<T, R, true>(v, tw, o * (RL::N / (ST * R)));
            #pragma unroll
            for(unsigned r=0; r<R; r++) cell[r * ST * LPB] = v[r];
        }else{
            if (valid){
                const unsigned k0 = dif_output_index<RL>(p0);
                #pragma unroll
                for(unsigned r=0; r<R; r++){
                    const unsigned e = k0 + r * (M / R);        // engine output index: z_e = v_2e + i v_2e+1, stored swapped
                    #pragma unroll
                    for(unsigned h=0; h<2; h++){
                        unsigned row; T factor;
                        real_backward_target<T, KIND>(2 * e + h, M, row, factor);
                        T value = (h == 0 ? v[r].y : v[r].x) * factor;
                        if (do_scale) value *= scale;
                        if constexpr (SCATTER) *scatter_address<T>(sc.map, sc.row, static_cast<int>(row), sc.a, sc.b) = value;
                        else rout[static_cast<long long>(row) * ostride] = value;
                    }
                }
            }
        }
    }
}

template<typename T, typename RL, int TPL, int LPB, int MINB, int KIND, bool BWD, bool SCATTER>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_real_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned M = RL::N, NR = 2 * RL::N;
    constexpr bool R2C = (KIND == real_r2c);
    cplx<T> *sm = reinterpret_cast<cplx<T>*>(smem_raw);
    T *rsm = reinterpret_cast<T*>(smem_raw);
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const unsigned line = (SCATTER ? scatter_tile_order<LPB>(a, blockIdx.x) : blockIdx.x) * LPB + t;
    const bool valid = line < a.nlines;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const cplx<T> *tx = reinterpret_cast<const cplx<T>*>(a.twiddle2);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    constexpr int P = RL::passes;
    const long long ioff = valid ? tile_line_offset(a.ig, a.count_a, line) : 0;
    const long long istride = a.ig.stride;

    scatter_ctx sc{nullptr, 0, 0, 0};
    long long ooff = 0;
    if constexpr (SCATTER){
        scatter_map *smap = reinterpret_cast<scatter_map*>(sm + static_cast<size_t>(RL::N) * LPB);   // behind the tile
        scatter_stage(smap, a.smap, shift);
        sc.map = smap;
        sc.b = static_cast<int>(line / static_cast<unsigned>(a.count_a));
        sc.a = static_cast<int>(line - static_cast<unsigned>(sc.b) * static_cast<unsigned>(a.count_a));
    }else{
        ooff = valid ? tile_line_offset(a.og, a.count_a, line) : 0;
    }
    const long long ostride = a.og.stride;

    if constexpr (!BWD){
        // ---- forward: real rows -> re/im slots of the packed sequence (asynchronous, nothing held in registers) ----------
        if (valid){
            const T *src = reinterpret_cast<const T*>(a.in) + ioff + static_cast<long long>(j) * istride;
            const long long hop = static_cast<long long>(TPL) * istride;
            #pragma unroll 4
            for(unsigned i = j; i < NR; i += TPL){
                const unsigned p = R2C ? i : ((i & 1) ? NR - 1 - (i >> 1) : (i >> 1));
                async_copy<sizeof(T)>(rsm + (2 * ((p >> 1) * LPB + t) + (p & 1)), src);
                src += hop;
            }
        }
        async_wait_all();
        __syncthreads();
        if constexpr (SCATTER) sc.row = valid ? scatter_row(sc.map, sc.a, sc.b) : 0;
    }else{
        // ---- backward: Z_k and Z_{M-k} from the pair (k, M-k) of the input, written swapped for the forward engine ---------
        if constexpr (SCATTER){ __syncthreads(); sc.row = valid ? scatter_row(sc.map, sc.a, sc.b) : 0; }
        if (valid){
            const T *rin = reinterpret_cast<const T*>(a.in) + (R2C ? 2 * ioff : ioff);
            const cplx<T> *cin = reinterpret_cast<const cplx<T>*>(rin);
            for(unsigned k = j; k <= M / 2; k += TPL){
                cplx<T> vk, vm;
                if constexpr (R2C){
                    vk = cin[static_cast<long long>(k) * istride]; vm = cin[static_cast<long long>(M - k) * istride];
                    if (k == 0){ vk.y = 0; vm.y = 0; }
                }else{
                    T yk, ynk, ymk, ypk;
                    if constexpr (KIND == real_cos){
                        yk = rin[k * istride]; ynk = (k == 0) ? T(0) : rin[(NR - k) * istride];
                        ymk = rin[(M - k) * istride]; ypk = rin[(M + k) * istride];
                    }else{
                        yk = rin[(NR - 1 - k) * istride]; ynk = (k == 0) ? T(0) : rin[(k - 1) * istride];
                        ymk = rin[(M - 1 + k) * istride]; ypk = rin[(M - 1 - k) * istride];
                    }
                    const cplx<T> wk = ldg_c<T>(tx + k), wm = ldg_c<T>(tx + (M - k));
                    vk = cmul(mk<T>(yk, -ynk), mk<T>(wk.x, -wk.y));
                    vm = cmul(mk<T>(ymk, -ypk), mk<T>(wm.x, -wm.y));
                }
                const cplx<T> A = mk<T>(vk.x + vm.x, vk.y - vm.y), B = mk<T>(vk.x - vm.x, vk.y + vm.y);
                const cplx<T> w = ldg_c<T>(tx + 4 * k);
                const cplx<T> wb = cmul(mk<T>(w.x, -w.y), B);
                const cplx<T> C = mk<T>(-wb.y, wb.x);
                sm[k * LPB + t] = mk<T>(A.y + C.y, A.x + C.x);
                if (k > 0) sm[(M - k) * LPB + t] = mk<T>(-(A.y - C.y), A.x - C.x);
            }
        }
        __syncthreads();
    }

    T *rout = SCATTER ? nullptr : reinterpret_cast<T*>(a.out) + ((R2C && !BWD) ? 2 * ooff : ooff);
    strided_real_pass<T, RL, 0, TPL, LPB, KIND, BWD, SCATTER>(sm, t, j, valid, rout, ostride, tw, scale, do_scale, sc);
    if constexpr (P > 1){
        __syncthreads();
        strided_real_pass<T, RL, 1, TPL, LPB, KIND, BWD, SCATTER>(sm, t, j, valid, rout, ostride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 2){
        __syncthreads();
        strided_real_pass<T, RL, 2, TPL, LPB, KIND, BWD, SCATTER>(sm, t, j, valid, rout, ostride, tw, scale, do_scale, sc);
    }
    if constexpr (P > 3){
        __syncthreads();
        strided_real_pass<T, RL, 3, TPL, LPB, KIND, BWD, SCATTER>(sm, t, j, valid, rout, ostride, tw, scale, do_scale, sc);
    }

    if constexpr (!BWD){
        // ---- forward epilogue: spectrum of the real sequence from the pair (k, M-k), then r2c store or DCT / DST twiddle ------
        __syncthreads();
        if (!valid) return;
        cplx<T> *cout = reinterpret_cast<cplx<T>*>(rout);
        auto put_real = [&](unsigned i, T value){
            if (do_scale) value *= scale;
            if constexpr (SCATTER) *scatter_address<T>(sc.map, sc.row, static_cast<int>(i), sc.a, sc.b) = value;
            else rout[static_cast<long long>(i) * ostride] = value;
        };
        const T half = static_cast<T>(0.5);
        for(unsigned k = j; k <= M / 2; k += TPL){
            const cplx<T> zk = sm[dif_position_of<RL>(k) * LPB + t], zm = sm[dif_position_of<RL>((M - k) % M) * LPB + t];
            const cplx<T> E = mk<T>((zk.x + zm.x) * half, (zk.y - zm.y) * half);
            const cplx<T> O = mk<T>((zk.y + zm.y) * half, (zm.x - zk.x) * half);
            const cplx<T> Pk = cmul(ldg_c<T>(tx + 4 * k), O);
            const cplx<T> vk = mk<T>(E.x + Pk.x, E.y + Pk.y), vm = mk<T>(E.x - Pk.x, -(E.y - Pk.y));
            if constexpr (R2C){
                cplx<T> xk = vk, xm = vm;
                if (do_scale){ xk.x *= scale; xk.y *= scale; xm.x *= scale; xm.y *= scale; }
                if constexpr (SCATTER){
                    *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(k), sc.a, sc.b) = xk;
                    *scatter_address<cplx<T>>(sc.map, sc.row, static_cast<int>(M - k), sc.a, sc.b) = xm;
                }else{
                    cout[static_cast<long long>(k) * ostride] = xk;
                    cout[static_cast<long long>(M - k) * ostride] = xm;
                }
            }else{
                const cplx<T> a1 = cmul(ldg_c<T>(tx + k), vk), a2 = cmul(ldg_c<T>(tx + (M - k)), vm);
                auto put_y = [&](unsigned p, T value){ put_real((KIND == real_sin) ? NR - 1 - p : p, value); };
                put_y(k, T(2) * a1.x);
                if (k > 0) put_y(NR - k, T(-2) * a1.y);
                put_y(M - k, T(2) * a2.x);
                if (k > 0) put_y(M + k, T(-2) * a2.y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// real-data transforms along contiguous lines, second generation (plain stores): two neighbouring lines x1, x2 form ONE complex
// line c = x1 + i x2 (see fft_strided_real2_kernel for the algebra) and the complex Stockham passes run at the full length n.
// Only the passes in between touch shared memory: the schedule starts and ends with the same radix R and a thread owns the
// butterflies q and n/R - q of the first and of the last pass, so that every pair (k, n-k) of the spectrum sits in the
// registers of ONE thread --
//   forward:  the first pass gathers its legs straight from global memory (cosine / sine: in Makhoul's order, the sign of the
//             sine transform on the way); the last pass separates the two lines from its pairs and stores them: y_k and
//             y_{n-k} (quarter-wave twiddle), or the two half spectra of r2c, unit stride across the threads;
//   backward: the first pass builds C_k and C_{n-k} from unit-stride loads of both lines; the last pass stores the two real
//             lines from its registers (cosine / sine: in Makhoul's order).
// Thread 0 of a line owns q = 0 and q = n/2R, whose pairs lie inside one butterfly (and the self-conjugate entries 0, n/2).
// ---------------------------------------------------------------------------------------------------------
template<typename T> __device__ __forceinline__ void pair_split(cplx<T> ck, cplx<T> cm, cplx<T> &v1, cplx<T> &v2){
    const T half = static_cast<T>(0.5);          // V1_k = (C_k + conj C_{n-k}) / 2,   V2_k = -i (C_k - conj C_{n-k}) / 2
    v1 = mk<T>((ck.x + cm.x) * half, (ck.y - cm.y) * half);
    v2 = mk<T>((ck.y + cm.y) * half, (cm.x - ck.x) * half);
}
template<typename T> __device__ __forceinline__ void pair_merge(cplx<T> v1, cplx<T> v2, cplx<T> &ck, cplx<T> &cm){
    ck = mk<T>(v1.x - v2.y, v1.y + v2.x);        // C_k = V1_k + i V2_k,   C_{n-k} = conj(V1_k) + i conj(V2_k)
    cm = mk<T>(v1.x + v2.y, v2.x - v1.y);
}

template<typename T, typename RL, int LPB, int MINB, int KIND, bool BWD>
__global__ void __launch_bounds__((RL::N / (2 * RL::radix(0))) * LPB, MINB) fft_contig_real2_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned N = RL::N;
    constexpr int P = RL::passes;
    constexpr unsigned R = RL::radix(0), NB = N / R;
    constexpr int TPL = NB / 2;
    static_assert(P >= 2 && RL::radix(P - 1) == R && R % 2 == 0, "the schedule starts and ends with the same even radix");
    constexpr bool R2C = (KIND == real_r2c);
    constexpr unsigned PITCH = pad_index(N) + 1;
    constexpr int N1 = RL::radix(0), N2 = N1 * RL::radix(1);
    const unsigned j = threadIdx.x % TPL, t = threadIdx.x / TPL;
    cplx<T> *row = reinterpret_cast<cplx<T>*>(smem_raw) + t * PITCH;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle0);
    const cplx<T> *tx = reinterpret_cast<const cplx<T>*>(a.twiddle2);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    const T factor = (R2C ? T(1) : T(2)) * (do_scale ? scale : T(1));
    const scatter_ctx sc{nullptr, 0, 0, 0};
    const unsigned npairs = static_cast<unsigned>(a.nlines / 2);
    const unsigned ntiles = (npairs + LPB - 1) / LPB;
    const unsigned qa = j, qb = (j == 0) ? NB / 2 : NB - j;        // the two butterflies of this thread, first and last pass
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        const unsigned pair = tile * LPB + t;
        const bool valid = pair < npairs;
        const long long i1 = valid ? tile_line_offset(a.ig, a.count_a, 2 * pair) : 0, i2 = valid ? tile_line_offset(a.ig, a.count_a, 2 * pair + 1) : 0;
        const long long o1 = valid ? tile_line_offset(a.og, a.count_a, 2 * pair) : 0, o2 = valid ? tile_line_offset(a.og, a.count_a, 2 * pair + 1) : 0;
        cplx<T> v[2][R];
        // f(C_k slot, C_{n-k} slot, k, self): every pair this thread owns, k <= n/2; `self`: k is 0 or n/2 and the two slots are one
        auto for_each_pair = [&](auto &&f){
            if (j != 0){
                #pragma unroll
                for(unsigned r=0; r<R/2; r++){
                    f(v[0][r], v[1][R - 1 - r], qa + r * NB, false);
                    f(v[1][r], v[0][R - 1 - r], qb + r * NB, false);
                }
            }else{
                f(v[0][0], v[0][0], 0u, true);
                f(v[0][R / 2], v[0][R / 2], N / 2, true);
                #pragma unroll
                for(unsigned r=1; r<R/2; r++) f(v[0][r], v[0][R - r], r * NB, false);
                #pragma unroll
                for(unsigned r=0; r<R/2; r++) f(v[1][r], v[1][R - 1 - r], NB / 2 + r * NB, false);
            }
        };
        // first pass of a Stockham schedule: no twiddles, butterfly q writes positions q R + r
        auto first_pass_out = [&](){
            #pragma unroll
            for(unsigned u=0; u<2; u++){
                const unsigned q = u ? qb : qa;
                butterfly<T, R>::run(v[u]);
                #pragma unroll
                for(unsigned r=0; r<R; r++) row[pad_index(q * R + r)] = v[u][r];
            }
        };
        // last pass: legs q + r NB from the row, twiddle W_n^(q r), butterfly; v[u][r] is then entry q + r NB of the result
        auto last_pass_in = [&](){
            #pragma unroll
            for(unsigned u=0; u<2; u++){
                const unsigned q = u ? qb : qa;
                #pragma unroll
                for(unsigned r=0; r<R; r++) v[u][r] = row[pad_index(q + r * NB)];
                apply_twiddles<T, R, true>(v[u], tw, static_cast<int>(q));
                butterfly<T, R>::run(v[u]);
            }
        };

        if constexpr (!BWD){
            // ---- first pass: legs straight from the two lines ---------------------------------------------------------------------------------
            {
                const T *in1 = reinterpret_cast<const T*>(a.in) + i1, *in2 = reinterpret_cast<const T*>(a.in) + i2;
                #pragma unroll
                for(unsigned u=0; u<2; u++){
                    const unsigned q = u ? qb : qa;
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        const unsigned p = q + r * NB;                                          // position in the (permuted) sequence
                        unsigned i = p;
                        bool upper = false;
                        if constexpr (!R2C){ upper = 2 * p >= N; i = upper ? 2 * (N - 1 - p) + 1 : 2 * p; }
                        T x1 = valid ? in1[i] : T(0), x2 = valid ? in2[i] : T(0);
                        if (KIND == real_sin && upper){ x1 = -x1; x2 = -x2; }
                        v[u][r] = mk<T>(x1, x2);
                    }
                }
                first_pass_out();
            }
            if constexpr (P > 2){ __syncthreads(); contig_pass<T, RL, 1, N1, TPL, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc); }
            if constexpr (P > 3){ __syncthreads(); contig_pass<T, RL, 2, N2, TPL, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc); }
            __syncthreads();
            last_pass_in();
            // ---- the two lines apart, straight from the registers ----------------------------------------------------------------------------
            if (valid){
                if constexpr (R2C){
                    cplx<T> *out1 = reinterpret_cast<cplx<T>*>(a.out) + o1, *out2 = reinterpret_cast<cplx<T>*>(a.out) + o2;
                    for_each_pair([&](cplx<T> &ck, cplx<T> &cm, unsigned k, bool){
                        cplx<T> v1, v2;
                        pair_split<T>(ck, cm, v1, v2);
                        out1[k] = mk<T>(v1.x * factor, v1.y * factor);
                        out2[k] = mk<T>(v2.x * factor, v2.y * factor);
                    });
                }else{
                    T *out1 = reinterpret_cast<T*>(a.out) + o1, *out2 = reinterpret_cast<T*>(a.out) + o2;
                    for_each_pair([&](cplx<T> &ck, cplx<T> &cm, unsigned k, bool self){
                        cplx<T> v1, v2;
                        pair_split<T>(ck, cm, v1, v2);
                        const cplx<T> w = ldg_c<T>(tx + k);
                        const cplx<T> z1 = cmul(w, v1), z2 = cmul(w, v2);
                        const unsigned lo = (KIND == real_sin) ? N - 1 - k : k;
                        out1[lo] = factor * z1.x; out2[lo] = factor * z2.x;
                        if (!self){
                            const unsigned hi = (KIND == real_sin) ? k - 1 : N - k;
                            out1[hi] = -factor * z1.y; out2[hi] = -factor * z2.y;
                        }
                    });
                }
            }
        }else{
            // ---- first pass: C_k and C_{n-k} from unit-stride loads of both lines, swapped for the forward engine ----------------------------
            if (valid){
                if constexpr (R2C){
                    const cplx<T> *in1 = reinterpret_cast<const cplx<T>*>(a.in) + i1, *in2 = reinterpret_cast<const cplx<T>*>(a.in) + i2;
                    for_each_pair([&](cplx<T> &ck, cplx<T> &cm, unsigned k, bool self){
                        cplx<T> v1 = in1[k], v2 = in2[k];
                        if (self){ v1.y = 0; v2.y = 0; }                 // c2r ignores the imaginary part of the self-conjugate entries
                        cplx<T> c, m;
                        pair_merge<T>(v1, v2, c, m);
                        ck = cswap(c);
                        if (!self) cm = cswap(m);
                    });
                }else{
                    const T *in1 = reinterpret_cast<const T*>(a.in) + i1, *in2 = reinterpret_cast<const T*>(a.in) + i2;
                    for_each_pair([&](cplx<T> &ck, cplx<T> &cm, unsigned k, bool self){
                        // V_k = conj(w_k) (y_k - i y_{n-k}), y_n := 0
                        const unsigned lo = (KIND == real_sin) ? N - 1 - k : k, hi = (KIND == real_sin) ? k - 1 : N - k;
                        const T y1k = in1[lo], y2k = in2[lo];
                        const T y1m = (k == 0) ? T(0) : in1[hi], y2m = (k == 0) ? T(0) : in2[hi];
                        const cplx<T> w = ldg_c<T>(tx + k);
                        const cplx<T> v1 = cmul(mk<T>(y1k, -y1m), mk<T>(w.x, -w.y));
                        const cplx<T> v2 = cmul(mk<T>(y2k, -y2m), mk<T>(w.x, -w.y));
                        cplx<T> c, m;
                        pair_merge<T>(v1, v2, c, m);
                        ck = cswap(c);
                        if (!self) cm = cswap(m);
                    });
                }
            }else{
                #pragma unroll
                for(unsigned r=0; r<R; r++){ v[0][r] = mk<T>(0, 0); v[1][r] = mk<T>(0, 0); }
            }
            first_pass_out();
            if constexpr (P > 2){ __syncthreads(); contig_pass<T, RL, 1, N1, TPL, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc); }
            if constexpr (P > 3){ __syncthreads(); contig_pass<T, RL, 2, N2, TPL, false, false>(row, j, valid, nullptr, nullptr, 1, 1, tw, scale, do_scale, sc); }
            __syncthreads();
            last_pass_in();
            // ---- entry p of the result (swapped: .y is line 1, .x is line 2); cosine / sine: x_2e = 2 v_e, x_2e+1 = 2 v_{n-1-e} --------------
            if (valid){
                T *out1 = reinterpret_cast<T*>(a.out) + o1, *out2 = reinterpret_cast<T*>(a.out) + o2;
                #pragma unroll
                for(unsigned u=0; u<2; u++){
                    const unsigned q = u ? qb : qa;
                    #pragma unroll
                    for(unsigned r=0; r<R; r++){
                        const unsigned p = q + r * NB;
                        unsigned i = p;
                        T f = factor;
                        if constexpr (!R2C){
                            const bool upper = 2 * p >= N;
                            i = upper ? 2 * (N - 1 - p) + 1 : 2 * p;
                            if (KIND == real_sin && upper) f = -f;
                        }
                        out1[i] = f * v[u][r].y; out2[i] = f * v[u][r].x;
                    }
                }
            }
        }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// real-data transforms along the middle / slow axis, second generation (plain stores): TWO adjacent real lines make ONE
// complex line.  Neighbouring lines are adjacent in memory, so a row of 2 LPB reals IS a row of LPB complex numbers
// c = x1 + i x2: the tile is loaded exactly like a tile of the complex kernel (16-byte asynchronous copies, no
// de-interleaving, no 8-byte copies), transformed with the complex passes at the full length n, and the two lines are
// separated afterwards from the pair (k, n-k):   V1_k = (C_k + conj C_{n-k}) / 2,   V2_k = -i (C_k - conj C_{n-k}) / 2.
//   r2c:  row k of the output holds (V1_k, V2_k), k <= n/2 -- two adjacent complex numbers, one 32-byte store
//   DCT-II: rows enter in Makhoul's order (a row permutation of the load); y_k = 2 Re(w_k V_k), y_{n-k} = -2 Im(w_k V_k)
//   backward: C_k = V1_k + i V2_k and C_{n-k} = conj V1_k + i conj V2_k are built from the rows k and n-k, the complex passes
//   run on swapped data, and the rows of the result are stored as pairs of reals (DCT-III: permuted, times two)
// Same algebra as fft_strided_real_kernel, on n x LPB complex numbers instead of n/2 x 2 LPB: the same shared memory and
// the same instruction mix per byte as the complex kernel, which runs at the HBM roofline.  Needs an even number of
// adjacent lines and rows aligned to a complex number; other shapes keep the first-generation kernel.
// ---------------------------------------------------------------------------------------------------------
template<typename T, typename RL, int TPL, int LPB, int MINB, int KIND, bool BWD>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_real2_kernel(fft_args a0){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    const fft_args a = batch_entry(a0, shift);
    constexpr unsigned N = RL::N;                  // the real length
    constexpr bool R2C = (KIND == real_r2c);
    constexpr int P = RL::passes;
    cplx<T> *sm = reinterpret_cast<cplx<T>*>(smem_raw);
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle0);
    const cplx<T> *tx = reinterpret_cast<const cplx<T>*>(a.twiddle2);     // W_{4n}^k
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    const scatter_ctx sc{nullptr, 0, 0, 0};
    const unsigned npairs = static_cast<unsigned>((a.nlines + 1) / 2);    // complex lines
    const unsigned half_a = static_cast<unsigned>(a.count_a / 2);         // complex lines per row of the box
    const unsigned ntiles = (npairs + LPB - 1) / LPB;
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        const unsigned pair = tile * LPB + t;
        const bool valid = pair < npairs;
        // offsets of the first line of the pair, in reals
        const unsigned pb = pair / half_a, pa = pair - pb * half_a;
        const long long ioff = valid ? (2LL * pa * a.ig.stride_a + static_cast<long long>(pb) * a.ig.stride_b) : 0;
        const long long ooff = valid ? (2LL * pa * a.og.stride_a + static_cast<long long>(pb) * a.og.stride_b) : 0;

        if constexpr (!BWD){
            // ---- load: rows of two adjacent reals, in Makhoul's order for the cosine / sine transforms -----------------------------
            if (valid){
                const T *src = reinterpret_cast<const T*>(a.in) + ioff;
                #pragma unroll 4
                for(unsigned i = j; i < N; i += TPL){
                    const unsigned p = R2C ? i : ((i & 1) ? N - 1 - (i >> 1) : (i >> 1));
                    async_copy<sizeof(cplx<T>)>(sm + p * LPB + t, src + static_cast<long long>(i) * a.ig.stride);
                }
            }
            async_wait_all();
            __syncthreads();
            // ---- complex passes, the spectrum stays in the tile (digit-reversed positions) -----------------------------------------
            if constexpr (KIND == real_sin){
                // DST-II: the odd samples fill the upper half of the permuted sequence and carry a minus sign
                for(unsigned i = N / 2 + j; i < N; i += TPL){ cplx<T> x = sm[i * LPB + t]; sm[i * LPB + t] = mk<T>(-x.x, -x.y); }
                __syncthreads();
            }
            if constexpr (P > 1){ conv_forward_pass<T, RL, 0, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            if constexpr (P > 2){ conv_forward_pass<T, RL, 1, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            if constexpr (P > 3){ conv_forward_pass<T, RL, 2, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            {
                constexpr unsigned S = P - 1, R = RL::radix(S), NB = N / R;
                #pragma unroll 1
                for(unsigned u=0; u<NB/TPL; u++){
                    cplx<T> *cell = sm + (j + u * TPL) * R * LPB + t;
                    cplx<T> v[R];
                    #pragma unroll
                    for(unsigned r=0; r<R; r++) v[r] = cell[r * LPB];
                    butterfly<T, R>::run(v);
                    #pragma unroll
                    for(unsigned r=0; r<R; r++) cell[r * LPB] = v[r];
                }
            }
            __syncthreads();
            // ---- the two lines apart: pair (k, n-k) ----------------------------------------------------------------------------------
            if (valid){
                const T half = static_cast<T>(0.5);
                T *rout = reinterpret_cast<T*>(a.out) + (R2C ? 2 * ooff : ooff);
                for(unsigned k = j; k <= N / 2; k += TPL){
                    const cplx<T> ck = sm[dif_position_of<RL>(k) * LPB + t], cm = sm[dif_position_of<RL>((N - k) % N) * LPB + t];
                    cplx<T> v1 = mk<T>((ck.x + cm.x) * half, (ck.y - cm.y) * half);
                    cplx<T> v2 = mk<T>((ck.y + cm.y) * half, (cm.x - ck.x) * half);
                    if constexpr (R2C){
                        if (do_scale){ v1.x *= scale; v1.y *= scale; v2.x *= scale; v2.y *= scale; }
                        cplx<T> *dst = reinterpret_cast<cplx<T>*>(rout) + static_cast<long long>(k) * a.og.stride;     // complex row k: (line 1, line 2)
                        dst[0] = v1; dst[1] = v2;
                    }else{
                        const cplx<T> w = ldg_c<T>(tx + k);
                        const cplx<T> z1 = cmul(w, v1), z2 = cmul(w, v2);
                        const T two = do_scale ? T(2) * scale : T(2);
                        auto put = [&](unsigned p, T y1, T y2){
                            const unsigned row = (KIND == real_sin) ? N - 1 - p : p;
                            *reinterpret_cast<cplx<T>*>(rout + static_cast<long long>(row) * a.og.stride) = mk<T>(y1, y2);
                        };
                        put(k, two * z1.x, two * z2.x);
                        if (k > 0 && 2 * k != N) put(N - k, -two * z1.y, -two * z2.y);
                    }
                }
            }
        }else{
            // ---- backward: C_k and C_{n-k} from the rows k and n-k, written swapped for the forward engine -----------------------------------
            // cosine / sine: the rows arrive asynchronously (the sine transform reads the reversed input: row i lands at n-1-i) and
            // every pair is rewritten in place by the thread that read it
            if constexpr (!R2C){
                if (valid){
                    const T *src = reinterpret_cast<const T*>(a.in) + ioff;
                    #pragma unroll 4
                    for(unsigned i = j; i < N; i += TPL){
                        const unsigned p = (KIND == real_sin) ? N - 1 - i : i;
                        async_copy<sizeof(cplx<T>)>(sm + p * LPB + t, src + static_cast<long long>(i) * a.ig.stride);
                    }
                }
                async_wait_all();
                __syncthreads();
            }
            if (valid){
                const T *rin = reinterpret_cast<const T*>(a.in) + (R2C ? 2 * ioff : ioff);
                for(unsigned k = j; k <= N / 2; k += TPL){
                    cplx<T> v1, v2;
                    if constexpr (R2C){
                        const cplx<T> *src = reinterpret_cast<const cplx<T>*>(rin) + static_cast<long long>(k) * a.ig.stride;
                        v1 = src[0]; v2 = src[1];
                        if (k == 0 || 2 * k == N){ v1.y = 0; v2.y = 0; }       // c2r ignores the imaginary part of the self-conjugate entries
                    }else{
                        // V_k = conj(w_k) (y_k - i y_{n-k}), y_n := 0
                        const cplx<T> yk = sm[k * LPB + t];
                        const cplx<T> ym = (k == 0) ? mk<T>(0, 0) : sm[(N - k) * LPB + t];
                        const cplx<T> w = ldg_c<T>(tx + k);
                        v1 = cmul(mk<T>(yk.x, -ym.x), mk<T>(w.x, -w.y));
                        v2 = cmul(mk<T>(yk.y, -ym.y), mk<T>(w.x, -w.y));
                    }
                    // C_k = V1_k + i V2_k,  C_{n-k} = conj(V1_k) + i conj(V2_k); stored swapped
                    const cplx<T> ck = mk<T>(v1.x - v2.y, v1.y + v2.x);
                    const cplx<T> cm = mk<T>(v1.x + v2.y, v2.x - v1.y);
                    sm[k * LPB + t] = cswap(ck);
                    if (k > 0 && 2 * k != N) sm[(N - k) * LPB + t] = cswap(cm);
                }
            }
            __syncthreads();
            if constexpr (P > 1){ conv_forward_pass<T, RL, 0, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            if constexpr (P > 2){ conv_forward_pass<T, RL, 1, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            if constexpr (P > 3){ conv_forward_pass<T, RL, 2, TPL, LPB>(sm, t, j, tw); __syncthreads(); }
            {
                // last pass: natural index of leg r is k0 + r N/R; the two reals of a row are stored together
                constexpr unsigned S = P - 1, R = RL::radix(S), NB = N / R;
                T *rout = reinterpret_cast<T*>(a.out) + ooff;
                const T factor = (R2C ? T(1) : T(2)) * (do_scale ? scale : T(1));
                #pragma unroll 1
                for(unsigned u=0; u<NB/TPL; u++){
                    const unsigned p0 = (j + u * TPL) * R;
                    cplx<T> *cell = sm + p0 * LPB + t;
                    cplx<T> v[R];
                    #pragma unroll
                    for(unsigned r=0; r<R; r++) v[r] = cell[r * LPB];
                    butterfly<T, R>::run(v);
                    if (valid){
                        const unsigned k0 = dif_output_index<RL>(p0);
                        #pragma unroll
                        for(unsigned r=0; r<R; r++){
                            const unsigned e = k0 + r * (N / R);          // position in the sequence v (swapped: .y is line 1, .x is line 2)
                            unsigned row; T f = factor;
                            if constexpr (R2C) row = e;
                            else{
                                row = (e < (N + 1) / 2) ? 2 * e : 2 * (N - 1 - e) + 1;
                                if (KIND == real_sin && (row & 1)) f = -f;
                            }
                            *reinterpret_cast<cplx<T>*>(rout + static_cast<long long>(row) * a.og.stride) = mk<T>(v[r].y * f, v[r].x * f);
                        }
                    }
                }
            }
        }
        if (tile + gridDim.x < ntiles) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// generic kernel: any N that fits shared memory twice.  Stockham passes by prime factor; each thread produces
// one output element per pass:  out[o] = sum_r in[q + r*N/p] * W_N^(r*base mod N).
// Load/store "modes" make it the engine for c2c, r2c, c2r and the r2r (DCT/DST) transforms:
// ---------------------------------------------------------------------------------------------------------
enum generic_mode : int {
    mode_c2c = 0,
    mode_r2c = 1,     // real input line of length N, complex output of length N/2+1
    mode_c2r = 2,     // complex input N/2+1 (Hermitian half), real output N (unnormalised)
    mode_dct2 = 3,    // REDFT10:   y_k = 2 sum x_i cos(pi (2i+1) k / 2n)                 (heffte cos forward)
    mode_dct3 = 4,    // 2*REDFT01: y_i = 2 x_0 + 4 sum_{k>=1} x_k cos(pi k (2i+1) / 2n)  (heffte cos backward)
    mode_dst2 = 5,    // RODFT10:   y_k = 2 sum x_i sin(pi (2i+1)(k+1) / 2n)              (heffte sin forward)
    mode_dst3 = 6,    // 2*RODFT01                                                        (heffte sin backward)
    mode_dct1 = 7     // cos1: REDFT00 (forward), 2*REDFT00 (backward) -- selected by args.backward
};

struct generic_args {
    const void *in;
    void *out;
    const void *twiddle;     // W_M^k for k < M where M = engine length (N, or 4n / 4(n-1) for some r2r modes)
    line_geom ig, og;
    long long nlines;
    int count_a;
    int backward;
    double scale;
    int n;                   // user-visible line length
    int m;                   // engine (complex FFT) length
    int mode;
    int lpb;                 // lines per block
    int lines_fast;          // 1: thread index runs across lines first (neighbour lines adjacent in memory)
    int nfactors;
    int factors[24];
    const scatter_map *smap;   // device pointer or null: fused reshape on the store side (read from global memory, L1-resident)
    long long in_step, out_step, scatter_step, local_shift, local_step;   // batched transforms, see fft_args
};

template<typename T>
__device__ __forceinline__ cplx<T> generic_load(generic_args const &a, const void *base, long long off, int i){
    // returns element i (0 <= i < m) of the complex engine input built from the user line
    const int n = a.n;
    switch(a.mode){
        case mode_c2c: {
            cplx<T> x = reinterpret_cast<const cplx<T>*>(base)[off + (long long)i * a.ig.stride];
            return a.backward ? cswap(x) : x;
        }
        case mode_r2c: {
            T x = reinterpret_cast<const T*>(base)[off + (long long)i * a.ig.stride];
            return mk<T>(x, 0);
        }
        case mode_c2r: {
            // Hermitian extension, then backward = swap trick
            int h = n / 2;
            cplx<T> x;
            if (i <= h) x = reinterpret_cast<const cplx<T>*>(base)[off + (long long)i * a.ig.stride];
            else { x = reinterpret_cast<const cplx<T>*>(base)[off + (long long)(n - i) * a.ig.stride]; x.y = -x.y; }
            if (i == 0 || (2 * i == n)) x.y = 0;   // c2r ignores the imaginary part of the self-conjugate entries
            return cswap(x);
        }
        case mode_dct2: {
            // Makhoul reordering: v_i = x_{2i} (i < ceil(n/2)), v_{n-1-i} = x_{2i+1}
            int src = (2 * i < n) ? 2 * i : 2 * (n - 1 - i) + 1;
            T x = reinterpret_cast<const T*>(base)[off + (long long)src * a.ig.stride];
            return mk<T>(x, 0);
        }
        case mode_dst2: {
            // DST-II of x equals the reversed DCT-II of (-1)^i x_i
            int src = (2 * i < n) ? 2 * i : 2 * (n - 1 - i) + 1;
            T x = reinterpret_cast<const T*>(base)[off + (long long)src * a.ig.stride];
            return mk<T>((src & 1) ? -x : x, 0);
        }
        default: {
            // dct3 / dst3 / dct1 build their spectrum in a pre-pass (see generic kernel), never through here
            return mk<T>(0, 0);
        }
    }
}

// element i (0 <= i < m) of the complex engine input of the line at element offset `off`, every mode; w4n: W_{4n}^k, k < n
template<typename T>
__device__ __forceinline__ cplx<T> generic_input(generic_args const &a, long long off, int i, const cplx<T> *w4n){
    const int n = a.n, m = a.m;
    if (a.mode == mode_dct3 || a.mode == mode_dst3){
        // inverse of the Makhoul post-twiddle: V_k = e^{+i pi k / 2n} (y_k - i y_{n-k}), y_n := 0;
        // for the sine variant y is replaced by the reversed input (y_k -> x_{n-1-k}).
        const T *src = reinterpret_cast<const T*>(a.in);
        T yk, ynk;
        if (a.mode == mode_dct3){
            yk  = src[off + (long long)i * a.ig.stride];
            ynk = (i == 0) ? T(0) : src[off + (long long)(n - i) * a.ig.stride];
        }else{
            yk  = src[off + (long long)(n - 1 - i) * a.ig.stride];
            ynk = (i == 0) ? T(0) : src[off + (long long)(i - 1) * a.ig.stride];
        }
        cplx<T> w = ldg_c<T>(w4n + i);            // e^{+i pi k/2n} = conj(W_{4n}^k)
        cplx<T> z = mk<T>(yk, -ynk);
        return cswap(cmul(z, mk<T>(w.x, -w.y)));  // swapped: the engine runs forward
    }
    if (a.mode == mode_dct1){
        // even extension of length m = 2(n-1): s_i = x_i (i < n), s_{m-i} = x_i
        int src = (i < n) ? i : m - i;
        return mk<T>(reinterpret_cast<const T*>(a.in)[off + (long long)src * a.ig.stride], 0);
    }
    return generic_load<T>(a, a.in, off, i);
}

// output element i of a line from the engine result res(k) (natural order), written to `where`
template<typename T, typename R>
__device__ __forceinline__ void generic_output(generic_args const &a, R const &res, int i, const cplx<T> *w4n, void *where, T scale){
    const int n = a.n;
    switch(a.mode){
        case mode_c2c: {
            cplx<T> x = res(i);
            if (a.backward) x = cswap(x);
            x.x *= scale; x.y *= scale;
            *static_cast<cplx<T>*>(where) = x;
        } break;
        case mode_r2c: {
            cplx<T> x = res(i);
            x.x *= scale; x.y *= scale;
            *static_cast<cplx<T>*>(where) = x;
        } break;
        case mode_c2r: {
            // engine ran forward on swapped input: result = swap(ifft); real part sits in .y
            *static_cast<T*>(where) = res(i).y * scale;
        } break;
        case mode_dct2: {
            // y_k = 2 Re( e^{-i pi k / 2n} V_k ),  e^{-i pi k/2n} = W_{4n}^k
            cplx<T> z = cmul(res(i), ldg_c<T>(w4n + i));
            *static_cast<T*>(where) = T(2) * z.x * scale;
        } break;
        case mode_dst2: {
            cplx<T> z = cmul(res(n - 1 - i), ldg_c<T>(w4n + (n - 1 - i)));
            *static_cast<T*>(where) = T(2) * z.x * scale;
        } break;
        case mode_dct3: case mode_dst3: {
            // v = ifft(V) * n is in res (swapped: real part in .y); x_{2i} = v_i, x_{2i+1} = v_{n-1-i}
            int srcpos = (i & 1) ? (n - 1 - (i >> 1)) : (i >> 1);
            T v = res(srcpos).y * T(2);
            if (a.mode == mode_dst3 && (i & 1)) v = -v;
            *static_cast<T*>(where) = v * scale;
        } break;
        case mode_dct1: {
            T v = res(i).x;
            if (a.backward) v *= T(2);
            *static_cast<T*>(where) = v * scale;
        } break;
    }
}

template<typename T>
__global__ void fft_generic_kernel(generic_args a){
    B200_DYN_SMEM(smem_raw);
    batch_shift shift;
    {
        const long long e = blockIdx.y;
        a.in = static_cast<const char*>(a.in) + e * a.in_step;
        if (a.out != nullptr) a.out = static_cast<char*>(a.out) + e * a.out_step;
        shift.all = e * a.scatter_step;
        shift.local = a.local_shift + e * a.local_step;
    }
    const int m = a.m, n = a.n;
    const int lpb = a.lpb;
    cplx<T> *buf0 = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T> *buf1 = buf0 + (size_t)lpb * m;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const long long line0 = (long long)blockIdx.x * lpb;
    const int total = lpb * m;

    // ---- load ------------------------------------------------------------------------------------------
    for(int idx = threadIdx.x; idx < total; idx += blockDim.x){
        int t, i;
        if (a.lines_fast){ t = idx % lpb; i = idx / lpb; } else { i = idx % m; t = idx / m; }
        long long line = line0 + t;
        cplx<T> x = mk<T>(0, 0);
        if (line < a.nlines) x = generic_input<T>(a, line_offset(a.ig, a.count_a, line), i, tw + m);
        buf0[t * m + i] = x;
    }
    __syncthreads();

    // ---- Stockham passes by factor -----------------------------------------------------------------------
    cplx<T> *src = buf0, *dst = buf1;
    int ns = 1;
    for(int f=0; f<a.nfactors; f++){
        const int p = a.factors[f];
        const int np = m / p;            // butterflies per line
        const int step = m / (ns * p);
        for(int idx = threadIdx.x; idx < total; idx += blockDim.x){
            int t = idx / m, o = idx - t * m;
            int k = o % ns;
            int rp = (o / ns) % p;
            int blk = o / (ns * p);
            int q = blk * ns + k;
            long long base = ((long long)k * step + (long long)rp * np) % m;
            const cplx<T> *s = src + t * m + q;
            cplx<T> acc = s[0];
            int e = 0;
            for(int r=1; r<p; r++){
                e += (int)base; if (e >= m) e -= m;
                acc = cadd(acc, cmul(s[r * np], ldg_c<T>(tw + e)));
            }
            dst[t * m + o] = acc;
        }
        __syncthreads();
        cplx<T> *tmp = src; src = dst; dst = tmp;
        ns *= p;
    }

    // ---- store -------------------------------------------------------------------------------------------
    const T scale = static_cast<T>(a.scale);
    int nout;
    switch(a.mode){
        case mode_r2c: nout = n / 2 + 1; break;
        default: nout = n;
    }
    const int total_out = lpb * nout;
    for(int idx = threadIdx.x; idx < total_out; idx += blockDim.x){
        int t, i;
        if (a.lines_fast){ t = idx % lpb; i = idx / lpb; } else { i = idx % nout; t = idx / nout; }
        long long line = line0 + t;
        if (line >= a.nlines) continue;
        const cplx<T> *res = src + t * m;
        // destination of element i of this line: plain strided output, or the fused reshape
        void *where;
        {
            const bool complex_out = (a.mode == mode_c2c || a.mode == mode_r2c);
            if (a.smap != nullptr){
                const int lb = static_cast<int>(line / a.count_a), la = static_cast<int>(line - static_cast<long long>(lb) * a.count_a);
                const int row = scatter_row(a.smap, la, lb);
                where = complex_out ? static_cast<void*>(scatter_address_shifted<cplx<T>>(a.smap, row, i, la, lb, shift))
                                    : static_cast<void*>(scatter_address_shifted<T>(a.smap, row, i, la, lb, shift));
            }else{
                const long long pos = line_offset(a.og, a.count_a, line) + (long long)i * a.og.stride;
                where = complex_out ? static_cast<void*>(reinterpret_cast<cplx<T>*>(a.out) + pos) : static_cast<void*>(reinterpret_cast<T*>(a.out) + pos);
            }
        }
        generic_output<T>(a, [res](int k){ return res[k]; }, i, tw + m, where, scale);
    }
}

} // namespace b200
