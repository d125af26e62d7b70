// Transforms of ANY length: the composite engine behind b200_fft1d plans whose lines neither fit the register / shared-memory
// kernels nor the generic shared-memory kernel, and behind large prime lengths (where the generic kernel costs O(N^2) per line).
// The reference reaches arbitrary lengths through cuFFT (include/heffte_backend_cuda.h:356-368) and, in its stock backend,
// through composite and Rader plans (include/stock_fft/heffte_stock_algos.h:43-69).
//
// A chunk of L lines is gathered into a workspace laid out [position][line] -- the layout of the strided kernels -- through the
// load stage of the generic kernel (c2c, r2c, c2r, DCT / DST pre-processing: generic_input), transformed there, and scattered
// back through its store stage (generic_output).  The transform of the workspace is
//   * four-step, for a length m = n1 n2 (position j = j1 n2 + j2): transforms of length n1 over j1, twiddles W_m^(k1 j2), transforms of
//     length n2 over j2; the result for k = k1 + n1 k2 sits at position k1 n2 + k2.  Both steps are ordinary batched plans of the
//     kernels in fft_device.cuh (sub-plans), each ONE launch for the whole chunk;
//   * Bluestein's chirp-z for lengths that do not split (primes): x_j w_j zero-padded to a power of two m >= 2n - 1, the four-step
//     transform, a pointwise product with the transformed chirp, the inverse four-step (the steps undone in reverse order, which
//     takes the permuted order back to the natural one), and the chirp again.
// The backward direction rides on the re/im swap of the generic load / store stages, so the engine always runs forward.
#pragma once

#include "fft_device.cuh"

namespace b200 {

struct composite_args {
    generic_args g;            // the user-side description: pointers, geometry, mode, n, m = engine input length, scale, backward
    void *work;                // [m_fft][L] complex
    long long L;               // lines per chunk (row length of the workspace)
    long long line0, lines;    // the lines of this chunk
    long long m_fft, n1, n2;   // length of the workspace transform and its split (n2 == 1: no split)
    const void *twiddle;       // W_{m_fft}^t, t < m_fft
    const void *chirp;         // Bluestein: w_j = exp(-i pi j^2 / E), j < E  (null: the direct four-step transform)
    const void *bhat;          // Bluestein: transform of the chirp in workspace order, times 1 / m_fft
    const void *w4n;           // r2r modes: W_{4n}^k, k <= n
    int conjugate;             // twiddle kernel: multiply by the conjugates (inverse four-step)
    batch_shift shift;         // fused reshape on the store side: re-basing of the destinations (see fft_args)
};

// user lines -> workspace (times the chirp, zero padded)
template<typename T>
__global__ void __launch_bounds__(256) composite_load_kernel(composite_args c){
    cplx<T> *work = reinterpret_cast<cplx<T>*>(c.work);
    const cplx<T> *chirp = reinterpret_cast<const cplx<T>*>(c.chirp);
    const long long total = c.m_fft * c.L;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += step){
        const long long j = idx / c.L, t = idx - j * c.L;
        cplx<T> x = mk<T>(0, 0);
        if (t < c.lines && j < c.g.m){
            x = generic_input<T>(c.g, line_offset(c.g.ig, c.g.count_a, c.line0 + t), static_cast<int>(j), reinterpret_cast<const cplx<T>*>(c.w4n));
            if (chirp != nullptr) x = cmul(x, chirp[j]);
        }
        work[idx] = x;
    }
}

// four-step twiddles: position (k1 n2 + j2) times W^(k1 j2)
template<typename T>
__global__ void __launch_bounds__(256) composite_twiddle_kernel(composite_args c){
    cplx<T> *work = reinterpret_cast<cplx<T>*>(c.work);
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(c.twiddle);
    const long long total = c.m_fft * c.L;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += step){
        const long long p = idx / c.L;
        const long long k1 = p / c.n2, j2 = p - k1 * c.n2;
        cplx<T> w = tw[(k1 * j2) % c.m_fft];
        if (c.conjugate) w.y = -w.y;
        work[idx] = cmul(work[idx], w);
    }
}

// Bluestein: times the transformed chirp (stored in workspace order)
template<typename T>
__global__ void __launch_bounds__(256) composite_pointwise_kernel(composite_args c){
    cplx<T> *work = reinterpret_cast<cplx<T>*>(c.work);
    const cplx<T> *bhat = reinterpret_cast<const cplx<T>*>(c.bhat);
    const long long total = c.m_fft * c.L;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += step)
        work[idx] = cmul(work[idx], bhat[idx / c.L]);
}

// workspace -> user lines through the store stage of the generic kernel
template<typename T>
__global__ void __launch_bounds__(256) composite_store_kernel(composite_args c){
    const cplx<T> *work = reinterpret_cast<const cplx<T>*>(c.work);
    const cplx<T> *chirp = reinterpret_cast<const cplx<T>*>(c.chirp);
    const generic_args &a = c.g;
    const long long nout = (a.mode == mode_r2c) ? a.n / 2 + 1 : a.n;
    const long long total = nout * c.L;
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    const T scale = static_cast<T>(a.scale);
    for(long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += step){
        const long long i = idx / c.L, t = idx - i * c.L;
        if (t >= c.lines) continue;
        const long long line = c.line0 + t;
        // element k of the transform of this line: direct four-step leaves it at position (k % n1) n2 + k / n1, Bluestein at k
        auto res = [&](int k){
            if (chirp != nullptr) return cmul(work[static_cast<long long>(k) * c.L + t], chirp[k]);
            const long long pos = (c.n2 == 1) ? k : (static_cast<long long>(k) % c.n1) * c.n2 + static_cast<long long>(k) / c.n1;
            return work[pos * c.L + t];
        };
        const bool complex_out = (a.mode == mode_c2c || a.mode == mode_r2c);
        void *where;
        if (a.smap != nullptr){
            const int lb = static_cast<int>(line / a.count_a), la = static_cast<int>(line - static_cast<long long>(lb) * a.count_a);
            const int row = scatter_row(a.smap, la, lb);
            where = complex_out ? static_cast<void*>(scatter_address_shifted<cplx<T>>(a.smap, row, static_cast<int>(i), la, lb, c.shift))
                                : static_cast<void*>(scatter_address_shifted<T>(a.smap, row, static_cast<int>(i), la, lb, c.shift));
        }else{
            const long long pos = line_offset(a.og, a.count_a, line) + i * a.og.stride;
            where = complex_out ? static_cast<void*>(reinterpret_cast<cplx<T>*>(a.out) + pos) : static_cast<void*>(reinterpret_cast<T*>(a.out) + pos);
        }
        generic_output<T>(a, res, static_cast<int>(i), reinterpret_cast<const cplx<T>*>(c.w4n), where, scale);
    }
}

} // namespace b200
