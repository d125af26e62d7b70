// Developer micro-benchmark (not part of the product): the tile of the 512-point fp64 strided kernel loaded by TMA
// (cp.async.bulk.tensor, one elected thread, mbarrier completion) instead of per-thread cp.async copies -- the staging the
// north_star names.  Same passes (strided_pass of the product), bit-identical output required.  Middle and slow axis of 512^3.
#include "../heffte_b200/csrc/fft_host_plan.h"
#include <cuda.h>
#include <cstdio>
#include <cstdint>
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

__device__ __forceinline__ unsigned smem_u32(const void *p){ return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// PIPE = 1: one tile buffer, the next tile is requested after the passes of this one; PIPE = 2: two buffers, the load of the next
// tile is in flight while this one is transformed
template<typename T, typename RL, int TPL, int LPB, int MINB, bool BWD, int PIPE, int ROWS = 256>
__global__ void __launch_bounds__(TPL * LPB, MINB) fft_strided_tma_kernel(fft_args a, const __grid_constant__ CUtensorMap tmap){
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr unsigned N = RL::N;
    constexpr unsigned TILE_BYTES = N * LPB * sizeof(cplx<T>);
    constexpr int P = RL::passes;
    unsigned long long *bars = reinterpret_cast<unsigned long long*>(smem_raw + static_cast<size_t>(PIPE) * TILE_BYTES);
    const unsigned t = threadIdx.x % LPB, j = threadIdx.x / LPB;
    const cplx<T> *tw = reinterpret_cast<const cplx<T>*>(a.twiddle);
    const T scale = static_cast<T>(a.scale);
    const bool do_scale = a.scale != 1.0;
    const scatter_ctx sc{nullptr, 0, 0, 0};
    if (threadIdx.x == 0){
        for(int s=0; s<PIPE; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const unsigned ntiles = static_cast<unsigned>(a.nlines / LPB);
    auto request = [&](unsigned tile, int slot){
        const unsigned line0 = tile * LPB;
        const unsigned b = line0 / static_cast<unsigned>(a.count_a);
        const unsigned a0 = line0 - b * static_cast<unsigned>(a.count_a);
        const unsigned bar = smem_u32(bars + slot);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(TILE_BYTES) : "memory");
        #pragma unroll
        for(unsigned r0 = 0; r0 < N; r0 += ROWS){
            const unsigned dst = smem_u32(smem_raw + static_cast<size_t>(slot) * TILE_BYTES + static_cast<size_t>(r0) * LPB * sizeof(cplx<T>));
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                         :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(static_cast<int>(2 * a0 * (sizeof(T) == 8 ? 1 : 1))), "r"(static_cast<int>(r0)), "r"(static_cast<int>(b)), "r"(bar) : "memory");
        }
    };
    unsigned phase[PIPE];
    for(int s=0; s<PIPE; s++) phase[s] = 0;
    unsigned k = 0;
    if (threadIdx.x == 0 && blockIdx.x < ntiles) request(blockIdx.x, 0);
    for(unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++){
        const int slot = (PIPE == 1) ? 0 : static_cast<int>(k % PIPE);
        if (PIPE == 2 && threadIdx.x == 0 && tile + gridDim.x < ntiles) request(tile + gridDim.x, (slot + 1) % PIPE);
        {   // wait for the tile
            const unsigned bar = smem_u32(bars + slot);
            unsigned done = 0;
            while(!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(phase[slot]) : "memory");
            phase[slot] ^= 1;
        }
        cplx<T> *sm = reinterpret_cast<cplx<T>*>(smem_raw + static_cast<size_t>(slot) * TILE_BYTES);
        const unsigned line = tile * LPB + t;
        cplx<T> *gout = reinterpret_cast<cplx<T>*>(a.out) + tile_line_offset(a.og, a.count_a, line);
        strided_pass<T, RL, 0, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc);
        if constexpr (P > 1){ __syncthreads(); strided_pass<T, RL, 1, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if constexpr (P > 2){ __syncthreads(); strided_pass<T, RL, 2, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if constexpr (P > 3){ __syncthreads(); strided_pass<T, RL, 3, TPL, LPB, BWD, false>(sm, t, j, true, gout, a.og.stride, tw, scale, do_scale, sc); }
        if (tile + gridDim.x < ntiles){
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // generic accesses of this buffer before the next bulk write
            __syncthreads();
            if (PIPE == 1 && threadIdx.x == 0) request(tile + gridDim.x, 0);
        }
    }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(){
    L l;
    const int n = 512; const long long elems = (long long)n*n*n; const double gb = 2.0 * elems * 16 * 1e-9;
    double2 *x, *y1, *y2; CK(cudaMalloc(&x, elems * 16)); CK(cudaMalloc(&y1, elems * 16)); CK(cudaMalloc(&y2, elems * 16));
    {
        std::vector<double2> h(elems);
        for(long long i=0; i<elems; i++){ h[i].x = (double)((i * 2654435761u) % 1000) * 1e-3; h[i].y = (double)((i * 40503u) % 977) * 1e-3; }
        CK(cudaMemcpy(x, h.data(), elems*16, cudaMemcpyHostToDevice));
    }
    encode_fn encode = nullptr;
    { void *f = nullptr; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q)); if (q != cudaDriverEntryPointSuccess){ printf("no cuTensorMapEncodeTiled\n"); return 1; } encode = (encode_fn) f; }
    auto report = [&](const char *name, float ms){ printf("%-70s %8.3f ms  %7.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
    fft_args a{}; a.twiddle = twiddles<double>(n); a.twiddle2 = nullptr; a.nlines = elems / n; a.backward = 0; a.scale = 1.0; a.smap = nullptr;
    using R888 = radix_list<8,8,8,1>;
    constexpr int LPB = 8, TPL = 32;
    for(int dim=1; dim<=2; dim++){
        long long count_b, stride_b;
        if (dim == 1){ a.ig = a.og = line_geom{n, 1, (long long)n*n}; a.count_a = n; count_b = n; stride_b = (long long)n*n; }
        else { a.ig = a.og = line_geom{(long long)n*n, 1, 0}; a.count_a = n*n; count_b = 1; stride_b = elems; }
        CUtensorMap tmap, tmap128, tmap64rows, tmap_nopromo;
        cuuint64_t gdim[3] = {(cuuint64_t) 2 * a.count_a, (cuuint64_t) n, (cuuint64_t) count_b};
        cuuint64_t gstride[2] = {(cuuint64_t) a.ig.stride * 16, (cuuint64_t) stride_b * 16};
        cuuint32_t box[3] = {2 * LPB, 256, 1}, box64[3] = {2 * LPB, 64, 1}, estr[3] = {1, 1, 1};
        CUresult rc = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc == CUDA_SUCCESS) rc = encode(&tmap128, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc == CUDA_SUCCESS) rc = encode(&tmap_nopromo, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc == CUDA_SUCCESS) rc = encode(&tmap64rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, x, gdim, gstride, box64, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS){ printf("cuTensorMapEncodeTiled failed: %d\n", (int) rc); return 1; }
        printf("-- fp64 512 strided, dim %d (out of place: x -> y)\n", dim);
        a.in = x; a.out = y1;
        report("cp.async tile (product kernel) <8,8,8> TPL32 LPB8 minb3", timeit([&]{ launch_strided<double, R888, TPL, LPB, 3, false>(a, l); }));
        a.out = y2;
        auto run_tma = [&](auto kernel, int pipe, long long blocks){
            size_t smem = (size_t) pipe * n * LPB * 16 + 64;
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
            kernel<<<(unsigned) blocks, TPL * LPB, smem>>>(a, tmap);
        };
        long long const tiles = a.nlines / LPB;
        report("TMA tile, one tile per CTA, minb3", timeit([&]{ run_tma(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1>, 1, tiles); }));
        {
            auto run_with = [&](auto kernel, CUtensorMap const &m){ size_t smem = (size_t) n * LPB * 16 + 64; cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024); kernel<<<(unsigned) tiles, TPL * LPB, smem>>>(a, m); };
            report("TMA tile, L2 promotion 128 B", timeit([&]{ run_with(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1>, tmap128); }));
            report("TMA tile, no L2 promotion", timeit([&]{ run_with(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1>, tmap_nopromo); }));
            report("TMA tile, eight requests of 64 rows", timeit([&]{ run_with(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1, 64>, tmap64rows); }));
            report("TMA tile, minb2 (two CTAs per SM)", timeit([&]{ run_with(fft_strided_tma_kernel<double, R888, TPL, LPB, 2, false, 1>, tmap); }));
        }
        {
            CK(cudaDeviceSynchronize());
            std::vector<double2> h1(1 << 20), h2(1 << 20);
            long long bad = 0;
            for(long long off : {0LL, elems / 2, elems - (1LL << 20)}){
                CK(cudaMemcpy(h1.data(), y1 + off, h1.size() * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), y2 + off, h2.size() * 16, cudaMemcpyDeviceToHost));
                for(size_t i=0; i<h1.size(); i++) if (h1[i].x != h2[i].x || h1[i].y != h2[i].y) bad++;
            }
            printf("   TMA output vs product kernel: %lld differing entries of %d checked\n", bad, 3 << 20);
        }
        report("TMA tile, persistent 148 x 3 CTAs, one buffer", timeit([&]{ run_tma(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1>, 1, 148 * 3); }));
        report("TMA tile, persistent 148 CTAs, two buffers (minb1)", timeit([&]{ run_tma(fft_strided_tma_kernel<double, R888, TPL, LPB, 1, false, 2>, 2, 148); }));
        CK(cudaMemset(y2, 0, elems * 16));
        report("TMA tile, persistent 148 x 3 CTAs, one buffer (again, for the check)", timeit([&]{ run_tma(fft_strided_tma_kernel<double, R888, TPL, LPB, 3, false, 1>, 1, 148 * 3); }, 2));
        {
            CK(cudaDeviceSynchronize());
            std::vector<double2> h1(1 << 20), h2(1 << 20);
            long long bad = 0;
            for(long long off : {0LL, elems / 3, elems - (1LL << 20)}){
                CK(cudaMemcpy(h1.data(), y1 + off, h1.size() * 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), y2 + off, h2.size() * 16, cudaMemcpyDeviceToHost));
                for(size_t i=0; i<h1.size(); i++) if (h1[i].x != h2[i].x || h1[i].y != h2[i].y) bad++;
            }
            printf("   persistent TMA output vs product kernel: %lld differing entries of %d checked\n", bad, 3 << 20);
        }
    }
    return 0;
}
