// Device code of the packers, scaling and real<->complex conversion.
// Replaces the reference kernels in src/heffte_backend_cuda.cu: direct_packer (:61-85), transpose_unpacker (:90-133),
// simple_scal (:138-145) and real_complex_convert (:44-56).  Differences: 64-bit indexing throughout (the reference
// uses int offsets), element-wide (8/16-byte) accesses instead of scalar 4/8-byte ones, one launch per box instead of a
// 1024-thread block per line, and a shared-memory tile for permuting copies so both sides stay coalesced.
#pragma once

#include "cuda_compat.h"
#include "fft_device.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// scatter copy: the first reshape of a transform (user box -> the boxes of the first FFT stage) written straight into
// the destination boxes, local or on a peer GPU over NVLink.  One pass: no send buffer, no receive buffer, no unpack.
// Coordinates of the map: k = fast index of my box, a = mid, b = slow.
// ---------------------------------------------------------------------------------------------------------
struct scatter_copy_args {
    const void *src;
    const scatter_map *smap;     // device pointer
    int nfast, nmid, nslow;
    long long line, plane;       // strides of my box
    int tf;                      // threads along the fast axis (power of two <= 256)
    long long in_step, scatter_step, local_shift, local_step;   // batched transforms (bytes), see fft_args
};

template<typename V>
__global__ void __launch_bounds__(256) scatter_copy_kernel(scatter_copy_args a){
    B200_DYN_SMEM(map_raw);
    scatter_map *smap = reinterpret_cast<scatter_map*>(map_raw);
    const long long entry = blockIdx.y;
    scatter_stage(smap, a.smap, batch_shift{entry * a.scatter_step, a.local_shift + entry * a.local_step});
    __syncthreads();
    const V *src = reinterpret_cast<const V*>(static_cast<const char*>(a.src) + entry * a.in_step);
    const int tx = threadIdx.x & (a.tf - 1), ty = threadIdx.x / a.tf;
    const int rows = blockDim.x / a.tf;                         // lines handled side by side by one CTA
    const long long nlines = static_cast<long long>(a.nmid) * a.nslow;
    // planes are visited round-robin over the nb destination ranges of the slow axis: the CTAs that run side by side then write
    // to every destination GPU at once instead of all of them to the same one (the ranks of a plan walk their boxes in step, so
    // a plane-by-plane sweep would aim every sender at the same receiver at the same time)
    const int ranges = (smap->nb > 1 && a.nslow % smap->nb == 0) ? smap->nb : 1;
    const int per_range = a.nslow / ranges;
    for(long long line = static_cast<long long>(blockIdx.x) * rows + ty; line < nlines; line += static_cast<long long>(gridDim.x) * rows){
        const int visit = static_cast<int>(line / a.nmid), m = static_cast<int>(line - static_cast<long long>(visit) * a.nmid);
        const int s = (visit % ranges) * per_range + visit / ranges;
        const int row = scatter_row(smap, m, s);
        const V *from = src + s * a.plane + m * a.line;
        int f = tx;
        // four independent loads in flight per thread
        for(; f + 3 * a.tf < a.nfast; f += 4 * a.tf){
            V v0 = from[f], v1 = from[f + a.tf], v2 = from[f + 2 * a.tf], v3 = from[f + 3 * a.tf];
            *scatter_address<V>(smap, row, f, m, s) = v0;
            *scatter_address<V>(smap, row, f + a.tf, m, s) = v1;
            *scatter_address<V>(smap, row, f + 2 * a.tf, m, s) = v2;
            *scatter_address<V>(smap, row, f + 3 * a.tf, m, s) = v3;
        }
        for(; f < a.nfast; f += a.tf) *scatter_address<V>(smap, row, f, m, s) = from[f];
    }
}

// ---------------------------------------------------------------------------------------------------------
// barrier between the GPUs of a plan, stream ordered: rank r publishes `epoch` into slot r of every peer's flag array
// (peer memory, NVLink) and waits until every peer has published it into this rank's array.  All data written by the
// kernels enqueued before the barrier is visible to the peers once they pass it.
// ---------------------------------------------------------------------------------------------------------
constexpr int barrier_max_ranks = 64;
struct peer_barrier_args {
    unsigned long long *remote[barrier_max_ranks];   // remote[p]: address of slot `me` in the flag array of rank p
    unsigned long long *local;                       // my flag array (slot p is written by rank p)
    unsigned long long epoch;
    unsigned long long timeout_ns;                   // 0: wait for ever, like the MPI calls of the reference
    unsigned long long *timeout_word;                // mapped host memory: receives (epoch << 8 | peer + 1) when a wait expires
    int nranks, me;
};

#ifndef B200_HOST_EMULATION
// Spin until *flag >= target (system-scope acquire).  A late peer is NOT an error: by default the wait has no limit, exactly like
// a rank blocked in MPI_Alltoallv.  With a limit (HEFFTE_B200_BARRIER_TIMEOUT_S) an expired wait records itself in mapped host
// memory and gives up WITHOUT trapping -- a trap would poison the CUDA context of the whole job; the next transform call on this
// rank returns B200_ERR_PEER instead.
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long target, unsigned long long timeout_ns,
                                           unsigned long long *timeout_word, unsigned long long tag){
    unsigned long long seen = 0, start = 0, now;
    if (timeout_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(start));
    for(;;){
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
        if (seen >= target) return true;
        if (timeout_ns){
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - start > timeout_ns){
                if (timeout_word) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(timeout_word), "l"(tag) : "memory");
                return false;
            }
        }
    }
}

__global__ void __launch_bounds__(barrier_max_ranks) peer_barrier_kernel(peer_barrier_args a){
    const int p = threadIdx.x;
    if (p >= a.nranks || p == a.me) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(a.remote[p]), "l"(a.epoch) : "memory");
    spin_until(a.local + p, a.epoch, a.timeout_ns, a.timeout_word, (a.epoch << 8) | static_cast<unsigned long long>(p + 1));
}
#endif

// ---------------------------------------------------------------------------------------------------------
// sub-boxes of one box copied to the same positions of another array with the same layout: after the last fused reshape of a
// plan the pieces received from the other GPUs move from the plan's arena into the caller's array.  One launch for all
// pieces (blockIdx.y) and all batch entries (blockIdx.z); rows are moved in 16-byte words when the geometry allows.
// ---------------------------------------------------------------------------------------------------------
constexpr int multi_copy_max = 16;
struct multi_copy_args {
    const void *src;
    void *dst;
    long long line, plane;                 // strides of the box, in elements
    long long offset[multi_copy_max];      // first element of every piece
    int nfast[multi_copy_max], nmid[multi_copy_max], nslow[multi_copy_max];
    int npieces;
    long long src_step, dst_step;          // bytes between batch entries
};

template<typename V, int RATIO>      // V: access type, RATIO: elements per access (geometry in elements of sizeof(V) / RATIO bytes)
__global__ void __launch_bounds__(256) multi_copy_kernel(multi_copy_args a){
    const int piece = blockIdx.y;
    const long long entry = blockIdx.z;
    const V *src = reinterpret_cast<const V*>(static_cast<const char*>(a.src) + entry * a.src_step);
    V *dst = reinterpret_cast<V*>(static_cast<char*>(a.dst) + entry * a.dst_step);
    const long long base = a.offset[piece] / RATIO, line = a.line / RATIO, plane = a.plane / RATIO;
    const int nf = a.nfast[piece] / RATIO, nm = a.nmid[piece];
    const long long rows = static_cast<long long>(nm) * a.nslow[piece];
    // a CTA moves whole rows; short rows share a CTA pass
    const int tf = (nf >= 256) ? 256 : ((nf >= 128) ? 128 : ((nf >= 64) ? 64 : 32));
    const int tx = threadIdx.x % tf, ty = threadIdx.x / tf, side = 256 / tf;
    for(long long row = static_cast<long long>(blockIdx.x) * side + ty; row < rows; row += static_cast<long long>(gridDim.x) * side){
        const long long s = row / nm, m = row - s * nm;
        const long long at = base + s * plane + m * line;
        for(int f = tx; f < nf; f += tf) dst[at + f] = src[at + f];
    }
}

struct copy3d_args {
    const void *src;
    void *dst;
    long long nfast, nmid, nslow;
    long long src_line, src_plane;   // strides of the source (elements)
    long long dst_line, dst_plane;   // strides of the destination
};

// dst[s*dst_plane + m*dst_line + f] = src[s*src_plane + m*src_line + f]
template<typename V>
__global__ void __launch_bounds__(256) copy3d_kernel(copy3d_args a){
    const V *src = reinterpret_cast<const V*>(a.src);
    V *dst = reinterpret_cast<V*>(a.dst);
    const long long total = a.nfast * a.nmid * a.nslow;
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step){
        long long line = idx / a.nfast;
        long long f = idx - line * a.nfast;
        long long s = line / a.nmid;
        long long m = line - s * a.nmid;
        dst[s * a.dst_plane + m * a.dst_line + f] = src[s * a.src_plane + m * a.src_line + f];
    }
}

struct permute_args {
    const void *src;
    void *dst;
    long long size[3];        // extents in DESTINATION order (fast, mid, slow)
    long long dst_stride[3];  // 1, line, plane
    long long src_stride[3];  // stride in the source of destination index 0, 1, 2
    // tiled variant: t = destination dimension that is unit-stride in the source, o = the remaining one
    long long nt, no;
    long long src_f, src_t, src_o;   // source strides of the destination-fast, tile and other dimensions
    long long dst_t, dst_o;          // destination strides of the tile and other dimensions
};

// Permuting copy through a 32x32 shared-memory tile: reads run along the source-fast axis, writes along the
// destination-fast axis, so both sides are coalesced.  grid = (tiles_f, tiles_t, size[other]).
template<typename V>
__global__ void __launch_bounds__(256) permute_tile_kernel(permute_args a){
    B200_DYN_SMEM(tile_raw);
    V (*tile)[33] = reinterpret_cast<V (*)[33]>(tile_raw);
    const V *src = reinterpret_cast<const V*>(a.src);
    V *dst = reinterpret_cast<V*>(a.dst);
    const long long f0 = (long long)blockIdx.x * 32, t0 = (long long)blockIdx.y * 32;
    const long long nf = a.size[0], nt = a.nt;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for(long long o = blockIdx.z; o < a.no; o += gridDim.z){
        // read: tx runs along the source-fast axis
        #pragma unroll
        for(int r=0; r<4; r++){
            long long f = f0 + ty + 8 * r, t = t0 + tx;
            if (f < nf && t < nt)
                tile[ty + 8 * r][tx] = src[f * a.src_f + t * a.src_t + o * a.src_o];
        }
        __syncthreads();
        // write: tx runs along the destination-fast axis
        #pragma unroll
        for(int r=0; r<4; r++){
            long long f = f0 + tx, t = t0 + ty + 8 * r;
            if (f < nf && t < nt)
                dst[f + t * a.dst_t + o * a.dst_o] = tile[tx][ty + 8 * r];
        }
        __syncthreads();
    }
}

// fallback for shapes where no destination dimension is unit-stride in the source
template<typename V>
__global__ void __launch_bounds__(256) permute_simple_kernel(permute_args a){
    const V *src = reinterpret_cast<const V*>(a.src);
    V *dst = reinterpret_cast<V*>(a.dst);
    const long long total = a.size[0] * a.size[1] * a.size[2];
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step){
        long long line = idx / a.size[0];
        long long f = idx - line * a.size[0];
        long long s = line / a.size[1];
        long long m = line - s * a.size[1];
        dst[f + m * a.dst_stride[1] + s * a.dst_stride[2]] = src[f * a.src_stride[0] + m * a.src_stride[1] + s * a.src_stride[2]];
    }
}

struct scale_args { void *data; long long count; double factor; };

template<typename T>
__global__ void __launch_bounds__(256) scale_kernel(scale_args a){
    T *data = reinterpret_cast<T*>(a.data);
    const T factor = static_cast<T>(a.factor);
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += step) data[i] *= factor;
}

struct pointwise_args { void *data; const void *multiplier; long long count; double factor; };

// data[i] = data[i] * factor * M[i]  (M null: data[i] itself): the pointwise product of a spectral operator, unfused form
template<typename C, typename T>
__global__ void __launch_bounds__(256) pointwise_kernel(pointwise_args a){
    C *data = reinterpret_cast<C*>(a.data);
    const C *mult = reinterpret_cast<const C*>(a.multiplier);
    const T factor = static_cast<T>(a.factor);
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += step){
        C x = data[i];
        x.x *= factor; x.y *= factor;
        const C m = (mult != nullptr) ? mult[i] : x;
        C r; r.x = x.x * m.x - x.y * m.y; r.y = x.x * m.y + x.y * m.x;
        data[i] = r;
    }
}

struct convert_args { const void *src; void *dst; long long count; };

template<typename T, typename C>
__global__ void __launch_bounds__(256) real_to_complex_kernel(convert_args a){
    const T *src = reinterpret_cast<const T*>(a.src);
    C *dst = reinterpret_cast<C*>(a.dst);
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += step){
        C z; z.x = src[i]; z.y = 0;
        dst[i] = z;
    }
}
template<typename T, typename C>
__global__ void __launch_bounds__(256) complex_to_real_kernel(convert_args a){
    const C *src = reinterpret_cast<const C*>(a.src);
    T *dst = reinterpret_cast<T*>(a.dst);
    const long long step = (long long)gridDim.x * blockDim.x;
    for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += step) dst[i] = src[i].x;
}

} // namespace b200
