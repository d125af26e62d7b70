// Device code of the packers, scaling and real<->complex conversion.
// Replaces the reference kernels in src/heffte_backend_cuda.cu: direct_packer (:61-85), transpose_unpacker (:90-133),
// simple_scal (:138-145) and real_complex_convert (:44-56).  Differences: 64-bit indexing throughout (the reference
// uses int offsets), element-wide (8/16-byte) accesses instead of scalar 4/8-byte ones, one launch per box instead of a
// 1024-thread block per line, and a shared-memory tile for permuting copies so both sides stay coalesced.
#pragma once

#include "cuda_compat.h"

namespace b200 {

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
