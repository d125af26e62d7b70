// Host-side launch logic of the packers / scaling / conversion kernels, templated on the Launcher policy
// (CUDA stream in the product, CPU emulation in tests/emul).
#pragma once

#include <algorithm>
#include <cstdint>

#include "../../include/heffte_b200_kernels.h"
#include "pack_device.cuh"

namespace b200 {

constexpr int num_sms = 148;   // B200

inline long long stream_blocks(long long total, int threads, int per_sm = 16){
    long long blocks = (total + threads - 1) / threads;
    return std::max<long long>(1, std::min<long long>(blocks, (long long)num_sms * per_sm));
}

template<typename Launcher>
int launch_copy3d(int elem_bytes, copy3d_args const &a, Launcher &L){
    long long const total = a.nfast * a.nmid * a.nslow;
    if (total <= 0) return B200_SUCCESS;
    long long const blocks = stream_blocks(total, 256);
    switch(elem_bytes){
        case 4:  return L.launch(copy3d_kernel<float>,   blocks, 256, 0, a);
        case 8:  return L.launch(copy3d_kernel<double>,  blocks, 256, 0, a);
        case 16: return L.launch(copy3d_kernel<double2>, blocks, 256, 0, a);
        default: return B200_ERR_INVALID;
    }
}

// the pieces must not be empty; elem_bytes in {4, 8, 16}
template<typename Launcher>
int launch_multi_copy(int elem_bytes, multi_copy_args const &a, int batch, Launcher &L){
    if (a.npieces <= 0 or batch <= 0) return B200_SUCCESS;
    // widest access (16 bytes) that divides every row length, offset and stride
    int ratio = 16 / elem_bytes;
    auto fits = [&](int r){
        if ((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(a.dst)) % (static_cast<uintptr_t>(r) * elem_bytes)) return false;
        if (a.line % r or a.plane % r or (a.src_step % (r * elem_bytes)) or (a.dst_step % (r * elem_bytes))) return false;
        for(int i=0; i<a.npieces; i++) if (a.nfast[i] % r or a.offset[i] % r) return false;
        return true;
    };
    while(ratio > 1 and not fits(ratio)) ratio /= 2;
    long long most_rows = 1;
    for(int i=0; i<a.npieces; i++) most_rows = std::max<long long>(most_rows, static_cast<long long>(a.nmid[i]) * a.nslow[i]);
    long long const gx = std::max<long long>(1, std::min<long long>(most_rows, (long long) num_sms * 8 / std::max(1, a.npieces)));
    int const bytes = elem_bytes * ratio;
    switch(bytes){
        case 4:  return L.launch3(multi_copy_kernel<float, 1>, gx, a.npieces, batch, 256, 0, a);
        case 8:  return (ratio == 2) ? L.launch3(multi_copy_kernel<double, 2>, gx, a.npieces, batch, 256, 0, a) : L.launch3(multi_copy_kernel<double, 1>, gx, a.npieces, batch, 256, 0, a);
        case 16: return (ratio == 4) ? L.launch3(multi_copy_kernel<double2, 4>, gx, a.npieces, batch, 256, 0, a)
                                     : ((ratio == 2) ? L.launch3(multi_copy_kernel<double2, 2>, gx, a.npieces, batch, 256, 0, a)
                                                     : L.launch3(multi_copy_kernel<double2, 1>, gx, a.npieces, batch, 256, 0, a));
        default: return B200_ERR_INVALID;
    }
}

template<typename Launcher>
int launch_scatter_copy(int elem_bytes, scatter_copy_args a, Launcher &L){
    long long const nlines = static_cast<long long>(a.nmid) * a.nslow;
    if (nlines <= 0 or a.nfast <= 0) return B200_SUCCESS;
    int tf = 1;
    while(tf < a.nfast and tf < 256) tf *= 2;
    a.tf = tf;
    long long const rows = 256 / tf;
    long long const blocks = std::max<long long>(1, std::min<long long>((nlines + rows - 1) / rows, (long long)num_sms * 16));
    size_t const smem = sizeof(scatter_map);
    switch(elem_bytes){
        case 4:  return L.launch(scatter_copy_kernel<float>,   blocks, 256, smem, a);
        case 8:  return L.launch(scatter_copy_kernel<double>,  blocks, 256, smem, a);
        case 16: return L.launch(scatter_copy_kernel<double2>, blocks, 256, smem, a);
        default: return B200_ERR_INVALID;
    }
}

// general permuting copy: dst[f + m*dl + s*dp] = src[f*ss0 + m*ss1 + s*ss2]
template<typename Launcher>
int launch_permute(int elem_bytes, permute_args a, Launcher &L){
    long long const total = a.size[0] * a.size[1] * a.size[2];
    if (total <= 0) return B200_SUCCESS;
    int tile_dim = -1;
    if (a.src_stride[0] != 1){
        if (a.src_stride[1] == 1) tile_dim = 1;
        else if (a.src_stride[2] == 1) tile_dim = 2;
    }
    if (tile_dim < 0){
        long long const blocks = stream_blocks(total, 256);
        switch(elem_bytes){
            case 4:  return L.launch(permute_simple_kernel<float>,   blocks, 256, 0, a);
            case 8:  return L.launch(permute_simple_kernel<double>,  blocks, 256, 0, a);
            case 16: return L.launch(permute_simple_kernel<double2>, blocks, 256, 0, a);
            default: return B200_ERR_INVALID;
        }
    }
    int const other_dim = 3 - tile_dim;
    a.nt = a.size[tile_dim]; a.no = a.size[other_dim];
    a.src_f = a.src_stride[0]; a.src_t = a.src_stride[tile_dim]; a.src_o = a.src_stride[other_dim];
    a.dst_t = a.dst_stride[tile_dim]; a.dst_o = a.dst_stride[other_dim];
    long long const gx = (a.size[0] + 31) / 32, gy = (a.nt + 31) / 32;
    long long const gz = std::min<long long>(a.no, 65535);
    if (gy > 65535) return B200_ERR_UNSUPPORTED;
    switch(elem_bytes){
        case 4:  return L.launch3(permute_tile_kernel<float>,   gx, gy, gz, 256, 32 * 33 * (size_t) elem_bytes, a);
        case 8:  return L.launch3(permute_tile_kernel<double>,  gx, gy, gz, 256, 32 * 33 * (size_t) elem_bytes, a);
        case 16: return L.launch3(permute_tile_kernel<double2>, gx, gy, gz, 256, 32 * 33 * (size_t) elem_bytes, a);
        default: return B200_ERR_INVALID;
    }
}

// heffte pack_plan_3d semantics (reference include/heffte_pack3d.h:31-45, 137-193)
inline permute_args transpose_unpack_args(long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                                          long long buff_line_stride, long long buff_plane_stride, int map0, int map1, int map2,
                                          const void *src, void *dst){
    permute_args a;
    a.src = src; a.dst = dst;
    a.size[0] = nfast; a.size[1] = nmid; a.size[2] = nslow;
    a.dst_stride[0] = 1; a.dst_stride[1] = line_stride; a.dst_stride[2] = plane_stride;
    long long const bstride[3] = {1, buff_line_stride, buff_plane_stride};
    int const map[3] = {map0, map1, map2};
    for(int k=0; k<3; k++) a.src_stride[map[k]] = bstride[k];
    a.nt = a.no = 0; a.src_f = a.src_t = a.src_o = a.dst_t = a.dst_o = 0;
    return a;
}

template<typename Launcher>
int launch_scale(int precision, long long count, void *data, double factor, Launcher &L){
    if (count <= 0) return B200_SUCCESS;
    scale_args a{data, count, factor};
    long long const blocks = stream_blocks(count, 256);
    if (precision == B200_PREC_FLOAT) return L.launch(scale_kernel<float>, blocks, 256, 0, a);
    return L.launch(scale_kernel<double>, blocks, 256, 0, a);
}

template<typename Launcher>
int launch_convert(int precision, bool to_complex, long long count, const void *src, void *dst, Launcher &L){
    if (count <= 0) return B200_SUCCESS;
    convert_args a{src, dst, count};
    long long const blocks = stream_blocks(count, 256);
    if (to_complex){
        if (precision == B200_PREC_FLOAT) return L.launch(real_to_complex_kernel<float, float2>, blocks, 256, 0, a);
        return L.launch(real_to_complex_kernel<double, double2>, blocks, 256, 0, a);
    }
    if (precision == B200_PREC_FLOAT) return L.launch(complex_to_real_kernel<float, float2>, blocks, 256, 0, a);
    return L.launch(complex_to_real_kernel<double, double2>, blocks, 256, 0, a);
}

} // namespace b200
