// C ABI of the packers, scaling and real<->complex conversion (see include/heffte_b200_kernels.h).
#include "pack_host.h"
#include <cstdlib>
#include "runtime.h"

using namespace b200;

extern "C" {

int b200_direct_pack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                     long long line_stride, long long plane_stride, const void *src, void *dst, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    copy3d_args a{src, dst, nfast, nmid, nslow, line_stride, plane_stride, nfast, nfast * nmid};
    int rc = launch_copy3d(elem_bytes, a, L);
    return (rc == B200_ERR_INVALID) ? fail(rc, "element size must be 4, 8 or 16 bytes") : rc;
}

int b200_direct_unpack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                       long long line_stride, long long plane_stride, const void *src, void *dst, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    copy3d_args a{src, dst, nfast, nmid, nslow, nfast, nfast * nmid, line_stride, plane_stride};
    int rc = launch_copy3d(elem_bytes, a, L);
    return (rc == B200_ERR_INVALID) ? fail(rc, "element size must be 4, 8 or 16 bytes") : rc;
}

int b200_copy_subbox(int elem_bytes, long long nfast, long long nmid, long long nslow,
                     long long src_line_stride, long long src_plane_stride, long long dst_line_stride, long long dst_plane_stride,
                     const void *src, void *dst, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    copy3d_args a{src, dst, nfast, nmid, nslow, src_line_stride, src_plane_stride, dst_line_stride, dst_plane_stride};
    int rc = launch_copy3d(elem_bytes, a, L);
    return (rc == B200_ERR_INVALID) ? fail(rc, "element size must be 4, 8 or 16 bytes") : rc;
}

int b200_transpose_unpack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                          long long line_stride, long long plane_stride,
                          long long buff_line_stride, long long buff_plane_stride,
                          int map0, int map1, int map2, const void *src, void *dst, void *stream){
    int seen[3] = {0, 0, 0};
    int const map[3] = {map0, map1, map2};
    for(int k=0; k<3; k++){
        if (map[k] < 0 or map[k] > 2) return fail(B200_ERR_INVALID, "map must be a permutation of 0,1,2");
        seen[map[k]]++;
    }
    if (seen[0] != 1 or seen[1] != 1 or seen[2] != 1) return fail(B200_ERR_INVALID, "map must be a permutation of 0,1,2");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    permute_args a = transpose_unpack_args(nfast, nmid, nslow, line_stride, plane_stride, buff_line_stride, buff_plane_stride,
                                           map0, map1, map2, src, dst);
    int rc = launch_permute(elem_bytes, a, L);
    if (rc == B200_ERR_INVALID) return fail(rc, "element size must be 4, 8 or 16 bytes");
    if (rc == B200_ERR_UNSUPPORTED) return fail(rc, "box too large for the tiled permutation");
    return rc;
}

int b200_scatter_copy_batch(int elem_bytes, long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                            const void *src, const void *device_scatter_map, void *stream,
                            int batch, long long in_step, long long scatter_step, long long local_shift, long long local_step){
    if (device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "null scatter map");
    if (nfast > 2147483647LL or nmid > 2147483647LL or nslow > 2147483647LL) return fail(B200_ERR_UNSUPPORTED, "box too large");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    L.batch = batch;
    scatter_copy_args a{src, static_cast<const scatter_map*>(device_scatter_map), (int) nfast, (int) nmid, (int) nslow, line_stride, plane_stride, 1,
                        in_step, scatter_step, local_shift, local_step};
    int rc = launch_scatter_copy(elem_bytes, a, L);
    return (rc == B200_ERR_INVALID) ? fail(rc, "element size must be 4, 8 or 16 bytes") : rc;
}
int b200_scatter_copy(int elem_bytes, long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                      const void *src, const void *device_scatter_map, void *stream){
    return b200_scatter_copy_batch(elem_bytes, nfast, nmid, nslow, line_stride, plane_stride, src, device_scatter_map, stream, 1, 0, 0, 0, 0);
}

int b200_copy_subboxes(int elem_bytes, int npieces, const long long *offsets, const long long *nfast, const long long *nmid, const long long *nslow,
                       long long line_stride, long long plane_stride, const void *src, void *dst, void *stream,
                       int batch, long long src_step, long long dst_step){
    if (npieces < 0 or (npieces > 0 and (offsets == nullptr or nfast == nullptr or nmid == nullptr or nslow == nullptr))) return fail(B200_ERR_INVALID, "bad piece list");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    for(int first = 0; first < npieces; first += multi_copy_max){
        multi_copy_args a{};
        a.src = src; a.dst = dst; a.line = line_stride; a.plane = plane_stride; a.src_step = src_step; a.dst_step = dst_step;
        for(int i = first; i < npieces and a.npieces < multi_copy_max; i++){
            if (nfast[i] <= 0 or nmid[i] <= 0 or nslow[i] <= 0) continue;
            if (nfast[i] > 2147483647LL or nmid[i] > 2147483647LL or nslow[i] > 2147483647LL) return fail(B200_ERR_UNSUPPORTED, "box too large");
            a.offset[a.npieces] = offsets[i];
            a.nfast[a.npieces] = static_cast<int>(nfast[i]); a.nmid[a.npieces] = static_cast<int>(nmid[i]); a.nslow[a.npieces] = static_cast<int>(nslow[i]);
            a.npieces++;
        }
        int rc = launch_multi_copy(elem_bytes, a, batch, L);
        if (rc == B200_ERR_INVALID) return fail(rc, "element size must be 4, 8 or 16 bytes");
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

// Limit of a wait for a peer, from HEFFTE_B200_BARRIER_TIMEOUT_S (seconds, fractions allowed).  Default 0: no limit -- a rank that
// is late entering forward() / backward() (host I/O between steps, a debugger, load imbalance) is simply waited for, as MPI does.
unsigned long long b200_peer_timeout_ns(void){
    static unsigned long long value = []{
        const char *e = std::getenv("HEFFTE_B200_BARRIER_TIMEOUT_S");
        double const seconds = (e != nullptr) ? std::atof(e) : 0.0;
        return (seconds > 0.0) ? static_cast<unsigned long long>(seconds * 1e9) : 0ULL;
    }();
    return value;
}
// One word of mapped, portable host memory per process: written by a device-side wait that expired.
unsigned long long* b200_peer_timeout_word(void){
#ifdef B200_HOST_EMULATION
    static unsigned long long word = 0;
    return &word;
#else
    static unsigned long long *word = []() -> unsigned long long* {
        void *p = nullptr;
        if (cudaHostAlloc(&p, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess){ cudaGetLastError(); return nullptr; }
        *static_cast<volatile unsigned long long*>(p) = 0;
        return static_cast<unsigned long long*>(p);
    }();
    return word;
#endif
}
unsigned long long b200_peer_timed_out(void){
    unsigned long long *w = b200_peer_timeout_word();
    return (w != nullptr) ? *static_cast<volatile unsigned long long*>(w) : 0ULL;
}

int b200_peer_barrier(int nranks, int me, void *const *remote_slots, void *local_flags, unsigned long long epoch, void *stream){
    if (nranks < 1 or nranks > barrier_max_ranks or me < 0 or me >= nranks) return fail(B200_ERR_INVALID, "bad rank count for the peer barrier");
    peer_barrier_args a{};
    for(int p=0; p<nranks; p++) a.remote[p] = static_cast<unsigned long long*>(remote_slots[p]);
    a.local = static_cast<unsigned long long*>(local_flags);
    a.epoch = epoch; a.nranks = nranks; a.me = me;
    a.timeout_ns = b200_peer_timeout_ns();
    a.timeout_word = b200_peer_timeout_word();
#ifdef B200_HOST_EMULATION
    return B200_SUCCESS;   // tests/emul: streams are synchronous and the thread-ranks rendezvous in after_peer_barrier()
#else
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    return L.launch(peer_barrier_kernel, 1, barrier_max_ranks, 0, a);
#endif
}

int b200_pointwise_multiply(int precision, long long count, void *data, const void *multiplier, double factor, void *stream){
    if (count <= 0) return B200_SUCCESS;
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    pointwise_args a{data, multiplier, count, factor};
    long long const blocks = stream_blocks(count, 256);
    if (precision == B200_PREC_FLOAT) return L.launch(pointwise_kernel<float2, float>, blocks, 256, 0, a);
    return L.launch(pointwise_kernel<double2, double>, blocks, 256, 0, a);
}

int b200_scale(int precision, long long count, void *data, double factor, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    return launch_scale(precision, count, data, factor, L);
}

int b200_convert_r2c(int precision, long long count, const void *real_src, void *complex_dst, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    return launch_convert(precision, true, count, real_src, complex_dst, L);
}
int b200_convert_c2r(int precision, long long count, const void *complex_src, void *real_dst, void *stream){
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    return launch_convert(precision, false, count, complex_src, real_dst, L);
}

} // extern "C"
