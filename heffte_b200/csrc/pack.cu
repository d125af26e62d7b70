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

int b200_scatter_copy(int elem_bytes, long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                      const void *src, const void *device_scatter_map, void *stream){
    if (device_scatter_map == nullptr) return fail(B200_ERR_INVALID, "null scatter map");
    if (nfast > 2147483647LL or nmid > 2147483647LL or nslow > 2147483647LL) return fail(B200_ERR_UNSUPPORTED, "box too large");
    cuda_launcher L{static_cast<cudaStream_t>(stream)};
    scatter_copy_args a{src, static_cast<const scatter_map*>(device_scatter_map), (int) nfast, (int) nmid, (int) nslow, line_stride, plane_stride, 1};
    int rc = launch_scatter_copy(elem_bytes, a, L);
    return (rc == B200_ERR_INVALID) ? fail(rc, "element size must be 4, 8 or 16 bytes") : rc;
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
