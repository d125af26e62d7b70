/*
 * heffte_b200_kernels.h -- C ABI of the device layer of the B200 backend (libheffte_b200.so).
 *
 * This is the thin layer the host C++ (heffte::backend::b200 plug-in, include/heffte_b200.hpp) calls instead of
 * cuFFT and the reference's own CUDA kernels.  Plain pointers, sizes and a cudaStream_t (passed as void*);
 * every function returns 0 on success or a non-zero error code (b200_last_error() gives the text) and never
 * throws.  All data pointers are DEVICE pointers.  There is no CPU fallback: without a CUDA device the calls fail.
 *
 * Each entry point cites the reference interface it replaces (paths relative to icl-utk-edu/heffte v2.4.1).
 */
#ifndef HEFFTE_B200_KERNELS_H
#define HEFFTE_B200_KERNELS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* precision / transform selectors */
#define B200_PREC_FLOAT  0
#define B200_PREC_DOUBLE 1

#define B200_FORWARD  0
#define B200_BACKWARD 1

/* kind of batched 1-D transform */
#define B200_C2C   0   /* complex <-> complex, in-place capable                                              */
#define B200_R2C   1   /* forward: real n -> complex n/2+1 ; backward: complex n/2+1 -> real n (unnormalised) */
#define B200_COS   2   /* forward REDFT10 (DCT-II), backward 2*REDFT01 (DCT-III)  == reference cufft_cos      */
#define B200_SIN   3   /* forward RODFT10 (DST-II), backward 2*RODFT01 (DST-III)  == reference cufft_sin      */
#define B200_COS1  4   /* forward REDFT00 (DCT-I),  backward 2*REDFT00            == reference cufft_cos1     */

/* error codes */
#define B200_SUCCESS            0
#define B200_ERR_INVALID        1
#define B200_ERR_UNSUPPORTED    2
#define B200_ERR_CUDA           3
#define B200_ERR_NCCL           4
#define B200_ERR_NO_DEVICE      5
#define B200_ERR_PEER           6   /* a peer GPU did not arrive within HEFFTE_B200_BARRIER_TIMEOUT_S (default: no limit) */

const char* b200_last_error(void);
/* number of kernels launched by this library since load (used by bench.py for "gpu_launches") */
long long b200_launch_count(void);
int b200_device_count(void);

/*
 * Device memory and transfers for callers that do not include the CUDA headers (the C++ front-end include/heffte_b200.hpp).
 * Replace heffte::backend::data_manipulator<tag::gpu> (include/heffte_backend_cuda.h:230-284) and heffte::gpu::transfer
 * (include/heffte_backend_data_transfer.h:29-183).  `stream` is a cudaStream_t passed as void* (NULL = default stream);
 * the copies are asynchronous with respect to the host unless the host memory is pageable.
 */
int b200_device_alloc(size_t bytes, void **device_pointer);
int b200_device_free(void *device_pointer);
int b200_copy_to_device(const void *host, void *device, size_t bytes, void *stream);
int b200_copy_to_host(const void *device, void *host, size_t bytes, void *stream);
int b200_copy_on_device(const void *source, void *destination, size_t bytes, void *stream);
/* synchronous copy between any two host / device pointers (what a GPU-aware MPI does with the buffers it is handed) */
int b200_copy_any(void *destination, const void *source, size_t bytes);
int b200_stream_synchronize(void *stream);
/* a non-blocking CUDA stream (cudaStream_t as void*).  Ranks that are host threads of one process (heffte_comm_create_threads)
 * need ONE STREAM PER RANK: the stream-ordered barrier between the ranks of a plan cannot complete on a shared stream.
 * Replaces the stream handling of heffte::backend::device_instance<tag::gpu> (include/heffte_backend_cuda.h:201-215). */
int b200_stream_create(void **stream);
int b200_stream_destroy(void *stream);
int b200_device_set(int device);

/*
 * Batched strided 1-D FFT plan.
 * Replaces: heffte::plan_cufft / plan_cufft_r2c (include/heffte_backend_cuda.h:346-422, 580-621), i.e.
 * cufftMakePlanMany(size, howmany, stride, dist) -- but with TWO batch dimensions so the "blocks" loop the
 * reference needs for the middle dimension (heffte_backend_cuda.h:452, 496-499) is a single launch, and with
 * separate input/output geometry so packing/transposition can be fused into the transform.
 *
 * Line l in [0, count_a*count_b): a = l % count_a, b = l / count_a.
 * Element i of line l is at  base + a*stride_a + b*stride_b + i*stride  (units: elements of that side's type:
 * complex for C2C, real on the real side of R2C / r2r).
 */
typedef struct {
    long long stride, stride_a, stride_b;
} b200_line_geom;

typedef struct {
    int precision;          /* B200_PREC_* */
    int kind;               /* B200_C2C ... */
    long long n;            /* transform length (real-space length for R2C and r2r) */
    long long count_a, count_b;
    b200_line_geom in;      /* geometry of the forward input  == backward output */
    b200_line_geom out;     /* geometry of the forward output == backward input  */
} b200_fft1d_desc;

typedef struct b200_fft1d_plan_s* b200_fft1d_plan;

int b200_fft1d_create(const b200_fft1d_desc *desc, b200_fft1d_plan *plan);
int b200_fft1d_destroy(b200_fft1d_plan plan);
/* Replaces cufftExec{C2C,Z2Z,R2C,D2Z,C2R,Z2D} (heffte_backend_cuda.h:494-524, 694-727) and, through `scale`,
 * the separate scaling kernel (src/heffte_backend_cuda.cu:138-145, 471-478).  in == out is allowed for C2C/r2r
 * when the two geometries coincide. */
int b200_fft1d_execute(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream);
/* The same transform restricted to the lines with b in [b_begin, b_begin + b_count) (b = line / count_a): a slab of the box.
 * `in` / `out` are the addresses of the whole box.  Lets the caller run two transforms slab by slab so that the second one
 * finds its input in the L2 cache (csrc/transform.cpp, HEFFTE_B200_L2_SLAB_MB). */
int b200_fft1d_execute_range(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream,
                             long long b_begin, long long b_count);
/*
 * Same transform with the FOLLOWING RESHAPE FUSED INTO THE STORE: every output element is written straight into the box
 * of the rank that owns it after the reshape -- in local memory or in a peer GPU's memory mapped over NVLink.
 * `device_scatter_map` points to a b200 scatter map in device memory (built by the plan, csrc/scatter_build.h).
 * Replaces, in one kernel: cufftExec* + direct_packer::pack + MPI_Alltoallv + direct/transpose_packer::unpack
 * (reference src/heffte_reshape3d.cpp:365-443, include/heffte_backend_cuda.h:494-524, 800-829).
 */
int b200_fft1d_execute_scatter(b200_fft1d_plan plan, int direction, const void *in, const void *device_scatter_map, double scale, void *stream);
/* Batched variants (reference include/heffte_fft3d.h:391-414: forward(batch, ...)): `batch` entries in ONE launch.  Entry e reads
 * in + e * in_step bytes and writes out + e * out_step bytes; with a fused reshape it adds e * scatter_step bytes to every
 * destination and local_shift + e * local_step bytes more to the destinations that lie in this rank's own memory (that is how
 * the last stage of a plan lands its own part straight in the caller's array). */
int b200_fft1d_execute_batch(b200_fft1d_plan plan, int direction, const void *in, void *out, double scale, void *stream,
                             int batch, long long in_step, long long out_step);
int b200_fft1d_execute_scatter_batch(b200_fft1d_plan plan, int direction, const void *in, const void *device_scatter_map, double scale, void *stream,
                                     int batch, long long in_step, long long scatter_step, long long local_shift, long long local_step);
/*
 * TWO consecutive transforms of the same box -- one along the contiguous axis, one along the middle axis -- in ONE persistent
 * launch, plane by plane: the second finds the output of the first in the L2 cache, and when it carries a fused reshape
 * (device_scatter_map != NULL) the first, purely local pass hides behind the NVLink-bound stores of the second.
 * first: in -> mid (may alias); second: mid -> the scatter map, or in place when the map is NULL; `scale` rides on the second.
 * counters: device memory of batch * (extent of the slowest axis) unsigned ints (zeroed by the call); lag: planes between the
 * two fronts (<= 0: default).  Returns B200_ERR_UNSUPPORTED -- without an error text -- when the two plans have no paired
 * kernel (b200_fft1d_pairable() == 0): run them one after the other instead.
 */
int b200_fft1d_pairable(b200_fft1d_plan first, b200_fft1d_plan second);
int b200_fft1d_execute_pair(b200_fft1d_plan first, b200_fft1d_plan second, int direction, const void *in, void *mid,
                            const void *device_scatter_map, double scale, void *counters, int lag, void *stream,
                            int batch, long long in_step, long long mid_step, long long scatter_step, long long local_shift, long long local_step);
/*
 * The same idea with TWO launches on TWO streams (what the plan uses): the first, local transform runs on `side_stream` and
 * reports the planes it has stored; the second transform -- with a fused reshape, bound by NVLink -- runs on `stream` with a thin
 * grid (`thin_blocks` CTAs walking the tiles) and waits plane by plane, so the HBM traffic of the first hides behind the remote
 * stores of the second (tools/kbench_peer.cu: a thin remote-store kernel and an HBM-bound kernel overlap almost perfectly).
 * fork_event / join_event: two cudaEvent_t of the caller, used to order side_stream after the work in front and to join it.  map_nb: the number of destination ranges of
 * the scatter map along the slowest axis (the producer visits the planes in the consumer's order).  Any two complex fast-path
 * transforms of the same box that are not along its slowest axis (b200_fft1d_overlappable() == 1).
 */
int b200_fft1d_overlappable(b200_fft1d_plan first, b200_fft1d_plan second);
int b200_fft1d_execute_overlapped(b200_fft1d_plan first, b200_fft1d_plan second, int direction, const void *in, void *mid,
                                  const void *device_scatter_map, int map_nb, double scale, void *counters, void *stream, void *side_stream,
                                  void *fork_event, void *join_event,
                                  int batch, long long in_step, long long mid_step, long long scatter_step, long long local_shift, long long local_step,
                                  int thin_blocks);
/*
 * Fused spectral operator along the axis of the plan: forward transform, spectrum * scale * M, backward transform of every line
 * in ONE pass over memory, M = the spectrum itself (multiplier == NULL, the x[i] *= x[i] of the reference's
 * benchmarks/convolution.cpp:89-94) or a device array with the layout of the input box.  The result goes to `out` (may alias
 * `in`) or, with a scatter map, through the fused reshape of a backward stage.  Complex plans along a strided axis with a
 * power-of-two length (b200_fft1d_convolvable() == 1), else B200_ERR_UNSUPPORTED without an error text.
 */
int b200_fft1d_convolvable(b200_fft1d_plan plan);
int b200_fft1d_execute_convolve(b200_fft1d_plan plan, const void *in, void *out, const void *device_scatter_map, const void *multiplier,
                                double scale, void *stream, int batch, long long in_step, long long out_step,
                                long long scatter_step, long long local_shift, long long local_step);
/* pointwise complex product data[i] = data[i] * factor * M[i] (M == NULL: data[i] itself), the unfused form of the operator above */
int b200_pointwise_multiply(int precision, long long count, void *data, const void *multiplier, double factor, void *stream);
/* name of the kernel family the plan resolved to ("strided", "contig", "generic"), for tests and profiling */
const char* b200_fft1d_kernel_name(b200_fft1d_plan plan);

/*
 * Sub-box copy between a strided box and a dense buffer.
 * Replaces heffte::cuda::direct_pack / direct_unpack (src/heffte_backend_cuda.cu:61-85, 385-402):
 *   pack:   dst[(s*nmid + m)*nfast + f] = src[s*plane_stride + m*line_stride + f]
 *   unpack: dst[s*plane_stride + m*line_stride + f] = src[(s*nmid + m)*nfast + f]
 * elem_bytes in {4, 8, 16}.
 */
int b200_direct_pack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                     long long line_stride, long long plane_stride, const void *src, void *dst, void *stream);
int b200_direct_unpack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                       long long line_stride, long long plane_stride, const void *src, void *dst, void *stream);
/* Sub-box copy between two strided boxes (no reference counterpart: used after a fused reshape to move the sub-boxes received
 * from other GPUs out of the plan's arena into the caller's array):  dst[s*dplane + m*dline + f] = src[s*splane + m*sline + f] */
int b200_copy_subbox(int elem_bytes, long long nfast, long long nmid, long long nslow,
                     long long src_line_stride, long long src_plane_stride, long long dst_line_stride, long long dst_plane_stride,
                     const void *src, void *dst, void *stream);
/*
 * Unpack with axis permutation.
 * Replaces heffte::cuda::transpose_unpack (src/heffte_backend_cuda.cu:90-133, 404-441):
 *   idx = (f, m, s);  dst[s*plane_stride + m*line_stride + f] = src[idx[map0] + idx[map1]*buff_line_stride + idx[map2]*buff_plane_stride]
 */
int b200_transpose_unpack(int elem_bytes, long long nfast, long long nmid, long long nslow,
                          long long line_stride, long long plane_stride,
                          long long buff_line_stride, long long buff_plane_stride,
                          int map0, int map1, int map2, const void *src, void *dst, void *stream);
/*
 * Reshape of a box without a transform in front of it (the first reshape of a plan), written straight into the destination
 * boxes through a scatter map: pack + transfer + unpack of the reference in one pass (src/heffte_reshape3d.cpp:365-443).
 */
int b200_scatter_copy(int elem_bytes, long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                      const void *src, const void *device_scatter_map, void *stream);
/* batched variant: `batch` entries in one launch, entry e reads src + e * in_step bytes and adds e * scatter_step bytes to every
 * destination (local_shift + e * local_step more to the destinations inside this rank's own memory) */
int b200_scatter_copy_batch(int elem_bytes, long long nfast, long long nmid, long long nslow, long long line_stride, long long plane_stride,
                            const void *src, const void *device_scatter_map, void *stream,
                            int batch, long long in_step, long long scatter_step, long long local_shift, long long local_step);
/* Several sub-boxes of one box copied to the same positions of another array of the same layout, all pieces and all batch
 * entries in ONE launch (16-byte accesses where the geometry allows): the pieces a rank received from the other GPUs move
 * from the plan's arena into the caller's array.  offsets / nfast / nmid / nslow: npieces entries, in elements. */
int b200_copy_subboxes(int elem_bytes, int npieces, const long long *offsets, const long long *nfast, const long long *nmid, const long long *nslow,
                       long long line_stride, long long plane_stride, const void *src, void *dst, void *stream,
                       int batch, long long src_step, long long dst_step);
/*
 * Stream-ordered barrier between the GPUs of a plan over peer memory (stands where the reference blocks the host in
 * MPI_Alltoallv / MPI_Waitany, src/heffte_reshape3d.cpp:388-402, 662): remote_slots[p] is the address, in rank p's flag
 * array, of the slot that belongs to rank `me`; local_flags is this rank's array; epoch increases by one per barrier.
 */
int b200_peer_barrier(int nranks, int me, void *const *remote_slots, void *local_flags, unsigned long long epoch, void *stream);
/* Waiting for a peer has NO time limit by default (a late rank is waited for, as the reference waits in MPI).  With
 * HEFFTE_B200_BARRIER_TIMEOUT_S=<seconds> an expired wait gives up without trapping and records itself in mapped host memory:
 * b200_peer_timed_out() then returns non-zero ((epoch << 8) | peer + 1) and every later transform call returns B200_ERR_PEER. */
unsigned long long b200_peer_timeout_ns(void);
unsigned long long* b200_peer_timeout_word(void);
unsigned long long b200_peer_timed_out(void);
/* Replaces heffte::cuda::scale_data (src/heffte_backend_cuda.cu:138-145, 471-478): data[i] *= factor over `count` reals. */
int b200_scale(int precision, long long count, void *data, double factor, void *stream);
/* Replaces heffte::cuda::convert (src/heffte_backend_cuda.cu:44-56, 352-361): real -> complex (zero imaginary) and complex -> real. */
int b200_convert_r2c(int precision, long long count, const void *real_src, void *complex_dst, void *stream);
int b200_convert_c2r(int precision, long long count, const void *complex_src, void *real_dst, void *stream);

#ifdef __cplusplus
}
#endif

#endif
