/*
 * heffte_b200.h -- the drop-in C ABI of the B200 backend: the same entry points, argument meaning, return codes
 * and struct layouts as the reference's C interface (icl-utk-edu/heffte v2.4.1, include/heffte_c.h:25-256 and
 * include/heffte_c_defines.h:54-162, implemented in src/heffte_c.cpp:193-498), for the new backend id
 * Heffte_BACKEND_B200.  This is what the reference's Python (python/heffte.py:42-100, ctypes) and Fortran bindings
 * load; INTEGRATION.md shows the binding a maintainer would add.
 *
 * One deliberate difference: the image has no MPI, and the data path is NCCL over NVLink, so the communicator
 * argument is a `heffte_comm` handle created by the functions at the top of this file instead of an MPI_Comm.
 * With a real MPI the maintainer-side stub builds the handle from the MPI communicator.
 *
 * All `input`/`output`/`workspace` pointers are DEVICE pointers, exactly like the reference's cuFFT backend.
 * The *_host variants at the end are the end-to-end convenience path (pinned-host staging + H2D/D2H inside).
 */
#ifndef HEFFTE_B200_H
#define HEFFTE_B200_H

#include "heffte_b200_kernels.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: reference include/heffte_c_defines.h:54-162 ------------------------------------------------ */
#define Heffte_BACKEND_STOCK   0
#define Heffte_BACKEND_FFTW    1
#define Heffte_BACKEND_MKL     2
#define Heffte_BACKEND_CUFFT  10
#define Heffte_BACKEND_ROCFFT 11
#define Heffte_BACKEND_B200   12   /* new: c2c and r2c */
#define Heffte_BACKEND_B200_COS   13   /* new: DCT-II/III  (reference template tag cufft_cos)  */
#define Heffte_BACKEND_B200_SIN   14   /* new: DST-II/III  (reference template tag cufft_sin)  */
#define Heffte_BACKEND_B200_COS1  15   /* new: DCT-I       (reference template tag cufft_cos1) */

#define Heffte_SUCCESS 0

#define Heffte_RESHAPE_ALGORITHM_ALLTOALLV  0
#define Heffte_RESHAPE_ALGORITHM_P2P_PLINED 1
#define Heffte_RESHAPE_ALGORITHM_P2P        2
#define Heffte_RESHAPE_ALGORITHM_ALLTOALL   3

#define Heffte_SCALE_NONE      0
#define Heffte_SCALE_FULL      1
#define Heffte_SCALE_SYMMETRIC 2

/* reference include/heffte_c_defines.h:113-125 */
typedef struct{
    int use_reorder;
    int algorithm;
    int use_pencils;
    int use_gpu_aware;
} heffte_plan_options;

/* reference include/heffte_c_defines.h:133-146 */
typedef struct{
    int backend_type;
    int using_r2c;
    void *fft;
} heffte_fft_plan;
typedef heffte_fft_plan* heffte_plan;

/* ---- communicator handle (stands where the reference takes an MPI_Comm) -------------------------------------- */
typedef struct heffte_comm_s* heffte_comm;

/* single rank, no communication library */
int heffte_comm_create_self(heffte_comm *comm);
/* one rank per GPU over NCCL: rank 0 obtains an id, the caller ships the 128 bytes to every rank (torch.distributed,
 * MPI_Bcast, a file ...), then every rank creates its communicator on its current CUDA device */
int heffte_comm_nccl_unique_id(void *id128);
int heffte_comm_create_nccl(int rank, int size, const void *id128, heffte_comm *comm);
/* caller-provided transport: host allgather for planning and a device exchange callback (see csrc/comm.h) */
typedef int (*heffte_allgather_fn)(void *context, const void *mine, void *all, size_t bytes);
typedef int (*heffte_exchange_fn)(void *context, int nsend, const int *send_peer, void *const *send_ptr, const size_t *send_bytes,
                                  int nrecv, const int *recv_peer, void *const *recv_ptr, const size_t *recv_bytes, void *stream);
int heffte_comm_create_callbacks(int rank, int size, heffte_allgather_fn gather, heffte_exchange_fn exchange, void *context, heffte_comm *comm);
/* `size` ranks inside ONE process, one host thread per rank, rank r on CUDA device devices[r] (NULL: all on device 0; the
 * same device may repeat -- this is how the test-suite runs multi-rank plans on one GPU).  Writes `size` handles; each must
 * be used from its own host thread and all plan calls are collective over the group. */
int heffte_comm_create_threads(int size, const int *devices, heffte_comm *comms);
int heffte_comm_rank(heffte_comm comm);
int heffte_comm_size(heffte_comm comm);
int heffte_comm_destroy(heffte_comm comm);

/* ---- the reference C API (include/heffte_c.h) ----------------------------------------------------------------------- */
/* heffte_plan_options::use_pencils: 0 (slabs) and 1 (pencils) are executed as given, like the reference; this value lets the planner
 * pick the decomposition that moves the fewest bytes over NVLink at the busiest GPU (also what a NULL options pointer means) */
#define Heffte_B200_DECOMPOSITION_AUTO 2
/* heffte_c.h:33  */ int heffte_set_default_options(int backend, heffte_plan_options *options);
/* heffte_c.h:67  */ int heffte_plan_create(int backend, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                                            int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                                            heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan);
/* heffte_c.h:77  */ int heffte_plan_create_r2c(int backend, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                                                int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                                                int r2c_direction, heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan);
/* same, on a caller-owned CUDA stream (C++ API: fft3d(stream, inbox, outbox, comm, options), include/heffte_fft3d.h:292-297) */
int heffte_plan_create_stream(int backend, void *cuda_stream, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                              int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                              int r2c_direction /* -1 unless r2c */, heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan);
/* same with the C++-only option plan_options::use_subcomm(num_subranks) (include/heffte_plan_logic.h:100-129): the intermediate
 * stages of the transform are held by the first num_subranks ranks only (small problems on many GPUs); <= 0 or >= the size of
 * the communicator means all ranks */
int heffte_plan_create_subcomm(int backend, void *cuda_stream, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                               int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                               int r2c_direction /* -1 unless r2c */, heffte_comm const comm, heffte_plan_options const *options,
                               int num_subranks, heffte_plan *plan);
/* same with 64-bit box coordinates (C++ API: box3d<long long>, test/test_longlong.cpp): nothing is truncated on the way in */
int heffte_plan_create64(int backend, void *cuda_stream, long long const inbox_low[3], long long const inbox_high[3], int const *inbox_order,
                         long long const outbox_low[3], long long const outbox_high[3], int const *outbox_order,
                         int r2c_direction /* -1 unless r2c */, heffte_comm const comm, heffte_plan_options const *options,
                         int num_subranks, heffte_plan *plan);
/* heffte_c.h:87  */ int heffte_plan_destroy(heffte_plan plan);
/* heffte_c.h:93  */ int heffte_size_inbox(heffte_plan const plan);
/* heffte_c.h:98  */ int heffte_size_outbox(heffte_plan const plan);
/* heffte_c.h:103 */ int heffte_size_workspace(heffte_plan const plan);
/* heffte_c.h:108 */ int heffte_get_backend(heffte_plan const plan);
/* heffte_c.h:113 */ int heffte_is_r2c(heffte_plan const plan);
/* 64-bit variants (the reference C API returns int) */
long long heffte_size_inbox64(heffte_plan const plan);
long long heffte_size_outbox64(heffte_plan const plan);
long long heffte_size_workspace64(heffte_plan const plan);
double heffte_get_scale_factor(heffte_plan const plan, int scale);

/* heffte_c.h:136-175 forward transforms, device pointers */
void heffte_forward_s2c(heffte_plan const plan, float const *input, void *output, int scale);
void heffte_forward_c2c(heffte_plan const plan, void const *input, void *output, int scale);
void heffte_forward_d2z(heffte_plan const plan, double const *input, void *output, int scale);
void heffte_forward_z2z(heffte_plan const plan, void const *input, void *output, int scale);
void heffte_forward_s2c_buffered(heffte_plan const plan, float const *input, void *output, void *workspace, int scale);
void heffte_forward_c2c_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale);
void heffte_forward_d2z_buffered(heffte_plan const plan, double const *input, void *output, void *workspace, int scale);
void heffte_forward_z2z_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale);
/* heffte_c.h:198-256 backward transforms */
void heffte_backward_c2s(heffte_plan const plan, void const *input, float *output, int scale);
void heffte_backward_c2c(heffte_plan const plan, void const *input, void *output, int scale);
void heffte_backward_z2d(heffte_plan const plan, void const *input, double *output, int scale);
void heffte_backward_z2z(heffte_plan const plan, void const *input, void *output, int scale);
void heffte_backward_c2s_buffered(heffte_plan const plan, void const *input, float *output, void *workspace, int scale);
void heffte_backward_c2c_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale);
void heffte_backward_z2d_buffered(heffte_plan const plan, void const *input, double *output, void *workspace, int scale);
void heffte_backward_z2z_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale);

/* real-to-real (the reference C API has no r2r entry points; these follow its naming): s = float, d = double */
void heffte_forward_s2s_buffered(heffte_plan const plan, float const *input, float *output, float *workspace, int scale);
void heffte_forward_d2d_buffered(heffte_plan const plan, double const *input, double *output, double *workspace, int scale);
void heffte_backward_s2s_buffered(heffte_plan const plan, float const *input, float *output, float *workspace, int scale);
void heffte_backward_d2d_buffered(heffte_plan const plan, double const *input, double *output, double *workspace, int scale);

/* generic entry used by the C++ header and the Python binding: precision B200_PREC_*, direction B200_FORWARD/BACKWARD,
 * batch >= 1 (C++ API forward(batch, ...), include/heffte_fft3d.h:391-414); returns an error code instead of void */
int heffte_execute(heffte_plan const plan, int precision, int direction, int batch, void const *input, void *output, void *workspace, int scale);
/*
 * Fused spectral operator (the caller pattern of the reference's benchmarks/convolution.cpp:86-97 -- forward(scale), pointwise
 * product, backward -- as ONE plan-level call, complex-to-complex plans):  output = backward( forward(input) * factor(scale) * M ),
 * M = the spectrum itself when multiplier == NULL (the benchmark's x[i] *= x[i]) or a device array laid out over the box
 * returned by heffte_convolve_box() (this rank's part of the spectrum in the plan's last forward stage, pencils along one axis).
 * The two brick reshapes around the product are not executed and the last forward transform, the product and the first
 * backward transform run in one kernel where the axis allows; the result comes back in the layout of the input box.
 */
int heffte_convolve(heffte_plan const plan, int precision, void const *input, void *output, void *workspace, void const *multiplier, int scale);
int heffte_convolve_box(heffte_plan const plan, long long low[3], long long high[3], int order[3]);
/* COLLECTIVE over the ranks of the plan: sets up the peer-memory data plane (arena of 3 x batch boxes, peer mapping, scatter maps)
 * for transforms of up to `batch` entries ahead of time.  Without it the first transform of each precision -- and the first one
 * with a larger batch -- does it: that call then blocks the host and must be entered by every rank with the same precision. */
int heffte_b200_prepare(heffte_plan const plan, int precision, int batch);
/* COLLECTIVE over the ranks of the plan: registers `bytes` of device memory the caller owns as an array the transforms of this plan
 * will be asked to WRITE (what ncclCommRegister is to NCCL; the reference has no counterpart: MPI moves the data there).  When a
 * transform of one entry (batch 1) gets a registered array as its output, the other GPUs store their part of the result straight
 * into it over NVLink and the final copy of the received sub-boxes disappears.  Contract: EVERY rank passes the array it registered
 * in the same call as the output of that transform (in place or not).  Returns Heffte_SUCCESS, or -- on every rank alike, with no
 * effect -- B200_ERR_UNSUPPORTED (2) when the memory cannot be shared with the other ranks (exchange mode, a memory pool without
 * CUDA IPC export, an array too small for this rank's box).  Registrations end with the plan or with a prepare() for more entries. */
int heffte_b200_register_buffer(heffte_plan const plan, int precision, void *device_array, size_t bytes);
/* local (not collective): forget a registered array, before its memory is released or handed to somebody else */
int heffte_b200_unregister_buffer(heffte_plan const plan, int precision, void *device_array);
/* same through pinned host staging: copies input host->device, transforms, copies the result device->host, synchronises */
int heffte_execute_host(heffte_plan const plan, int precision, int direction, int batch, void const *host_input, void *host_output, int scale);
/* 1 when the plan moves data between ranks through peer memory (NVLink stores fused into the FFT kernels), 0 when it uses
 * the communicator's send/receive path, -1 on a bad handle; meaningful after the first transform of that precision */
int heffte_b200_uses_peer_memory(heffte_plan const plan, int precision);
/* Device timing of the stages of the most recent transform in peer-memory mode (CUDA events on the plan's stream): enable,
 * run one transform, then collect (synchronises the stream).  Entry i: name (40 chars), milliseconds, bytes read + written in
 * local HBM and bytes stored into other GPUs over NVLink by that stage.  Returns the number of entries. */
int heffte_b200_stage_timing(heffte_plan const plan, int enable);
int heffte_b200_stage_times(heffte_plan const plan, int max_entries, char *names, double *ms, long long *local_bytes, long long *sent_bytes);
/* error text of the last failing call on this thread */
const char* heffte_last_error(void);

/* ---- plan introspection (pure host logic, usable without a GPU; used by the parity tests) ------------------------- */
/* boxes are 9 ints: low[3], high[3], order[3].  shapes_out receives 8*nranks boxes: in_shape[0..3] then out_shape[0..3]
 * (reference logic_plan3d, include/heffte_plan_logic.h:275-291; plan_operations, src/heffte_plan_logic.cpp:424-453) */
int heffte_b200_logic_plan(int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                           int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                           int *shapes_out, int *fft_direction, long long *index_count);
/* The plan a b200 transform EXECUTES for the same arguments: the reference's plan without the reorder of the intermediate boxes
 * and with those boxes assigned to the ranks so that the busiest GPU of every reshape moves as little as possible over NVLink
 * (csrc/plan_logic.h).  Results and reported sizes are those of the reference's plan; `swaps` receives the number of box
 * swaps the balancing applied.  HEFFTE_B200_REFERENCE_PLAN=1 makes both plans identical. */
int heffte_b200_execution_plan(int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                               int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                               int *shapes_out, int *fft_direction, int *swaps);
/* reference include/heffte_geometry.h:337-349, 643-691, 409-436 */
void heffte_b200_make_procgrid(int nprocs, int *grid2);
void heffte_b200_proc_setup_min_surface(int const *world_box, int nprocs, int *grid3);
void heffte_b200_split_world(int const *world_box, int const *grid3, int *boxes_out);
/* send (receive = 0) or receive (receive = 1) list of rank `me` for the reshape in -> out; pieces_out receives per entry
 * 14 long long: peer, offset, size[3], line, plane, buff_line, buff_plane, map[3], count, buffer_offset; returns the number
 * of entries or a negative error (reference src/heffte_reshape3d.cpp:125-206) */
int heffte_b200_reshape_pieces(int nranks, int const *inboxes, int const *outboxes, int me, int receive, long long *pieces_out, int max_pieces);
/* workspace / box sizes of a plan without creating device state (uses the same code path as the real plan) */
int heffte_b200_plan_sizes(int kind, int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                           int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                           long long *size_inbox, long long *size_outbox, long long *size_workspace);

#ifdef __cplusplus
}
#endif

#endif
