// The distributed 3-D transform object behind heffte::fft3d<backend::b200> / fft3d_r2c<backend::b200>.
// Replaces, for the b200 backend, the reference's fft3d::setup (include/heffte_fft3d.h:599-651), the reshape layer
// (include/heffte_reshape3d.h, src/heffte_reshape3d.cpp:365-443) and the stage driver
// (src/heffte_compute_transform.cpp:15-258).  Everything is enqueued on one CUDA stream; there is no host
// synchronisation between stages (the reference blocks on cudaStreamSynchronize before every MPI call,
// src/heffte_reshape3d.cpp:388).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "../../include/heffte_b200_kernels.h"
#include "comm.h"
#include "plan_logic.h"
#include "scatter_map.h"

namespace b200 {

enum transform_kind : int { kind_c2c = B200_C2C, kind_r2c = B200_R2C, kind_cos = B200_COS, kind_sin = B200_SIN, kind_cos1 = B200_COS1 };

struct piece {                 // one sub-box exchanged with one peer
    int peer;
    idx offset;                // position of the sub-box inside my box
    idx size[3];               // fast, mid, slow extents (in the order of MY box)
    idx line, plane;           // strides of my box
    idx buff_line, buff_plane; // receive side: strides of the packed message (sender order)
    int map[3];                // receive side: permutation (reference pack_plan_3d::map)
    bool permuted;             // receive side: orders differ
    idx count;
    idx buffer_offset;         // position of the message inside the send / receive buffer
};

// in_shape -> out_shape redistribution for one rank
class reshape_op {
public:
    reshape_op(shape const &in, shape const &out, int me, communicator *comm);
    idx input_count() const { return in_count; }
    idx output_count() const { return out_count; }
    idx workspace_count() const { return in_count + out_count; }
    bool is_local() const { return local_permute; }
    size_t num_peers_out() const { return sends.size(); }
    // src and dst may alias; workspace holds workspace_count() elements of elem_bytes
    int apply(int elem_bytes, const void *src, void *dst, void *workspace, cudaStream_t stream) const;
    std::vector<piece> const& send_list() const { return sends; }
    std::vector<piece> const& recv_list() const { return recvs; }
private:
    communicator *comm;
    int me;
    idx in_count = 0, out_count = 0;
    bool local_permute = false;
    std::vector<piece> sends, recvs;
};

// builds a reshape if the two shapes differ (null = no-op), mirroring the decision table of
// make_reshape3d (reference include/heffte_reshape3d.h:504-556)
std::unique_ptr<reshape_op> make_reshape(shape const &in, shape const &out, int me, communicator *comm);

class transform3d {
public:
    transform3d(transform_kind kind, box3 const &inbox, box3 const &outbox, int r2c_direction,
                communicator *comm, plan_options const &options, cudaStream_t stream);
    ~transform3d();

    idx size_inbox() const { return inbox_count; }
    idx size_outbox() const { return outbox_count; }
    idx size_workspace() const { return workspace_count; }
    double scale_factor(int scaling) const;
    logic_plan const& plan() const { return lp; }
    int traffic_swaps() const { return balanced_swaps; }
    transform_kind kind() const { return tkind; }
    int r2c_direction() const { return r2c_dir; }
    cudaStream_t stream() const { return cstream; }
    communicator* comm() const { return ccomm; }

    // precision B200_PREC_*; workspace may be null (a plan-owned buffer is used); scaling 0 none / 1 full / 2 symmetric
    int forward(int precision, int batch, const void *in, void *out, void *workspace, int scaling);
    int backward(int precision, int batch, const void *in, void *out, void *workspace, int scaling);
    // Fused spectral operator (reference benchmarks/convolution.cpp:86-97: forward(scale), pointwise product, backward), complex
    // plans only: out = backward( forward(in) * factor * M ) with M = the spectrum itself (multiplier == null: x *= x, the
    // benchmark's operator) or a caller array laid out in convolve_box().  The two brick reshapes around the product are elided
    // and the last forward transform, the product and the first backward transform run in ONE kernel where the plan allows.
    int convolve(int precision, const void *in, void *out, void *workspace, const void *multiplier, int scaling);
    // the box (of the plan's last forward stage) in which this rank provides / sees the spectrum during convolve()
    box3 convolve_box() const { return lp.out_shape[2][me]; }
    // collective: set up the peer-memory data plane for `batch` entries ahead of the first transform (else done lazily there)
    int prepare(int precision, int batch);

    // per-stage device timing of the most recent transform (peer-memory mode): enable, run, synchronise, collect
    struct stage_record { char name[40]; double ms; long long local_bytes; long long sent_bytes; };
    void enable_stage_timing(bool on){ timing = on; }
    std::vector<stage_record> collect_stage_times();

    // true once the plan runs its reshapes through peer memory (NVLink stores fused into the FFT kernels)
    bool uses_peer_memory(int precision) const { return peer[precision].active; }
    // Collective.  Registers `bytes` of caller memory at `ptr` as an array the transforms of this plan may be asked to write:
    // when forward / backward / convolve get it as their output, the other GPUs store their part of the result straight
    // into it (like a registered user buffer of a communication library) instead of into the plan's buffers, and the final
    // sub-box copy disappears.  EVERY rank must then pass the pointer it registered in the same call.  B200_ERR_UNSUPPORTED
    // (on every rank) when the memory cannot be shared or the plan does not run in peer-memory mode: nothing changes then.
    int register_buffer(int precision, void *ptr, size_t bytes);
    // local: forget a registered array (before its memory is released)
    int unregister_buffer(int precision, void *ptr);

private:
    enum run_mode { mode_forward = 0, mode_backward = 1, mode_convolve = 2 };
    int run(int precision, bool is_backward, const void *in, void *out, void *workspace, double scale);
    int run_local(int precision, int mode, int batch, const void *in, void *out, double scale, const void *multiplier);
    int run_peer(int precision, int mode, int batch, const void *in, void *out, double scale, const void *multiplier);
    int ensure_executors(int precision);
    void* ensure_workspace(int precision, int batch);
    bool ensure_peer(int precision, int batch);
    void release_peer(int precision);
    int peer_fence(int precision);
    void* pair_counters(size_t count);
    bool all_local() const;

    // ---- peer-memory mode: every reshape is fused into the store of the transform in front of it --------------------------
    // (stage 0 = the first reshape, a scatter copy; stage s = 1..3: transform s-1 followed by reshape s)
    struct peer_state {
        bool tried = false, active = false;
        void *arena = nullptr;                 // [flags | buffer 0 | buffer 1 | buffer 2], peer-mapped on every rank of the plan
        size_t entry_bytes = 0;                // room of one batch entry inside a buffer: the largest box of the plan
        int capacity = 0;                      // batch entries a buffer holds
        size_t buffer_bytes = 0;               // entry_bytes * capacity
        std::vector<void*> arenas;             // address of every rank's arena as seen from this device
        std::vector<void*> remote_slots;       // my slot in every rank's flag array
        unsigned long long epoch = 0;
        void *maps = nullptr;                  // device array of scatter maps, index ((view * 4 + stage) * 3 + buffer)
        std::vector<scatter_map> host_maps;    // host copy (the plan reads the cell counts)
        int next_buffer = 0;                   // the three buffers rotate: every remote write targets the buffer after the last one used
        bool fused[3][4] = {{false, false, false, false}, {false, false, false, false}, {false, false, false, false}};   // [view][stage]: the reshape moves data (global fact)
        char* buffer(int index) const { return static_cast<char*>(arena) + 4096 + static_cast<size_t>(index) * buffer_bytes; }
        int take(){ int const w = next_buffer; next_buffer = (next_buffer + 1) % 3; return w; }
        // caller arrays the peers may write directly: one scatter map of the last stage per view
        struct registered { void *ptr = nullptr; size_t bytes = 0; void *maps = nullptr; bool has[3] = {false, false, false}; };
        std::vector<registered> user;
    };
    bool stage_map(int precision, int view, int st, std::vector<void*> const &bases, scatter_map &map, std::string &why, bool count);
    peer_state peer[2];
    long long sent_elems[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};     // elements that leave this GPU in stage (view, st)
    long long stage_elems[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};    // elements this rank writes in that stage
    // The stages of a run as a forward-shaped sequence (reshape s, then the transform of stage s).  view 0: the forward transform;
    // view 1: the backward transform as the mirror image of the forward plan (what r2c plans and the spectral operator need);
    // view 2: the backward transform planned like a forward transform from the output boxes to the input boxes (`lb`): the
    // purely local transforms then come FIRST, where they overlap with the transfers of the stage behind them, instead of last.
    enum { view_forward = 0, view_mirror = 1, view_backward = 2 };
    logic_plan lb;
    bool lb_active = false;
    b200_fft1d_plan bexec[2][3];
    shape const& vin(int view, int s) const { return (view == view_forward) ? lp.in_shape[s] : ((view == view_mirror) ? lp.out_shape[3-s] : lb.in_shape[s]); }
    shape const& vout(int view, int s) const { return (view == view_forward) ? lp.out_shape[s] : ((view == view_mirror) ? lp.in_shape[3-s] : lb.out_shape[s]); }
    int vdim(int view, int e) const { return (view == view_forward) ? lp.fft_direction[e] : ((view == view_mirror) ? lp.fft_direction[2-e] : lb.fft_direction[e]); }
    int real_id(int view, int e) const { return (view == view_mirror) ? 2 - e : e; }      // which executor of the r2c chain this is
    b200_fft1d_plan vexec(int precision, int view, int e) const {
        return (view == view_forward) ? exec[precision][e] : ((view == view_mirror) ? exec[precision][2-e] : bexec[precision][e]);
    }
    bool timing = false;
    std::vector<cudaEvent_t> marks;
    std::vector<stage_record> pending;
    void mark(const char *name, long long local_bytes, long long sent_bytes);
    void *counters = nullptr;           // per-plane counters of two overlapped launches
    size_t counters_count = 0;
    cudaStream_t side_stream = nullptr; // the local transform of an overlapped pair runs here
    cudaEvent_t fork_event = nullptr, join_event = nullptr;
    bool ensure_side_stream();

    transform_kind tkind;
    int r2c_dir;
    communicator *ccomm;
    cudaStream_t cstream;
    logic_plan lp;
    int me;
    idx inbox_count, outbox_count, comm_count, temp_count;
    idx workspace_count;            // what size_workspace() reports: the reference's figure for this geometry
    idx exec_workspace_count;       // what the executed plan needs on the exchange path
    int balanced_swaps = 0;         // box swaps applied by balance_traffic()
    idx workspace_layout(logic_plan const &p, idx &comm_elements, idx &temp_elements) const;
    idx l2_slab_planes(int first, int second, int elem_bytes) const;
    double base_scale;
    std::unique_ptr<reshape_op> fwd[4], bwd[4];
    b200_fft1d_plan exec[2][3];
    bool exec_ready[2] = {false, false};
    void *own_workspace = nullptr;
    size_t own_workspace_bytes = 0;
};

} // namespace b200
