#include "transform.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scatter_build.h"

namespace b200 {

// defined in fft1d.cu
void set_error(std::string const &message);
int fail(int code, std::string const &message);

namespace {

inline char* advance(void *p, idx elements, int elem_bytes){ return static_cast<char*>(p) + elements * elem_bytes; }
inline const char* advance(const void *p, idx elements, int elem_bytes){ return static_cast<const char*>(p) + elements * elem_bytes; }

// geometry of the batch of lines of `box` that run along `dim`.  The two other axes are kept apart
// (a = the faster one, b = the slower one) so that a line knows its box coordinates: the fused reshape needs them.
void line_layout(box3 const &box, int dim, b200_line_geom &g, long long &count_a, long long &count_b){
    idx const strides[3] = {1, box.osize(0), box.osize(0) * box.osize(1)};
    int const pos = box.position_of(dim);
    int const a_pos = (pos == 0) ? 1 : 0, b_pos = (pos == 2) ? 1 : 2;
    g.stride = strides[pos]; g.stride_a = strides[a_pos]; g.stride_b = strides[b_pos];
    count_a = box.osize(a_pos); count_b = box.osize(b_pos);
}

bool shapes_differ(shape const &a, shape const &b){
    return not (extents_match(a, b) and a[0].same_order(b[0]));
}

} // namespace

// ------------------------------------------------------------------------------------------------------------
// reshape
// ------------------------------------------------------------------------------------------------------------
reshape_op::reshape_op(shape const &in, shape const &out, int rank, communicator *c) : comm(c), me(rank){
    box3 const &mine_in = in[me], &mine_out = out[me];
    in_count = mine_in.count();
    out_count = mine_out.count();
    int const n = static_cast<int>(in.size());

    auto receive_piece = [&](int peer, box3 const &source_box){
        box3 ov = mine_out.overlap(source_box);
        piece p{};
        p.peer = peer;
        p.offset = mine_out.offset_of(ov.low);
        for(int d=0; d<3; d++) p.size[d] = ov.osize(d);
        p.line = mine_out.osize(0);
        p.plane = mine_out.osize(0) * mine_out.osize(1);
        p.permuted = not source_box.same_order(mine_out);
        p.buff_line = ov.size(source_box.order[0]);
        p.buff_plane = ov.size(source_box.order[0]) * ov.size(source_box.order[1]);
        for(int j=0; j<3; j++) p.map[j] = mine_out.position_of(source_box.order[j]);
        p.count = ov.count();
        return p;
    };

    if (extents_match(in, out)){
        // same boxes, new order: a local permutation of my own data
        local_permute = true;
        if (not mine_out.empty()) recvs.push_back(receive_piece(me, mine_in));
        return;
    }

    idx send_offset = 0, recv_offset = 0;
    for(int i=0; i<n; i++){
        int const peer = (i + me + 1) % n;      // same visiting order as the reference: self comes last
        box3 ov = mine_in.overlap(out[peer]);
        if (not ov.empty()){
            piece p{};
            p.peer = peer;
            p.offset = mine_in.offset_of(ov.low);
            for(int d=0; d<3; d++) p.size[d] = ov.osize(d);
            p.line = mine_in.osize(0);
            p.plane = mine_in.osize(0) * mine_in.osize(1);
            p.count = ov.count();
            p.buffer_offset = send_offset;
            send_offset += p.count;
            sends.push_back(p);
        }
        box3 ov_in = mine_out.overlap(in[peer]);
        if (not ov_in.empty()){
            piece p = receive_piece(peer, in[peer]);
            p.buffer_offset = recv_offset;
            recv_offset += p.count;
            recvs.push_back(p);
        }
    }
}

int reshape_op::apply(int elem_bytes, const void *src, void *dst, void *workspace, cudaStream_t stream) const {
    auto unpack = [&](piece const &p, const void *buffer) -> int {
        void *target = advance(dst, p.offset, elem_bytes);
        if (p.permuted)
            return b200_transpose_unpack(elem_bytes, p.size[0], p.size[1], p.size[2], p.line, p.plane, p.buff_line, p.buff_plane,
                                         p.map[0], p.map[1], p.map[2], buffer, target, stream);
        return b200_direct_unpack(elem_bytes, p.size[0], p.size[1], p.size[2], p.line, p.plane, buffer, target, stream);
    };

    if (local_permute){
        if (recvs.empty()) return B200_SUCCESS;
        const void *from = src;
        if (src == dst){
            if (cudaMemcpyAsync(workspace, src, static_cast<size_t>(in_count) * elem_bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed in local reshape");
            from = workspace;
        }
        return unpack(recvs[0], from);
    }

    char *send_buffer = static_cast<char*>(workspace);
    char *recv_buffer = advance(workspace, in_count, elem_bytes);
    const void *self_message = nullptr;

    std::vector<transfer> outgoing, incoming;
    for(auto const &p : sends){
        char *slot = send_buffer + p.buffer_offset * elem_bytes;
        int rc = b200_direct_pack(elem_bytes, p.size[0], p.size[1], p.size[2], p.line, p.plane,
                                  advance(src, p.offset, elem_bytes), slot, stream);
        if (rc) return rc;
        if (p.peer == me) self_message = slot;
        else outgoing.push_back({p.peer, slot, static_cast<size_t>(p.count) * elem_bytes});
    }
    for(auto const &p : recvs)
        if (p.peer != me) incoming.push_back({p.peer, recv_buffer + p.buffer_offset * elem_bytes, static_cast<size_t>(p.count) * elem_bytes});

    {   // collective over the ranks of the plan, also with nothing to send or receive
        int rc = comm->exchange(outgoing, incoming, stream);
        if (rc) return fail(B200_ERR_NCCL, "exchange failed in reshape");
    }
    for(auto const &p : recvs){
        const void *message = (p.peer == me) ? self_message : recv_buffer + p.buffer_offset * elem_bytes;
        if (message == nullptr) return fail(B200_ERR_INVALID, "inconsistent self overlap in reshape");
        int rc = unpack(p, message);
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

std::unique_ptr<reshape_op> make_reshape(shape const &in, shape const &out, int me, communicator *comm){
    if (extents_match(in, out)){
        if (in[0].same_order(out[0])) return nullptr;
        if (out[me].empty()) return nullptr;
    }
    return std::unique_ptr<reshape_op>(new reshape_op(in, out, me, comm));
}

// ------------------------------------------------------------------------------------------------------------
// transform
// ------------------------------------------------------------------------------------------------------------
transform3d::transform3d(transform_kind kind, box3 const &inbox, box3 const &outbox, int r2c_direction,
                         communicator *comm, plan_options const &options, cudaStream_t stream)
    : tkind(kind), r2c_dir((kind == kind_r2c) ? r2c_direction : -1), ccomm(comm), cstream(stream){
    for(int p=0; p<2; p++) for(int i=0; i<3; i++) exec[p][i] = nullptr;
    me = comm->rank();
    int const n = comm->size();
#ifndef B200_HOST_EMULATION
    // ranks that share one process share the default stream: the stream-ordered barrier between them could never complete
    if (n > 1 and stream == nullptr and std::strcmp(comm->kind(), "threads") == 0)
        throw std::runtime_error("ranks that are host threads of one process need one CUDA stream per rank (create the plan on a stream)");
#endif

    // plan-time allgather of (inbox, outbox): 18 64-bit integers per rank (reference include/heffte_geometry.h:707-718)
    std::vector<long long> mine(18), all(18 * static_cast<size_t>(n));
    for(int d=0; d<3; d++){
        mine[d] = inbox.low[d]; mine[3+d] = inbox.high[d]; mine[6+d] = inbox.order[d];
        mine[9+d] = outbox.low[d]; mine[12+d] = outbox.high[d]; mine[15+d] = outbox.order[d];
    }
    if (comm->allgather(mine.data(), all.data(), 18 * sizeof(long long)) != 0)
        throw std::runtime_error("allgather of the boxes failed");
    shape ins, outs;
    for(int r=0; r<n; r++){
        long long const *b = all.data() + 18 * r;
        ins.push_back(box3({{b[0], b[1], b[2]}}, {{b[3], b[4], b[5]}}, {{(int) b[6], (int) b[7], (int) b[8]}}));
        outs.push_back(box3({{b[9], b[10], b[11]}}, {{b[12], b[13], b[14]}}, {{(int) b[15], (int) b[16], (int) b[17]}}));
    }

    plan_options effective = options;
    // the cosine / sine executors work on contiguous lines only in the reference (include/heffte_plan_logic.h:206-224), which
    // therefore forces the reorder; the reference-shaped plan below keeps that so that the reported sizes agree
    if (kind == kind_cos or kind == kind_sin or kind == kind_cos1) effective.use_reorder = true;
    // (1) the reference's plan, box for box: it defines what the caller sees (size_workspace, tests/test_plan_logic.py)
    logic_plan const reference_plan = make_logic_plan(ins, outs, r2c_dir, effective, me);
    idx ref_comm = 0, ref_temp = 0;
    workspace_count = workspace_layout(reference_plan, ref_comm, ref_temp);
    // (2) the plan that is executed (plan_logic.h: no reorder of the intermediate boxes, traffic balancing)
    lp = make_execution_plan(ins, outs, r2c_dir, effective, me, &balanced_swaps);

    // backward transforms of complex and real-to-real plans: planned like a forward transform from the output boxes to the
    // input boxes (see transform.h, view_backward); r2c plans keep the mirror image (the real transform must come last)
    for(int p=0; p<2; p++) for(int i=0; i<3; i++) bexec[p][i] = nullptr;
    if (kind != kind_r2c and n > 1 and std::getenv("HEFFTE_B200_MIRRORED_BACKWARD") == nullptr){
        try{
            lb = make_execution_plan(outs, ins, -1, effective, me);
            lb_active = true;
        }catch(std::exception &){ lb_active = false; }
    }
    {   // the executed plan depends on per-process switches (HEFFTE_B200_REFERENCE_PLAN, HEFFTE_B200_DECOMPOSITION): every rank
        // must have arrived at the same boxes, or the fused stores would land at wrong addresses
        unsigned long long h = 1469598103934665603ULL;
        auto mix = [&](long long v){ h ^= static_cast<unsigned long long>(v); h *= 1099511628211ULL; };
        for(int s=0; s<4; s++)
            for(shape const *sh : {&lp.in_shape[s], &lp.out_shape[s]})
                for(box3 const &b : *sh) for(int d=0; d<3; d++){ mix(b.low[d]); mix(b.high[d]); mix(b.order[d]); }
        for(int d=0; d<3; d++) mix(lp.fft_direction[d]);
        mix(lb_active ? 1 : 0);
        if (lb_active){
            for(int s=0; s<4; s++)
                for(shape const *sh : {&lb.in_shape[s], &lb.out_shape[s]})
                    for(box3 const &b : *sh) for(int d=0; d<3; d++){ mix(b.low[d]); mix(b.high[d]); mix(b.order[d]); }
            for(int d=0; d<3; d++) mix(lb.fft_direction[d]);
        }
        std::vector<unsigned long long> all_hashes(static_cast<size_t>(n));
        if (comm->allgather(&h, all_hashes.data(), sizeof(h)) != 0) throw std::runtime_error("allgather of the plan signature failed");
        for(unsigned long long v : all_hashes)
            if (v != h) throw std::runtime_error("the ranks planned different transforms (do HEFFTE_B200_REFERENCE_PLAN / HEFFTE_B200_DECOMPOSITION differ between the ranks?)");
    }
    inbox_count = lp.in_shape[0][me].count();
    outbox_count = lp.out_shape[3][me].count();
    base_scale = 1.0 / static_cast<double>(lp.index_count);
    if (kind == kind_cos or kind == kind_sin) base_scale /= 64.0;
    if (kind == kind_cos1) base_scale = 1.0 / (64.0 * (lp.fft_sizes[0] - 1) * (lp.fft_sizes[1] - 1) * (lp.fft_sizes[2] - 1));

    for(int i=0; i<4; i++){
        fwd[i] = make_reshape(lp.in_shape[i], lp.out_shape[i], me, comm);
        bwd[3-i] = make_reshape(lp.out_shape[i], lp.in_shape[i], me, comm);
    }
    // the exchange path lays its buffers out like the reference (below); when the executed plan needs more room than the
    // reference's (reported) workspace, the plan uses a buffer of its own instead of the caller's
    exec_workspace_count = workspace_layout(lp, comm_count, temp_count);
}

// workspace layout (reference include/heffte_fft3d.h:625-633, include/heffte_fft3d_r2c.h:335-339): returns the total
idx transform3d::workspace_layout(logic_plan const &p, idx &comm_elements, idx &temp_elements) const {
    auto moves = [&](shape const &in, shape const &out){        // same decision as make_reshape
        if (extents_match(in, out)){
            if (in[0].same_order(out[0])) return false;
            if (out[me].empty()) return false;
        }
        return true;
    };
    comm_elements = 0;
    bool last_backward = false;
    for(int i=0; i<4; i++){
        if (moves(p.in_shape[i], p.out_shape[i])) comm_elements = std::max(comm_elements, p.in_shape[i][me].count() + p.out_shape[i][me].count());
        if (moves(p.out_shape[i], p.in_shape[i])){
            comm_elements = std::max(comm_elements, p.in_shape[i][me].count() + p.out_shape[i][me].count());
            if (i == 0) last_backward = true;      // bwd[3]: the last reshape of the backward transform
        }
    }
    temp_elements = 0;
    for(int i=0; i<3; i++){
        idx boxed = (i == 0 and tkind == kind_r2c) ? p.in_shape[1][me].count() : p.out_shape[i][me].count();
        temp_elements = std::max(temp_elements, boxed);
    }
    idx last_chunk = 0;
    if (tkind != kind_r2c and last_backward) last_chunk = (p.out_shape[0][me].count() + 1) / 2;
    return comm_elements + temp_elements + last_chunk;
}

void transform3d::release_peer(int precision){
    peer_state &P = peer[precision];
    if (P.arena){
        cudaStreamSynchronize(cstream);
        if (not P.arenas.empty()) ccomm->unmap_peers(P.arenas);
        cudaFree(P.arena);
        P.arena = nullptr;
    }
    P.arenas.clear();
    if (P.maps){ cudaFree(P.maps); P.maps = nullptr; }
    for(auto &u : P.user) if (u.maps) cudaFree(u.maps);
    P.user.clear();                                // registrations do not survive a rebuilt arena: the caller registers again
    P.active = false;
}

transform3d::~transform3d(){
    for(int p=0; p<2; p++) release_peer(p);
    for(int p=0; p<2; p++) for(int i=0; i<3; i++) if (exec[p][i]) b200_fft1d_destroy(exec[p][i]);
    for(int p=0; p<2; p++) for(int i=0; i<3; i++) if (bexec[p][i]) b200_fft1d_destroy(bexec[p][i]);
    if (own_workspace) cudaFree(own_workspace);
    if (counters) cudaFree(counters);
    if (side_stream){ cudaStreamSynchronize(side_stream); cudaStreamDestroy(side_stream); }
    if (fork_event) cudaEventDestroy(fork_event);
    if (join_event) cudaEventDestroy(join_event);
    for(cudaEvent_t e : marks) cudaEventDestroy(e);
}

// ------------------------------------------------------------------------------------------------------------
// peer-memory mode
// ------------------------------------------------------------------------------------------------------------
// Collective over the ranks of the plan (first transform of each precision, or prepare(); again when a larger batch arrives):
// allocate and peer-map the arena -- three rotating buffers of `batch` entries -- and build the scatter maps of every stage.
// Any rank failing any step makes every rank fall back to the exchange() path.
bool transform3d::ensure_peer(int precision, int batch){
    peer_state &P = peer[precision];
    batch = std::max(batch, 1);
    if (P.tried and (not P.active or batch <= P.capacity)) return P.active;
    int const n = ccomm->size();
    if (n < 2 or n > 64){ P.tried = true; return false; }
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    bool any = false;
    int const nviews = lb_active ? 3 : 2;
    for(int s=0; s<4; s++){
        for(int v=0; v<nviews; v++) P.fused[v][s] = shapes_differ(vin(v, s), vout(v, s));
        any = any or P.fused[0][s];
    }
    if (not any){ P.tried = true; return false; }
    if (P.tried) release_peer(precision);          // a larger batch: every rank rebuilds (the call is collective)
    P.tried = true;                                // the flags of a new arena start at zero on every rank: so does the epoch
    P.epoch = 0;
    P.next_buffer = 0;

    // an entry holds the largest box any rank ever owns, in the widest element type of the plan (the same size on every
    // rank: the buffers of a peer sit at known offsets inside its arena)
    idx largest = 1;
    for(int s=0; s<4; s++)
        for(int r=0; r<n; r++){
            largest = std::max(largest, std::max(lp.in_shape[s][r].count(), lp.out_shape[s][r].count()));
            if (lb_active) largest = std::max(largest, std::max(lb.in_shape[s][r].count(), lb.out_shape[s][r].count()));
        }
    P.entry_bytes = ((static_cast<size_t>(largest) * (complex_data ? cplx_bytes : real_bytes) + 255) / 256) * 256;
    P.capacity = batch;
    P.buffer_bytes = P.entry_bytes * static_cast<size_t>(batch);
    size_t const arena_bytes = 4096 + 3 * P.buffer_bytes;
    int ok = 1;
    if (cudaMalloc(&P.arena, arena_bytes) != cudaSuccess){ P.arena = nullptr; cudaGetLastError(); ok = 0; }
    if (ok and (cudaMemset(P.arena, 0, 4096) != cudaSuccess or cudaDeviceSynchronize() != cudaSuccess)) ok = 0;
    // map_peers is collective: it is called by every rank even after a local failure (with a harmless null allocation vote)
    std::vector<void*> arenas;
    bool mapped = false;
    {
        std::vector<int> votes(n);
        if (ccomm->allgather(&ok, votes.data(), sizeof(int)) != 0) ok = 0;
        for(int v : votes) if (not v) ok = 0;
        if (ok) mapped = ccomm->map_peers(P.arena, arena_bytes, arenas);
    }
    if (not mapped){
        if (P.arena){ cudaFree(P.arena); P.arena = nullptr; }
        return false;
    }
    P.arenas = arenas;
    P.remote_slots.resize(n);
    for(int r=0; r<n; r++) P.remote_slots[r] = static_cast<char*>(arenas[r]) + sizeof(unsigned long long) * me;

    // scatter maps: ((view * 4 + stage) * 3 + buffer)
    std::vector<scatter_map> maps(36);
    std::string why;
    int built = 1;
    for(int view=0; view<nviews and built; view++){
        for(int st=0; st<4 and built; st++){
            if (not P.fused[view][st]) continue;
            for(int w=0; w<3; w++){
                std::vector<void*> bases(n);
                for(int r=0; r<n; r++) bases[r] = static_cast<char*>(arenas[r]) + 4096 + static_cast<size_t>(w) * P.buffer_bytes;
                if (not stage_map(precision, view, st, bases, maps[(view * 4 + st) * 3 + w], why, w == 0)){ built = 0; break; }
            }
        }
    }
    if (built and cudaMalloc(&P.maps, maps.size() * sizeof(scatter_map)) != cudaSuccess){ P.maps = nullptr; cudaGetLastError(); built = 0; }
    if (built and cudaMemcpy(P.maps, maps.data(), maps.size() * sizeof(scatter_map), cudaMemcpyHostToDevice) != cudaSuccess) built = 0;
    {
        std::vector<int> votes(n);
        if (ccomm->allgather(&built, votes.data(), sizeof(int)) != 0) built = 0;
        for(int v : votes) if (not v) built = 0;
    }
    if (not built){
        release_peer(precision);
        return false;
    }
    P.host_maps = maps;
    P.active = true;
    return true;
}

// the scatter map of stage (view, st) for destination boxes that start at bases[rank]
bool transform3d::stage_map(int precision, int view, int st, std::vector<void*> const &bases, scatter_map &map, std::string &why, bool count){
    int const n = ccomm->size();
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    shape const &dest = vout(view, st);
    // the box this rank writes in that stage, the axis of the transform in front of the reshape, the element size
    box3 const written = vin(view, st)[me];
    int k_pos = 0, bytes = complex_data ? cplx_bytes : real_bytes;
    if (st == 0){
        if (tkind == kind_r2c and view == view_forward) bytes = real_bytes;
    }else{
        int const e = st - 1;                                       // transform in front of this reshape
        if (not written.empty()) k_pos = written.position_of(vdim(view, e));
        if (tkind == kind_r2c and view == view_mirror and real_id(view, e) == 0) bytes = real_bytes;   // c2r output
    }
    if (count){
        stage_elems[view][st] = written.count();
        sent_elems[view][st] = 0;
        for(int r=0; r<n; r++) if (r != me) sent_elems[view][st] += written.overlap(dest[r]).count();
    }
    return build_scatter_map(written, k_pos, dest, bases, bytes, map, why, nullptr, me);
}

int transform3d::register_buffer(int precision, void *ptr, size_t bytes){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    int rc = ensure_executors(precision);
    if (rc) return rc;
    if (not ensure_peer(precision, 1)) return B200_ERR_UNSUPPORTED;        // the same answer on every rank
    peer_state &P = peer[precision];
    int const n = ccomm->size();
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    std::vector<void*> peers;
    bool const mapped = ccomm->map_user_buffer(ptr, peers);                // collective
    // which views can end in this array: the last reshape moves data, and the array holds the box that lands in it (on every rank)
    int const nviews = lb_active ? 3 : 2;
    int fits[3] = {0, 0, 0};
    for(int v=0; v<nviews and mapped; v++){
        if (not P.fused[v][3]) continue;
        bool const real_out = (tkind == kind_r2c and v != view_forward) or not complex_data;
        size_t const need = static_cast<size_t>(vout(v, 3)[me].count()) * (real_out ? real_bytes : cplx_bytes);
        fits[v] = (bytes >= need) ? 1 : 0;
    }
    std::vector<int> all(3 * static_cast<size_t>(n));
    if (ccomm->allgather(fits, all.data(), sizeof(fits)) != 0) return fail(B200_ERR_PEER, "allgather failed");
    for(int r=0; r<n; r++) for(int v=0; v<3; v++) if (not all[3 * static_cast<size_t>(r) + v]) fits[v] = 0;
    peer_state::registered entry;
    entry.ptr = ptr; entry.bytes = bytes;
    std::vector<scatter_map> maps(3);
    std::string why;
    int built = mapped ? 1 : 0;
    bool any = false;
    for(int v=0; v<nviews and built; v++){
        if (not fits[v]) continue;
        if (not stage_map(precision, v, 3, peers, maps[v], why, false)){ built = 0; break; }
        entry.has[v] = true;
        any = true;
    }
    if (built and any and cudaMalloc(&entry.maps, maps.size() * sizeof(scatter_map)) != cudaSuccess){ entry.maps = nullptr; cudaGetLastError(); built = 0; }
    if (built and any and cudaMemcpy(entry.maps, maps.data(), maps.size() * sizeof(scatter_map), cudaMemcpyHostToDevice) != cudaSuccess) built = 0;
    {
        std::vector<int> votes(n);
        if (ccomm->allgather(&built, votes.data(), sizeof(int)) != 0) built = 0;
        for(int v : votes) if (not v) built = 0;
    }
    if (not built or not any){
        if (entry.maps) cudaFree(entry.maps);
        return B200_ERR_UNSUPPORTED;
    }
    for(auto &u : P.user) if (u.ptr == ptr){ if (u.maps) cudaFree(u.maps); u = entry; return B200_SUCCESS; }
    P.user.push_back(entry);
    return B200_SUCCESS;
}

int transform3d::unregister_buffer(int precision, void *ptr){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    peer_state &P = peer[precision];
    for(size_t i=0; i<P.user.size(); i++){
        if (P.user[i].ptr != ptr) continue;
        cudaStreamSynchronize(cstream);          // a launch that reads the map may still be in flight
        if (P.user[i].maps) cudaFree(P.user[i].maps);
        P.user.erase(P.user.begin() + static_cast<long>(i));
        break;
    }
    return B200_SUCCESS;
}

bool transform3d::ensure_side_stream(){
    if (side_stream != nullptr) return true;
    if (cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking) != cudaSuccess){ side_stream = nullptr; cudaGetLastError(); return false; }
    if (cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming) != cudaSuccess or cudaEventCreateWithFlags(&join_event, cudaEventDisableTiming) != cudaSuccess){
        cudaGetLastError();
        return false;
    }
    return true;
}

int transform3d::prepare(int precision, int batch){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    int rc = ensure_executors(precision);
    if (rc) return rc;
    ensure_peer(precision, batch);
    return B200_SUCCESS;
}

void* transform3d::pair_counters(size_t count){
    if (count > counters_count){
        if (counters){ cudaStreamSynchronize(cstream); cudaFree(counters); counters = nullptr; counters_count = 0; }
        if (cudaMalloc(&counters, count * sizeof(unsigned)) != cudaSuccess){ cudaGetLastError(); return nullptr; }
        counters_count = count;
    }
    return counters;
}

void transform3d::mark(const char *name, long long local_bytes, long long sent_bytes){
    if (not timing) return;
    size_t const k = pending.size();
    if (marks.size() <= k){
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        marks.push_back(e);
    }
    stage_record r{};
    std::snprintf(r.name, sizeof(r.name), "%s", name);
    r.ms = 0; r.local_bytes = local_bytes; r.sent_bytes = sent_bytes;
    pending.push_back(r);
    cudaEventRecord(marks[k], cstream);
}

// call after the stream has been synchronised; entry i covers the time between mark i-1 and mark i
std::vector<transform3d::stage_record> transform3d::collect_stage_times(){
    std::vector<stage_record> out;
    for(size_t i=1; i<pending.size(); i++){
        float ms = 0;
        if (cudaEventElapsedTime(&ms, marks[i-1], marks[i]) != cudaSuccess) ms = -1;
        stage_record r = pending[i];
        r.ms = ms;
        out.push_back(r);
    }
    return out;
}

int transform3d::peer_fence(int precision){
    peer_state &P = peer[precision];
    P.epoch++;
    if (std::getenv("HEFFTE_B200_TRACE")) std::fprintf(stderr, "[b200 rank %d] fence %llu\n", me, P.epoch);
    ccomm->before_peer_barrier();
    int rc = b200_peer_barrier(ccomm->size(), me, P.remote_slots.data(), P.arena, P.epoch, cstream);
    ccomm->after_peer_barrier();
    return rc;
}

// One transform (or one fused spectral operator) with every reshape fused into the store of the kernel in front of it, for
// `batch` entries at once: every stage is ONE launch for all entries (grid.y) followed by ONE fence (reference: batch widens the
// messages of every reshape, src/heffte_reshape3d.cpp:379-384, 401-419).
// Data moves through three peer-mapped buffers that rotate: a stage that writes into other ranks' memory always targets the
// buffer after the last one used, on every rank alike, and a fence follows it.  When a rank writes buffer w, every rank has
// passed the fence of the previous remote stage, so it has finished all reads of what w held before (that was consumed two
// remote stages ago at the latest) -- no fence is needed at the start of a transform.
// A purely local transform in front of a fused one runs inside the same persistent kernel (fft_pair_kernel): its HBM traffic
// hides behind the NVLink-bound stores.
int transform3d::run_peer(int precision, int mode, int batch, const void *in, void *out, double scale, const void *multiplier){
    peer_state &P = peer[precision];
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    // Two-stream overlap of a local transform with the fused stage behind it: measured on 2 x B200 (profiles/r02_multi_2gpu) it
    // LOSES -- the fused kernels need every resident CTA to keep the link busy (thin grids: 400 instead of 690 GB/s), so the
    // local transform has no SM time to hide in.  Opt-in for experiments.
    static bool const allow_pair = (std::getenv("HEFFTE_B200_OVERLAP") != nullptr);
    static bool const allow_direct = (std::getenv("HEFFTE_B200_NO_DIRECT_OUTPUT") == nullptr);
    static bool const allow_registered = (std::getenv("HEFFTE_B200_NO_REGISTERED_OUTPUT") == nullptr);
    static int const thin_blocks = []{ const char *e = std::getenv("HEFFTE_B200_THIN_CTAS_PER_SM"); int const k = e ? std::atoi(e) : 2; return 148 * ((k > 0) ? k : 2); }();
    bool const shared_device = (std::strcmp(ccomm->kind(), "threads") == 0) and std::getenv("HEFFTE_B200_OVERLAP_ON_SHARED_DEVICE") == nullptr;
    static bool const trace = (std::getenv("HEFFTE_B200_TRACE") != nullptr);

    // the sequence of stages: (view, stage), see transform.h; a convolution replaces the last forward stage and the first stage of
    // the mirrored backward transform by the operator kernel, whose store is the reshape of that backward stage
    struct op { int dir, st; bool conv; };      // dir: the view the stage belongs to
    std::vector<op> ops;
    if (mode == mode_convolve){
        ops = {{view_forward, 0, false}, {view_forward, 1, false}, {view_forward, 2, false}, {view_mirror, 1, true}, {view_mirror, 2, false}, {view_mirror, 3, false}};
    }else{
        int const d = (mode == mode_backward) ? (lb_active ? view_backward : view_mirror) : view_forward;
        ops = {{d, 0, false}, {d, 1, false}, {d, 2, false}, {d, 3, false}};
    }
    auto executor_of = [](op const &o){ return o.st - 1; };           // index inside the view
    auto X = [&](op const &o){ return vexec(precision, o.dir, o.st - 1); };
    auto map_of = [&](op const &o, int w){ return static_cast<const char*>(P.maps) + sizeof(scatter_map) * static_cast<size_t>((o.dir * 4 + o.st) * 3 + w); };
    // element size on the input / output side of transform e of view `dir`
    auto bytes_of = [&](int e, int dir, bool output){
        if (tkind == kind_c2c) return cplx_bytes;
        if (tkind != kind_r2c) return real_bytes;
        if (real_id(dir, e) != 0) return cplx_bytes;
        return (output != (dir != view_forward)) ? cplx_bytes : real_bytes;   // r2c forward writes complex, c2r backward writes real
    };
    bool const last_is_backward = (ops.back().dir != view_forward);
    int const last_view = ops.back().dir;
    int const in_unit = (mode == mode_backward) ? (complex_data ? cplx_bytes : real_bytes)
                                                : ((tkind == kind_c2c) ? cplx_bytes : real_bytes);
    bool const real_out = (tkind == kind_r2c and last_is_backward) or not complex_data;
    int const out_unit = real_out ? real_bytes : cplx_bytes;
    long long const in_entry = static_cast<long long>((mode == mode_backward) ? outbox_count : inbox_count) * in_unit;
    long long const out_entry = static_cast<long long>(last_is_backward ? inbox_count : outbox_count) * out_unit;
    long long const arena_entry = static_cast<long long>(P.entry_bytes);

    // HEFFTE_B200_CHECK_REGISTERED=1 (debugging aid, one host allgather per call): the contract of register_buffer -- either every
    // rank passes the array it registered as the output of this call, or none does
    static bool const check_registered = (std::getenv("HEFFTE_B200_CHECK_REGISTERED") != nullptr);
    if (check_registered and not P.user.empty()){
        int mine = 0;
        for(auto const &u : P.user) if (u.ptr == out and u.has[last_view]) mine = 1;
        std::vector<int> all(static_cast<size_t>(ccomm->size()));
        if (ccomm->allgather(&mine, all.data(), sizeof(int)) != 0) return fail(B200_ERR_PEER, "allgather failed");
        for(int v : all) if (v != mine) return fail(B200_ERR_INVALID, "registered arrays: every rank must pass the array it registered as the output of the same call");
    }

    pending.clear();
    mark("start", 0, 0);
    int rc = B200_SUCCESS;
    const void *cur = in;
    long long cur_step = in_entry;
    int cur_buffer = -1;             // -1: caller memory
    bool landed_direct = false;
    int fences_done = 0;
    int const last_fft_op = static_cast<int>(ops.size()) - 1;

    for(size_t i=0; i<ops.size(); i++){
        op const &o = ops[i];
        bool const fused = P.fused[o.dir][o.st];
        if (o.st == 0){
            // ---- the first reshape of a transform: a scatter copy ------------------------------------------------------------
            if (not fused) continue;
            box3 const &box = vin(o.dir, 0)[me];
            int bytes = complex_data ? cplx_bytes : real_bytes;
            if (tkind == kind_r2c and o.dir == view_forward) bytes = real_bytes;
            int const w = P.take();
            if (not box.empty()){
                rc = b200_scatter_copy_batch(bytes, box.osize(0), box.osize(1), box.osize(2), box.osize(0), box.osize(0) * box.osize(1), cur, map_of(o, w),
                                             cstream, batch, cur_step, arena_entry, 0, 0);
                if (rc) return rc;
            }
            mark("reshape0 (scatter copy)", batch * (2 * stage_elems[o.dir][0] - sent_elems[o.dir][0]) * bytes, batch * sent_elems[o.dir][0] * bytes);
            rc = peer_fence(precision);
            if (rc) return rc;
            fences_done++;
            mark("fence", 0, 0);
            cur_buffer = w; cur = P.buffer(w); cur_step = arena_entry;
            continue;
        }
        int const e = executor_of(o);
        int const direction = (o.dir != view_forward) ? B200_BACKWARD : B200_FORWARD;
        b200_fft1d_plan const Xe = X(o);
        // the scaling rides on ONE stage of the plan -- a global choice: a rank whose box is empty in that stage must not scale
        // earlier, its data would be scaled again by the ranks that receive it.  Transforms: the last stage; operator: the product.
        double const stage_scale = (mode == mode_convolve) ? (o.conv ? scale : 1.0) : ((static_cast<int>(i) == last_fft_op) ? scale : 1.0);
        bool const type_changes = (tkind == kind_r2c and real_id(o.dir, e) == 0);          // r2c / c2r cannot run in place
        bool const is_last = (i + 1 == ops.size());
        bool later_fused = false;
        for(size_t t=i+1; t<ops.size(); t++) later_fused = later_fused or P.fused[ops[t].dir][ops[t].st];

        if (fused){
            int const w = P.take();                 // also when the stage ends in a registered array: the rotation stays in step
            long long local_shift = 0, local_step = 0;
            const char *map_here = map_of(o, w);
            bool into_registered = false;
            if (is_last and allow_direct){
                // The caller registered this array (register_buffer): every GPU stores its part of the result straight into it.
                // The other ranks write while I may still be reading `in` (an in-place call): allowed once a fence of THIS call
                // lies behind us, because `in` is only read by the first stage.
                if (allow_registered and batch == 1 and fences_done > 0){
                    for(auto const &u : P.user){
                        if (u.ptr == out and u.has[o.dir]){
                            map_here = static_cast<const char*>(u.maps) + sizeof(scatter_map) * static_cast<size_t>(o.dir);
                            into_registered = true;
                            break;
                        }
                    }
                }
                if (not into_registered){
                    // last stage: the part of my output that I produce myself goes straight into the caller's array (the cells of the
                    // map that stay in my own memory are re-based by the kernel); only what the other GPUs send lands in the arena
                    local_shift = static_cast<long long>(reinterpret_cast<intptr_t>(out)) - static_cast<long long>(reinterpret_cast<intptr_t>(P.buffer(w)));
                    local_step = out_entry - arena_entry;
                    landed_direct = true;
                }
            }
            if (trace) std::fprintf(stderr, "[b200 rank %d] stage (%d,%d) fused -> %s %d\n", me, o.dir, o.st, into_registered ? "registered array, skipped buffer" : "buffer", w);
            if (Xe){
                if (o.conv){
                    rc = b200_fft1d_execute_convolve(Xe, cur, nullptr, map_here, multiplier, stage_scale, cstream, batch, cur_step, 0, arena_entry, local_shift, local_step);
                    if (rc == B200_ERR_UNSUPPORTED){
                        // no operator kernel for this axis: transform in place, multiply, and let the backward kernel carry the reshape
                        void *here = const_cast<void*>(cur);
                        if (cur_buffer < 0) return fail(B200_ERR_UNSUPPORTED, "the fused spectral operator needs the plan's buffers");
                        rc = b200_fft1d_execute_batch(Xe, B200_FORWARD, cur, here, 1.0, cstream, batch, cur_step, cur_step);
                        for(int b=0; b<batch and rc == 0; b++)
                            rc = b200_pointwise_multiply(precision, vout(o.dir, e)[me].count(), static_cast<char*>(here) + b * cur_step, multiplier, stage_scale, cstream);
                        if (rc == 0) rc = b200_fft1d_execute_scatter_batch(Xe, B200_BACKWARD, cur, map_here, 1.0, cstream, batch, cur_step, arena_entry, local_shift, local_step);
                    }
                }else{
                    rc = b200_fft1d_execute_scatter_batch(Xe, direction, cur, map_here, stage_scale, cstream, batch, cur_step, arena_entry, local_shift, local_step);
                }
                if (rc) return rc;
            }
            {
                char label[40];
                int const axis = vdim(o.dir, e);      // the labels name the transformed dimension
                std::snprintf(label, sizeof(label), o.conv ? "fft%d * ifft%d + reshape%d (fused)" : "fft%d + reshape%d (fused)", axis, o.conv ? axis : o.st, o.st);
                long long const read_bytes = Xe ? static_cast<long long>(vout(o.dir, e)[me].count()) * bytes_of(e, o.dir, false) : 0;
                long long const wrote = stage_elems[o.dir][o.st] * bytes_of(e, o.dir, true), sent = sent_elems[o.dir][o.st] * bytes_of(e, o.dir, true);
                mark(label, batch * (read_bytes + wrote - sent), batch * sent);
            }
            rc = peer_fence(precision);
            if (rc) return rc;
            fences_done++;
            mark("fence", 0, 0);
            if (into_registered){ cur_buffer = -1; cur = out; cur_step = out_entry; }
            else{ cur_buffer = w; cur = P.buffer(w); cur_step = arena_entry; }
            continue;
        }

        // ---- a stage without data movement --------------------------------------------------------------------------------------
        void *dst;
        int dst_buffer;
        long long dst_step;
        // the caller's output can take the result once nothing moves any more -- except the complex intermediates of a
        // complex-to-real transform, which do not fit the real output array
        bool const fits_output = not (tkind == kind_r2c and last_is_backward and not is_last);
        if (not later_fused and fits_output){ dst = out; dst_buffer = -1; dst_step = out_entry; }
        else if (cur_buffer >= 0 and not type_changes){ dst = const_cast<void*>(cur); dst_buffer = cur_buffer; dst_step = cur_step; }
        else{ dst_buffer = P.take(); dst = P.buffer(dst_buffer); dst_step = arena_entry; }

        // followed by a fused stage of the same box: the two launches overlap on two streams, plane by plane -- this one, bound by
        // HBM, hides behind the remote stores of the next, which runs with a thin grid (tools/kbench_peer.cu).  Ranks that are host
        // threads sharing one GPU keep the launches apart: their thin kernels together could occupy every SM while they wait.
        bool paired = false;
        if (allow_pair and not shared_device and not o.conv and i + 1 < ops.size() and tkind == kind_c2c and Xe){
            op const &next = ops[i + 1];
            int const e2 = executor_of(next);
            if (P.fused[next.dir][next.st] and not next.conv and next.dir == o.dir and X(next) and b200_fft1d_overlappable(Xe, X(next)) and ensure_side_stream()){
                box3 const &box = vout(o.dir, e)[me];
                void *planes = pair_counters(static_cast<size_t>(box.osize(2)) * batch);
                if (planes != nullptr){
                    bool const next_last = (i + 2 == ops.size());
                    int const w = P.take();
                    long long local_shift = 0, local_step = 0;
                    if (next_last and allow_direct){
                        local_shift = static_cast<long long>(reinterpret_cast<intptr_t>(out)) - static_cast<long long>(reinterpret_cast<intptr_t>(P.buffer(w)));
                        local_step = out_entry - arena_entry;
                    }
                    double const next_scale = (mode != mode_convolve and static_cast<int>(i + 1) == last_fft_op) ? scale : 1.0;
                    int const nb = P.host_maps[static_cast<size_t>((next.dir * 4 + next.st) * 3 + w)].nb;
                    rc = b200_fft1d_execute_overlapped(Xe, X(next), direction, cur, dst, map_of(next, w), nb, next_scale, planes, cstream, side_stream,
                                                       fork_event, join_event, batch, cur_step, dst_step, arena_entry, local_shift, local_step, thin_blocks);
                    if (rc == B200_SUCCESS){
                        paired = true;
                        if (next_last and allow_direct) landed_direct = true;
                        if (trace) std::fprintf(stderr, "[b200 rank %d] stages (%d,%d) and (%d,%d) overlapped -> buffer %d\n", me, o.dir, o.st, next.dir, next.st, w);
                        char label[40];
                        std::snprintf(label, sizeof(label), "fft%d | fft%d + reshape%d (overlapped)", vdim(o.dir, e), vdim(next.dir, e2), next.st);
                        long long const count = box.count();
                        long long const wrote = stage_elems[next.dir][next.st] * cplx_bytes, sent = sent_elems[next.dir][next.st] * cplx_bytes;
                        mark(label, batch * (3 * count * cplx_bytes + wrote - sent), batch * sent);
                        rc = peer_fence(precision);
                        if (rc) return rc;
                        fences_done++;
                        mark("fence", 0, 0);
                        cur_buffer = w; cur = P.buffer(w); cur_step = arena_entry;
                        i++;                     // the fused stage is done
                    }else if (rc != B200_ERR_UNSUPPORTED) return rc;
                    else P.next_buffer = w;      // not taken after all: the rotation stays in step with the other ranks
                }
            }
        }
        if (paired) continue;

        if (Xe){
            if (o.conv){
                rc = b200_fft1d_execute_convolve(Xe, cur, dst, nullptr, multiplier, stage_scale, cstream, batch, cur_step, dst_step, 0, 0, 0);
                if (rc == B200_ERR_UNSUPPORTED){
                    rc = b200_fft1d_execute_batch(Xe, B200_FORWARD, cur, dst, 1.0, cstream, batch, cur_step, dst_step);
                    for(int b=0; b<batch and rc == 0; b++)
                        rc = b200_pointwise_multiply(precision, vout(o.dir, e)[me].count(), static_cast<char*>(dst) + b * dst_step, multiplier, stage_scale, cstream);
                    if (rc == 0) rc = b200_fft1d_execute_batch(Xe, B200_BACKWARD, dst, dst, 1.0, cstream, batch, dst_step, dst_step);
                }
            }else{
                rc = b200_fft1d_execute_batch(Xe, direction, cur, dst, stage_scale, cstream, batch, cur_step, dst_step);
            }
            if (rc) return rc;
        }
        {   // the mark is emitted on every rank, also with an empty box: all ranks report the same list of stages
            char label[40];
            std::snprintf(label, sizeof(label), o.conv ? "fft%d * ifft%d (local)" : "fft%d (local)", vdim(o.dir, e), vdim(o.dir, e));
            bool const real_stage = (tkind == kind_r2c and real_id(o.dir, e) == 0);
            // the real transform of an r2c plan: real box lp.out_shape[0], shortened complex box lp.in_shape[1]
            long long const count = Xe ? (real_stage ? lp.out_shape[0][me].count() : vout(o.dir, e)[me].count()) : 0;
            long long const half = Xe ? lp.in_shape[1][me].count() : 0;
            long long const out_count = real_stage ? ((o.dir != view_forward) ? count : half) : count;
            long long const in_count = (real_stage and o.dir != view_forward) ? half : count;
            mark(label, batch * (in_count * bytes_of(e, o.dir, false) + out_count * bytes_of(e, o.dir, true)), 0);
        }
        cur = dst; cur_buffer = dst_buffer; cur_step = dst_step;     // also without a transform (empty box): every rank follows the same buffers
    }

    if (cur != out and landed_direct){
        // what the other GPUs sent sits in the arena at its final position inside my box: move those sub-boxes only, all of
        // them and all batch entries in one launch
        shape const &from = vin(last_view, 3);
        box3 const &mine = vout(last_view, 3)[me];
        std::vector<long long> offsets, nf, nm, ns;
        long long moved = 0;
        for(int r=0; r<ccomm->size() and not mine.empty(); r++){
            if (r == me) continue;
            box3 const piece = mine.overlap(from[r]);
            if (piece.empty()) continue;
            offsets.push_back(mine.offset_of(piece.low));
            nf.push_back(piece.osize(0)); nm.push_back(piece.osize(1)); ns.push_back(piece.osize(2));
            moved += piece.count();
        }
        if (not offsets.empty()){
            rc = b200_copy_subboxes(out_unit, static_cast<int>(offsets.size()), offsets.data(), nf.data(), nm.data(), ns.data(),
                                    mine.osize(0), mine.osize(0) * mine.osize(1), cur, out, cstream, batch, cur_step, out_entry);
            if (rc) return rc;
        }
        mark("received sub-boxes to the caller's array", batch * 2 * moved * out_unit, 0);
    }else if (cur != out){
        idx const count = last_is_backward ? inbox_count : outbox_count;
        size_t const bytes = static_cast<size_t>(count) * out_unit;
        for(int b=0; b<batch; b++)
            if (count > 0 and cudaMemcpyAsync(static_cast<char*>(out) + b * out_entry, static_cast<const char*>(cur) + b * cur_step, bytes, cudaMemcpyDeviceToDevice, cstream) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed");
        mark("copy to the caller's array", batch * 2 * static_cast<long long>(bytes), 0);
    }
    return B200_SUCCESS;
}

// every stage of the plan is local to this rank (one rank, or boxes that never move): the batched launches need no arena
bool transform3d::all_local() const {
    for(int i=0; i<4; i++) if (fwd[i] or bwd[i]) return false;
    return true;
}

// Transforms without any data movement: three batched launches (grid.y = entries), in place where the types allow.  The fused
// spectral operator runs the last forward and the first backward transform in one kernel; single precision pairs the two
// transforms of the fast axes through the L2 cache (fft_pair_kernel, tools/kbench_pair.cu).
int transform3d::run_local(int precision, int mode, int batch, const void *in, void *out, double scale, const void *multiplier){
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    b200_fft1d_plan const *X = exec[precision];
    // pairing two local transforms through the L2 cache pays for small planes only (tools/kbench_pair.cu: 256^3 fp32 -19 %,
    // 512^3 fp64 +22 %) and not at all once the L2 is cold between transforms: opt-in
    static bool const allow_pair = (std::getenv("HEFFTE_B200_PAIR_LOCAL") != nullptr);
    bool const backward = (mode == mode_backward);
    long long const spatial_entry = static_cast<long long>(inbox_count) * ((tkind == kind_c2c) ? cplx_bytes : real_bytes);
    long long const spectral_entry = static_cast<long long>(outbox_count) * (complex_data ? cplx_bytes : real_bytes);
    long long const in_entry = backward ? spectral_entry : spatial_entry;
    long long const out_entry = (mode == mode_forward) ? spectral_entry : spatial_entry;
    int order[6], dirs[6], count = 0;
    if (mode == mode_forward){ for(int e=0; e<3; e++){ order[count] = e; dirs[count++] = B200_FORWARD; } }
    else if (mode == mode_backward){ for(int e=2; e>=0; e--){ order[count] = e; dirs[count++] = B200_BACKWARD; } }
    else{ order[0] = 0; order[1] = 1; order[2] = 2; order[3] = 1; order[4] = 0; dirs[0] = dirs[1] = B200_FORWARD; dirs[2] = 2; dirs[3] = dirs[4] = B200_BACKWARD; count = 5; }
    const void *cur = in;
    long long cur_step = in_entry;
    int rc = B200_SUCCESS;
    for(int i=0; i<count; i++){
        int const e = order[i];
        if (not X[e]) continue;
        bool const last = (i == count - 1);
        double const s = (mode == mode_convolve) ? ((dirs[i] == 2) ? scale : 1.0) : (last ? scale : 1.0);
        // r2c: the first transform changes the element type, so it writes straight into the output array; everything else in place
        void *dst = out;
        long long dst_step = out_entry;
        if (tkind == kind_r2c and backward and e != 0){
            // complex intermediates of a complex-to-real transform do not fit the real output: they stay in the plan's workspace
            void *scratch = ensure_workspace(precision, batch);
            if (scratch == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
            dst = scratch; dst_step = spectral_entry;
        }
        if (dirs[i] == 2){
            rc = b200_fft1d_execute_convolve(X[e], cur, dst, nullptr, multiplier, s, cstream, batch, cur_step, dst_step, 0, 0, 0);
            if (rc == B200_ERR_UNSUPPORTED){
                rc = b200_fft1d_execute_batch(X[e], B200_FORWARD, cur, dst, 1.0, cstream, batch, cur_step, dst_step);
                for(int b=0; b<batch and rc == 0; b++)
                    rc = b200_pointwise_multiply(precision, lp.out_shape[e][me].count(), static_cast<char*>(dst) + b * dst_step, multiplier, s, cstream);
                if (rc == 0) rc = b200_fft1d_execute_batch(X[e], B200_BACKWARD, dst, dst, 1.0, cstream, batch, dst_step, dst_step);
            }
        }else if (allow_pair and precision == B200_PREC_FLOAT and i + 1 < count and dirs[i + 1] == dirs[i] and X[order[i + 1]] and
                  b200_fft1d_pairable(X[e], X[order[i + 1]]) and pair_counters(static_cast<size_t>(lp.out_shape[e][me].osize(2)) * batch) != nullptr){
            bool const next_last = (i + 1 == count - 1);
            double const s2 = (mode == mode_convolve) ? 1.0 : (next_last ? scale : 1.0);
            rc = b200_fft1d_execute_pair(X[e], X[order[i + 1]], dirs[i], cur, dst, nullptr, s2, counters, 0, cstream, batch, cur_step, dst_step, 0, 0, 0);
            if (rc == B200_SUCCESS) i++;
            else if (rc == B200_ERR_UNSUPPORTED) rc = b200_fft1d_execute_batch(X[e], dirs[i], cur, dst, s, cstream, batch, cur_step, dst_step);
        }else{
            rc = b200_fft1d_execute_batch(X[e], dirs[i], cur, dst, s, cstream, batch, cur_step, dst_step);
        }
        if (rc) return rc;
        cur = dst; cur_step = dst_step;
    }
    if (cur != out){
        // nothing ran (empty box) or the data stopped in the scratch array: the output is what the input was
        idx const n = (mode == mode_forward) ? outbox_count : inbox_count;
        if (n > 0 and cur == in and in != out){
            for(int b=0; b<batch; b++)
                if (cudaMemcpyAsync(static_cast<char*>(out) + b * out_entry, static_cast<const char*>(in) + b * in_entry,
                                    static_cast<size_t>(std::min(in_entry, out_entry)), cudaMemcpyDeviceToDevice, cstream) != cudaSuccess)
                    return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed");
        }
    }
    return B200_SUCCESS;
}

// HEFFTE_B200_L2_SLAB_MB=<megabytes> (default: off) pairs two consecutive local transforms of the same box slab by slab:
// the first writes a slab of that many megabytes, the second reads it back from the L2 cache instead of HBM (126 MB on
// B200) and overwrites it in place, so the pair costs one read and one write of HBM instead of two of each.
// Returns the number of planes (index of the slowest axis) per slab, 0 when the pairing does not apply.
idx transform3d::l2_slab_planes(int first, int second, int elem_bytes) const {
    const char *setting = std::getenv("HEFFTE_B200_L2_SLAB_MB");
    if (setting == nullptr) return 0;
    double const megabytes = std::atof(setting);
    if (not (megabytes > 0)) return 0;
    box3 const &a = lp.out_shape[first][me], &b = lp.out_shape[second][me];
    if (a.empty() or not a.same_extent(b) or not a.same_order(b)) return 0;
    int const slow = a.order[2];
    if (lp.fft_direction[first] == slow or lp.fft_direction[second] == slow) return 0;
    double const plane_bytes = static_cast<double>(a.osize(0)) * static_cast<double>(a.osize(1)) * elem_bytes;
    idx planes = static_cast<idx>(megabytes * 1e6 / plane_bytes);
    planes = std::max<idx>(planes, 1);
    return (planes >= a.osize(2)) ? 0 : planes;       // one slab = the whole box: nothing to gain
}

double transform3d::scale_factor(int scaling) const {
    if (scaling == 0) return 1.0;
    return (scaling == 2) ? std::sqrt(base_scale) : base_scale;
}

int transform3d::ensure_executors(int precision){
    if (exec_ready[precision]) return B200_SUCCESS;
    for(int i=0; i<3; i++){
        box3 const &box = lp.out_shape[i][me];
        if (box.empty()) continue;
        int const dim = lp.fft_direction[i];
        b200_fft1d_desc d{};
        d.precision = precision;
        d.n = box.size(dim);
        line_layout(box, dim, d.in, d.count_a, d.count_b);
        d.out = d.in;
        if (tkind == kind_r2c){
            if (i == 0){
                d.kind = B200_R2C;
                long long ca, cb;
                line_layout(lp.in_shape[1][me], dim, d.out, ca, cb);   // the shortened complex box
            }else d.kind = B200_C2C;
        }else d.kind = static_cast<int>(tkind);
        int rc = b200_fft1d_create(&d, &exec[precision][i]);
        if (rc) return rc;
    }
    for(int i=0; i<3 and lb_active; i++){
        box3 const &box = lb.out_shape[i][me];
        if (box.empty()) continue;
        int const dim = lb.fft_direction[i];
        b200_fft1d_desc d{};
        d.precision = precision;
        d.n = box.size(dim);
        line_layout(box, dim, d.in, d.count_a, d.count_b);
        d.out = d.in;
        d.kind = static_cast<int>(tkind);
        int rc = b200_fft1d_create(&d, &bexec[precision][i]);
        if (rc) return rc;
    }
    exec_ready[precision] = true;
    return B200_SUCCESS;
}

void* transform3d::ensure_workspace(int precision, int batch){
    size_t const unit = (precision == B200_PREC_FLOAT ? 4 : 8) * ((tkind == kind_c2c or tkind == kind_r2c) ? 2 : 1);
    size_t const need = static_cast<size_t>(std::max(workspace_count, exec_workspace_count)) * unit * static_cast<size_t>(std::max(batch, 1)) + 64;
    if (need > own_workspace_bytes){
        if (own_workspace){ cudaStreamSynchronize(cstream); cudaFree(own_workspace); own_workspace = nullptr; own_workspace_bytes = 0; }
        if (cudaMalloc(&own_workspace, need) != cudaSuccess) return nullptr;
        own_workspace_bytes = need;
    }
    return own_workspace;
}

int transform3d::forward(int precision, int batch, const void *in, void *out, void *workspace, int scaling){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    if (ccomm->size() > 1 and b200_peer_timed_out()) return fail(B200_ERR_PEER, "a peer GPU did not reach a barrier within HEFFTE_B200_BARRIER_TIMEOUT_S");
    batch = std::max(batch, 1);
    int rc = ensure_executors(precision);
    if (rc) return rc;
    if (all_local()) return run_local(precision, mode_forward, batch, in, out, scale_factor(scaling), nullptr);
    if (ensure_peer(precision, batch)) return run_peer(precision, mode_forward, batch, in, out, scale_factor(scaling), nullptr);
    // the exchange path: pack -> exchange() -> unpack, entry by entry
    if (workspace == nullptr or exec_workspace_count > workspace_count){
        workspace = ensure_workspace(precision, 1);
        if (workspace == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
    }
    size_t const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    size_t const in_unit = (tkind == kind_c2c) ? 2 * real_bytes : real_bytes;
    size_t const out_unit = (tkind == kind_c2c or tkind == kind_r2c) ? 2 * real_bytes : real_bytes;
    for(int b=0; b<batch; b++){
        const char *src = static_cast<const char*>(in) + b * inbox_count * in_unit;
        char *dst = static_cast<char*>(out) + b * outbox_count * out_unit;
        rc = run(precision, false, src, dst, workspace, scale_factor(scaling));
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

int transform3d::backward(int precision, int batch, const void *in, void *out, void *workspace, int scaling){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    if (ccomm->size() > 1 and b200_peer_timed_out()) return fail(B200_ERR_PEER, "a peer GPU did not reach a barrier within HEFFTE_B200_BARRIER_TIMEOUT_S");
    batch = std::max(batch, 1);
    int rc = ensure_executors(precision);
    if (rc) return rc;
    if (all_local()) return run_local(precision, mode_backward, batch, in, out, scale_factor(scaling), nullptr);
    if (ensure_peer(precision, batch)) return run_peer(precision, mode_backward, batch, in, out, scale_factor(scaling), nullptr);
    if (workspace == nullptr or exec_workspace_count > workspace_count){
        workspace = ensure_workspace(precision, 1);
        if (workspace == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
    }
    size_t const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    size_t const in_unit = (tkind == kind_c2c or tkind == kind_r2c) ? 2 * real_bytes : real_bytes;
    size_t const out_unit = (tkind == kind_c2c) ? 2 * real_bytes : real_bytes;
    for(int b=0; b<batch; b++){
        const char *src = static_cast<const char*>(in) + b * outbox_count * in_unit;
        char *dst = static_cast<char*>(out) + b * inbox_count * out_unit;
        rc = run(precision, true, src, dst, workspace, scale_factor(scaling));
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

// forward(in) with `scaling`, times itself or times `multiplier` (an array over convolve_box()), backward -- complex plans whose
// inbox and outbox hold the same number of entries (the result comes back in the layout of the input)
int transform3d::convolve(int precision, const void *in, void *out, void *workspace, const void *multiplier, int scaling){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    if (tkind != kind_c2c) return fail(B200_ERR_UNSUPPORTED, "the spectral operator is defined for complex-to-complex plans");
    if (ccomm->size() > 1 and b200_peer_timed_out()) return fail(B200_ERR_PEER, "a peer GPU did not reach a barrier within HEFFTE_B200_BARRIER_TIMEOUT_S");
    int rc = ensure_executors(precision);
    if (rc) return rc;
    if (all_local()) return run_local(precision, mode_convolve, 1, in, out, scale_factor(scaling), multiplier);
    if (ensure_peer(precision, 1)) return run_peer(precision, mode_convolve, 1, in, out, scale_factor(scaling), multiplier);
    // the exchange path has no fused form: forward into the plan's scratch, product, backward.  The multiplier is laid out over
    // convolve_box(), which the unfused forward transform does not produce, so only the self-product is available here.
    if (multiplier != nullptr) return fail(B200_ERR_UNSUPPORTED, "a caller multiplier needs the peer-memory or the single-rank path");
    size_t const cplx_bytes = (precision == B200_PREC_FLOAT) ? 8 : 16;
    if (workspace == nullptr or exec_workspace_count > workspace_count){
        workspace = ensure_workspace(precision, 1);
        if (workspace == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
    }
    void *spectrum = nullptr;
    if (cudaMalloc(&spectrum, std::max<size_t>(static_cast<size_t>(outbox_count), 1) * cplx_bytes) != cudaSuccess) return fail(B200_ERR_CUDA, "cannot allocate the spectrum");
    rc = run(precision, false, in, spectrum, workspace, 1.0);
    if (rc == 0) rc = b200_pointwise_multiply(precision, outbox_count, spectrum, nullptr, scale_factor(scaling), cstream);
    if (rc == 0) rc = run(precision, true, spectrum, out, workspace, 1.0);
    cudaStreamSynchronize(cstream);
    cudaFree(spectrum);
    return rc;
}

int transform3d::run(int precision, bool is_backward, const void *in, void *out, void *workspace, double scale){
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    std::unique_ptr<reshape_op> const *R = is_backward ? bwd : fwd;
    int const E[3] = {is_backward ? 2 : 0, 1, is_backward ? 0 : 2};   // executor used after reshape s
    int const direction = is_backward ? B200_BACKWARD : B200_FORWARD;
    b200_fft1d_plan const *X = exec[precision];

    // the scaling rides on the last transform stage, on every rank (see run_peer)
    auto stage_scale = [&](int s){ return (s == 2) ? scale : 1.0; };

    // ---- complex-to-complex and real-to-real: one element type from end to end ---------------------------------
    if (tkind != kind_r2c){
        int const elem = (tkind == kind_c2c) ? cplx_bytes : real_bytes;
        void *temp = advance(workspace, comm_count, elem);
        int total_reshapes = 0, done_reshapes = 0;
        for(int s=0; s<4; s++) if (R[s]) total_reshapes++;
        const void *cur = in;
        bool writable = (in == out);
        for(int s=0; s<4; s++){
            if (R[s]){
                done_reshapes++;
                void *dst = (done_reshapes == total_reshapes) ? out : temp;
                int rc = R[s]->apply(elem, cur, dst, workspace, cstream);
                if (rc) return rc;
                cur = dst; writable = true;
            }
            if (s < 3 and X[E[s]]){
                void *dst = writable ? const_cast<void*>(cur) : ((done_reshapes < total_reshapes) ? temp : out);
                // two transforms of the same box with nothing between them, neither along the slowest axis: run them slab by slab
                // so that the second one finds its input in the L2 cache (opt-in, see l2_slab_planes())
                idx const planes = (s < 2 and not R[s+1] and X[E[s+1]]) ? l2_slab_planes(E[s], E[s+1], elem) : 0;
                if (planes > 0){
                    idx const count_b = lp.out_shape[E[s]][me].osize(2);
                    for(idx b0 = 0; b0 < count_b; b0 += planes){
                        idx const nb = std::min(planes, count_b - b0);
                        int rc = b200_fft1d_execute_range(X[E[s]], direction, cur, dst, stage_scale(s), cstream, b0, nb);
                        if (rc) return rc;
                        rc = b200_fft1d_execute_range(X[E[s+1]], direction, dst, dst, stage_scale(s+1), cstream, b0, nb);
                        if (rc) return rc;
                    }
                    cur = dst; writable = true;
                    s++;            // the second transform of the pair is done
                    continue;
                }
                int rc = b200_fft1d_execute(X[E[s]], direction, cur, dst, stage_scale(s), cstream);
                if (rc) return rc;
                cur = dst; writable = true;
            }
        }
        if (cur != out){
            idx const count = is_backward ? inbox_count : outbox_count;
            if (count > 0 and cudaMemcpyAsync(out, cur, static_cast<size_t>(count) * elem, cudaMemcpyDeviceToDevice, cstream) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed");
        }
        return B200_SUCCESS;
    }

    // ---- real-to-complex forward ----------------------------------------------------------------------------------
    void *temp = advance(workspace, comm_count, cplx_bytes);
    if (not is_backward){
        const void *real_in = in;
        if (R[0]){
            void *staged = workspace;
            void *scratch = advance(workspace, temp_count, cplx_bytes);
            int rc = R[0]->apply(real_bytes, in, staged, scratch, cstream);
            if (rc) return rc;
            real_in = staged;
        }
        int remaining = 0;
        for(int s=1; s<4; s++) if (R[s]) remaining++;
        void *cur = (remaining > 0) ? temp : out;
        if (X[0]){
            int rc = b200_fft1d_execute(X[0], B200_FORWARD, real_in, cur, stage_scale(0), cstream);
            if (rc) return rc;
        }
        for(int s=1; s<4; s++){
            if (R[s]){
                remaining--;
                void *dst = (remaining == 0) ? out : temp;
                int rc = R[s]->apply(cplx_bytes, cur, dst, workspace, cstream);
                if (rc) return rc;
                cur = dst;
            }
            if (s < 3 and X[E[s]]){
                int rc = b200_fft1d_execute(X[E[s]], B200_FORWARD, cur, cur, stage_scale(s), cstream);
                if (rc) return rc;
            }
        }
        return B200_SUCCESS;
    }

    // ---- complex-to-real backward -------------------------------------------------------------------------------------
    {
        if (R[0]){
            int rc = R[0]->apply(cplx_bytes, in, temp, workspace, cstream);
            if (rc) return rc;
        }else if (outbox_count > 0){
            if (cudaMemcpyAsync(temp, in, static_cast<size_t>(outbox_count) * cplx_bytes, cudaMemcpyDeviceToDevice, cstream) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed");
        }
        for(int s=0; s<2; s++){
            if (X[E[s]]){
                int rc = b200_fft1d_execute(X[E[s]], B200_BACKWARD, temp, temp, stage_scale(s), cstream);
                if (rc) return rc;
            }
            if (R[s+1]){
                int rc = R[s+1]->apply(cplx_bytes, temp, temp, workspace, cstream);
                if (rc) return rc;
            }
        }
        if (R[3]){
            void *real_buffer = workspace;
            idx const real_count = lp.out_shape[0][me].count();
            if (X[0]){
                int rc = b200_fft1d_execute(X[0], B200_BACKWARD, temp, real_buffer, stage_scale(2), cstream);
                if (rc) return rc;
            }
            void *scratch = advance(workspace, (real_count + 1) / 2, cplx_bytes);
            return R[3]->apply(real_bytes, real_buffer, out, scratch, cstream);
        }
        if (X[0]) return b200_fft1d_execute(X[0], B200_BACKWARD, temp, out, stage_scale(2), cstream);
        return B200_SUCCESS;
    }
}

} // namespace b200
