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

    {   // the executed plan depends on per-process switches (HEFFTE_B200_REFERENCE_PLAN, HEFFTE_B200_DECOMPOSITION): every rank
        // must have arrived at the same boxes, or the fused stores would land at wrong addresses
        unsigned long long h = 1469598103934665603ULL;
        auto mix = [&](long long v){ h ^= static_cast<unsigned long long>(v); h *= 1099511628211ULL; };
        for(int s=0; s<4; s++)
            for(shape const *sh : {&lp.in_shape[s], &lp.out_shape[s]})
                for(box3 const &b : *sh) for(int d=0; d<3; d++){ mix(b.low[d]); mix(b.high[d]); mix(b.order[d]); }
        for(int d=0; d<3; d++) mix(lp.fft_direction[d]);
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

transform3d::~transform3d(){
    for(int p=0; p<2; p++){
        peer_state &P = peer[p];
        if (P.arena){
            cudaStreamSynchronize(cstream);
            if (not P.arenas.empty()) ccomm->unmap_peers(P.arenas);
            cudaFree(P.arena);
        }
        if (P.maps) cudaFree(P.maps);
    }
    for(int p=0; p<2; p++) for(int i=0; i<3; i++) if (exec[p][i]) b200_fft1d_destroy(exec[p][i]);
    if (own_workspace) cudaFree(own_workspace);
    for(cudaEvent_t e : marks) cudaEventDestroy(e);
}

// ------------------------------------------------------------------------------------------------------------
// peer-memory mode
// ------------------------------------------------------------------------------------------------------------
// Collective over the ranks of the plan (first transform of each precision): allocate and peer-map the arena, build the
// scatter maps of every stage.  Any rank failing any step makes every rank fall back to the exchange() path.
bool transform3d::ensure_peer(int precision){
    peer_state &P = peer[precision];
    if (P.tried) return P.active;
    P.tried = true;
    int const n = ccomm->size();
    if (n < 2 or n > 64) return false;
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    bool any = false;
    for(int s=0; s<4; s++){
        P.fused[0][s] = shapes_differ(lp.in_shape[s], lp.out_shape[s]);
        P.fused[1][s] = shapes_differ(lp.out_shape[3-s], lp.in_shape[3-s]);
        any = any or P.fused[0][s];
    }
    if (not any) return false;

    // a buffer holds the largest box this rank ever owns, in the widest element type of the plan
    // (the same size on every rank: buffer 1 of a peer sits at a known offset inside its arena)
    idx largest = 1;
    for(int s=0; s<4; s++)
        for(int r=0; r<n; r++) largest = std::max(largest, std::max(lp.in_shape[s][r].count(), lp.out_shape[s][r].count()));
    P.buffer_bytes = ((static_cast<size_t>(largest) * (complex_data ? cplx_bytes : real_bytes) + 255) / 256) * 256;
    size_t const arena_bytes = 4096 + 2 * P.buffer_bytes;
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

    // scatter maps: ((direction * 4 + stage) * 2 + buffer)
    std::vector<scatter_map> maps(16);
    std::vector<int> owners(16 * scatter_max_cells, -1);
    std::string why;
    int built = 1;
    for(int dir=0; dir<2 and built; dir++){
        for(int st=0; st<4 and built; st++){
            if (not P.fused[dir][st]) continue;
            shape const &dest = (dir == 0) ? lp.out_shape[st] : lp.in_shape[3-st];
            // the box this rank writes in that stage, the axis of the transform in front of the reshape, the element size
            box3 written;
            int k_pos = 0, bytes = complex_data ? cplx_bytes : real_bytes;
            if (st == 0){
                written = (dir == 0) ? lp.in_shape[0][me] : lp.out_shape[3][me];
                if (tkind == kind_r2c and dir == 0) bytes = real_bytes;
            }else{
                int const e = (dir == 0) ? st - 1 : 3 - st;                 // executor in front of this reshape
                written = (dir == 0) ? lp.in_shape[st][me] : lp.out_shape[3-st][me];
                if (not written.empty()) k_pos = written.position_of(lp.fft_direction[e]);
                if (tkind == kind_r2c and dir == 1 and e == 0) bytes = real_bytes;   // c2r output
            }
            stage_elems[dir][st] = written.count();
            sent_elems[dir][st] = 0;
            for(int r=0; r<n; r++) if (r != me) sent_elems[dir][st] += written.overlap(dest[r]).count();
            for(int w=0; w<2; w++){
                std::vector<void*> bases(n);
                for(int r=0; r<n; r++) bases[r] = static_cast<char*>(arenas[r]) + 4096 + static_cast<size_t>(w) * P.buffer_bytes;
                if (not build_scatter_map(written, k_pos, dest, bases, bytes, maps[(dir * 4 + st) * 2 + w], why,
                                          owners.data() + static_cast<size_t>((dir * 4 + st) * 2 + w) * scatter_max_cells)){ built = 0; break; }
            }
        }
    }
    if (built and cudaMalloc(&P.maps, (maps.size() + 2) * sizeof(scatter_map)) != cudaSuccess){ P.maps = nullptr; cudaGetLastError(); built = 0; }
    if (built and cudaMemcpy(P.maps, maps.data(), maps.size() * sizeof(scatter_map), cudaMemcpyHostToDevice) != cudaSuccess) built = 0;
    {
        std::vector<int> votes(n);
        if (ccomm->allgather(&built, votes.data(), sizeof(int)) != 0) built = 0;
        for(int v : votes) if (not v) built = 0;
    }
    if (not built){
        ccomm->unmap_peers(P.arenas);
        P.arenas.clear();
        cudaFree(P.arena); P.arena = nullptr;
        if (P.maps){ cudaFree(P.maps); P.maps = nullptr; }
        return false;
    }
    P.host_maps = maps;
    P.owners = owners;
    P.active = true;
    return true;
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
    int rc = b200_peer_barrier(ccomm->size(), me, P.remote_slots.data(), P.arena, P.epoch, cstream);
    ccomm->after_peer_barrier();
    return rc;
}

// One transform with every reshape fused into the store of the kernel in front of it.  Data alternates between the two
// peer-mapped buffers; a fence follows every stage that writes into other ranks' memory.
int transform3d::run_peer(int precision, bool is_backward, const void *in, void *out, double scale){
    peer_state &P = peer[precision];
    int const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    int const cplx_bytes = 2 * real_bytes;
    bool const complex_data = (tkind == kind_c2c or tkind == kind_r2c);
    int const dir = is_backward ? 1 : 0;
    int const direction = is_backward ? B200_BACKWARD : B200_FORWARD;
    b200_fft1d_plan const *X = exec[precision];
    auto map_of = [&](int st, int w){ return static_cast<const char*>(P.maps) + sizeof(scatter_map) * static_cast<size_t>((dir * 4 + st) * 2 + w); };

    // the scaling rides on the LAST transform stage of the plan -- a global choice: a rank whose box is empty in that stage
    // must not scale earlier, its data would be scaled again by the ranks that receive it
    int const last_fft = 3;

    pending.clear();
    mark("start", 0, 0);
    // every peer has finished reading its buffers of the previous transform before anybody writes into them again
    int rc = peer_fence(precision);
    if (rc) return rc;
    mark("fence", 0, 0);
    unsigned touched = 0;            // buffers read or written locally since the last fence
    auto bytes_of = [&](int e, bool output){     // element size on the input / output side of executor e in this direction
        if (tkind == kind_c2c) return cplx_bytes;
        if (tkind != kind_r2c) return real_bytes;
        if (e != 0) return cplx_bytes;
        return (output != is_backward) ? cplx_bytes : real_bytes;   // r2c forward writes complex, c2r backward writes real
    };

    // the last reshape can deliver my own part straight into the caller's array (not for the real output of a c2r transform
    // that still has complex stages in the arena: there the arena element type differs only before stage 3, which is fine)
    bool const direct_local = P.fused[dir][3] and std::getenv("HEFFTE_B200_NO_DIRECT_OUTPUT") == nullptr;
    bool landed_direct = false;
    const void *cur = in;
    int cur_buffer = -1;             // -1: caller memory
    if (P.fused[dir][0]){
        box3 const &box = is_backward ? lp.out_shape[3][me] : lp.in_shape[0][me];
        int bytes = complex_data ? cplx_bytes : real_bytes;
        if (tkind == kind_r2c and not is_backward) bytes = real_bytes;
        if (not box.empty()){
            rc = b200_scatter_copy(bytes, box.osize(0), box.osize(1), box.osize(2), box.osize(0), box.osize(0) * box.osize(1), cur, map_of(0, 0), cstream);
            if (rc) return rc;
        }
        mark("reshape0 (scatter copy)", (2 * stage_elems[dir][0] - sent_elems[dir][0]) * bytes, sent_elems[dir][0] * bytes);
        rc = peer_fence(precision);
        if (rc) return rc;
        mark("fence", 0, 0);
        cur_buffer = 0; cur = P.buffer(0);
    }
    for(int st=1; st<4; st++){
        int const e = is_backward ? 3 - st : st - 1;
        double const stage_scale = (st == last_fft) ? scale : 1.0;
        bool const type_changes = (tkind == kind_r2c and e == 0);          // r2c / c2r cannot run in place
        if (P.fused[dir][st]){
            int const w = (cur_buffer < 0) ? 0 : (cur_buffer ^ 1);
            if (touched & (1u << w)){ rc = peer_fence(precision); if (rc) return rc; touched = 0; }
            const void *stage_map = map_of(st, w);
            if (st == 3 and direct_local){
                // last stage: the part of my output that I produce myself goes straight into the caller's array (the cells of the
                // map that point into my own arena are re-based); only what the other GPUs send lands in the arena
                if (P.patched_out[dir] != out or P.patched_buffer[dir] != w){
                    scatter_map patched = P.host_maps[(dir * 4 + 3) * 2 + w];
                    long long const arena_base = static_cast<long long>(reinterpret_cast<intptr_t>(P.buffer(w)));
                    const int *owner = P.owners.data() + static_cast<size_t>((dir * 4 + 3) * 2 + w) * scatter_max_cells;
                    for(int c=0; c<patched.ncells; c++)
                        if (owner[c] == me) patched.cell[c].base += static_cast<long long>(reinterpret_cast<intptr_t>(out)) - arena_base;
                    char *slot = static_cast<char*>(P.maps) + sizeof(scatter_map) * static_cast<size_t>(16 + dir);
                    if (cudaMemcpyAsync(slot, &patched, sizeof(scatter_map), cudaMemcpyHostToDevice, cstream) != cudaSuccess)
                        return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed (scatter map)");
                    if (cudaStreamSynchronize(cstream) != cudaSuccess) return fail(B200_ERR_CUDA, "stream synchronisation failed");   // `patched` is a local
                    P.patched_out[dir] = out; P.patched_buffer[dir] = w;
                }
                stage_map = static_cast<const char*>(P.maps) + sizeof(scatter_map) * static_cast<size_t>(16 + dir);
                landed_direct = true;
            }
            if (X[e]){
                rc = b200_fft1d_execute_scatter(X[e], direction, cur, stage_map, stage_scale, cstream);
                if (rc) return rc;
            }
            {
                char label[40];
                std::snprintf(label, sizeof(label), "fft%d + reshape%d (fused)", e, is_backward ? st : st);
                long long const read_bytes = X[e] ? static_cast<long long>(is_backward ? lp.out_shape[e][me].count() : lp.out_shape[e][me].count()) * bytes_of(e, false) : 0;
                long long const wrote = stage_elems[dir][st] * bytes_of(e, true), sent = sent_elems[dir][st] * bytes_of(e, true);
                mark(label, read_bytes + wrote - sent, sent);
            }
            rc = peer_fence(precision);
            if (rc) return rc;
            mark("fence", 0, 0);
            touched = 0;
            cur_buffer = w; cur = P.buffer(w);
        }else{
            bool later_fused = false;
            for(int t=st+1; t<4; t++) later_fused = later_fused or P.fused[dir][t];
            void *dst;
            int dst_buffer;
            // the caller's output can take the result once nothing moves any more -- except the complex intermediates of a
            // complex-to-real transform, which do not fit the real output array
            bool const fits_output = not (tkind == kind_r2c and is_backward and st < 3);
            if (not later_fused and fits_output){ dst = out; dst_buffer = -1; }
            else if (cur_buffer >= 0 and not type_changes){ dst = const_cast<void*>(cur); dst_buffer = cur_buffer; }
            else{ dst_buffer = (cur_buffer < 0) ? 0 : (cur_buffer ^ 1); dst = P.buffer(dst_buffer); }
            if (cur_buffer >= 0) touched |= 1u << cur_buffer;
            if (dst_buffer >= 0) touched |= 1u << dst_buffer;
            if (X[e]){
                rc = b200_fft1d_execute(X[e], direction, cur, dst, stage_scale, cstream);
                if (rc) return rc;
            }
            {   // the mark is emitted on every rank, also with an empty box: all ranks report the same list of stages
                char label[40];
                std::snprintf(label, sizeof(label), "fft%d (local)", e);
                long long const count = X[e] ? lp.out_shape[e][me].count() : 0;
                long long const out_count = (tkind == kind_r2c and e == 0) ? (is_backward ? count : (X[e] ? lp.in_shape[1][me].count() : 0)) : count;
                long long const in_count = (tkind == kind_r2c and e == 0 and is_backward) ? (X[e] ? lp.in_shape[1][me].count() : 0) : count;
                mark(label, in_count * bytes_of(e, false) + out_count * bytes_of(e, true), 0);
            }
            cur = dst; cur_buffer = dst_buffer;     // also without a transform (empty box): every rank follows the same buffers
        }
    }
    if (cur != out and landed_direct){
        // what the other GPUs sent sits in the arena at its final position inside my box: move those sub-boxes only
        bool const real_out = (tkind == kind_r2c and is_backward) or not complex_data;
        int const elem = real_out ? real_bytes : cplx_bytes;
        shape const &from = is_backward ? lp.out_shape[0] : lp.in_shape[3];
        box3 const &mine = is_backward ? lp.in_shape[0][me] : lp.out_shape[3][me];
        long long moved = 0;
        for(int r=0; r<ccomm->size() and not mine.empty(); r++){
            if (r == me) continue;
            box3 const piece = mine.overlap(from[r]);
            if (piece.empty()) continue;
            idx const offset = mine.offset_of(piece.low);
            rc = b200_copy_subbox(elem, piece.osize(0), piece.osize(1), piece.osize(2), mine.osize(0), mine.osize(0) * mine.osize(1),
                                  mine.osize(0), mine.osize(0) * mine.osize(1), advance(cur, offset, elem), advance(out, offset, elem), cstream);
            if (rc) return rc;
            moved += piece.count();
        }
        mark("received sub-boxes to the caller's array", 2 * moved * elem, 0);
    }else if (cur != out){
        idx const count = is_backward ? inbox_count : outbox_count;
        bool const real_out = (tkind == kind_r2c and is_backward) or not complex_data;
        size_t const bytes = static_cast<size_t>(count) * (real_out ? real_bytes : cplx_bytes);
        if (count > 0 and cudaMemcpyAsync(out, cur, bytes, cudaMemcpyDeviceToDevice, cstream) != cudaSuccess)
            return fail(B200_ERR_CUDA, "cudaMemcpyAsync failed");
        mark("copy to the caller's array", 2 * static_cast<long long>(bytes), 0);
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
    int rc = ensure_executors(precision);
    if (rc) return rc;
    bool const through_peers = ensure_peer(precision);
    if ((workspace == nullptr or exec_workspace_count > workspace_count) and not through_peers){
        workspace = ensure_workspace(precision, 1);
        if (workspace == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
    }
    size_t const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    size_t const in_unit = (tkind == kind_c2c) ? 2 * real_bytes : real_bytes;
    size_t const out_unit = (tkind == kind_c2c or tkind == kind_r2c) ? 2 * real_bytes : real_bytes;
    for(int b=0; b<std::max(batch, 1); b++){
        const char *src = static_cast<const char*>(in) + b * inbox_count * in_unit;
        char *dst = static_cast<char*>(out) + b * outbox_count * out_unit;
        rc = through_peers ? run_peer(precision, false, src, dst, scale_factor(scaling)) : run(precision, false, src, dst, workspace, scale_factor(scaling));
        if (rc) return rc;
    }
    return B200_SUCCESS;
}

int transform3d::backward(int precision, int batch, const void *in, void *out, void *workspace, int scaling){
    if (precision != B200_PREC_FLOAT and precision != B200_PREC_DOUBLE) return fail(B200_ERR_INVALID, "bad precision");
    if (ccomm->size() > 1 and b200_peer_timed_out()) return fail(B200_ERR_PEER, "a peer GPU did not reach a barrier within HEFFTE_B200_BARRIER_TIMEOUT_S");
    int rc = ensure_executors(precision);
    if (rc) return rc;
    bool const through_peers = ensure_peer(precision);
    if ((workspace == nullptr or exec_workspace_count > workspace_count) and not through_peers){
        workspace = ensure_workspace(precision, 1);
        if (workspace == nullptr) return fail(B200_ERR_CUDA, "cannot allocate the workspace");
    }
    size_t const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    size_t const in_unit = (tkind == kind_c2c or tkind == kind_r2c) ? 2 * real_bytes : real_bytes;
    size_t const out_unit = (tkind == kind_c2c) ? 2 * real_bytes : real_bytes;
    for(int b=0; b<std::max(batch, 1); b++){
        const char *src = static_cast<const char*>(in) + b * outbox_count * in_unit;
        char *dst = static_cast<char*>(out) + b * inbox_count * out_unit;
        rc = through_peers ? run_peer(precision, true, src, dst, scale_factor(scaling)) : run(precision, true, src, dst, workspace, scale_factor(scaling));
        if (rc) return rc;
    }
    return B200_SUCCESS;
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
