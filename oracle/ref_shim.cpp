/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
 *
 * C-ABI window onto the UNMODIFIED reference (icl-utk-edu/heffte v2.4.1), compiled in place from
 * /root/reference with the `stock` CPU backend (FFTW/MKL/MPI are absent from the image) against
 * the threads-as-ranks MPI stand-in in oracle/mpi_shim/.  The result, oracle/_ref/libheffte_ref.so,
 * is what tests/ uses as the ground truth ("kind": "reference") and what the committed golden
 * fixtures under tests/golden/ were generated from (tests/golden/make_golden.py).
 *
 * Nothing here re-implements the reference: every function forwards to the reference's own
 * templates (heffte::fft3d<backend::stock>, one_dim_backend<stock>::executor, plan_operations,
 * direct_packer/transpose_packer<tag::cpu>, compute_overlap_map_transpose_pack ...).
 */
#include "heffte.h"

#include <cstring>
#include <random>
#include <thread>

using namespace heffte;

namespace {

template<typename index_array>
box3d<> make_box(index_array const *nine){ // low[3], high[3], order[3]
    return box3d<>({nine[0], nine[1], nine[2]}, {nine[3], nine[4], nine[5]}, {nine[6], nine[7], nine[8]});
}
void put_box(box3d<> const &b, int *nine){
    for(int i=0; i<3; i++){ nine[i] = b.low[i]; nine[3+i] = b.high[i]; nine[6+i] = b.order[i]; }
}

plan_options make_options(int use_reorder, int algorithm, int use_pencils, int subranks){
    plan_options opts(use_reorder != 0, static_cast<reshape_algorithm>(algorithm), use_pencils != 0);
    if (subranks > 0) opts.use_num_subranks(subranks);
    return opts;
}

scale to_scale(int s){ return (s == 1) ? scale::full : ((s == 2) ? scale::symmetric : scale::none); }

template<typename backend_tag, typename precision>
int exec1d_c2c(box3d<> const box, int dim, int dir, void *data){
    using ctype = std::complex<precision>;
    auto fft = make_executor<backend_tag>(nullptr, box, dim);
    std::vector<ctype> work(fft->workspace_size());
    if (dir == 0) fft->forward(reinterpret_cast<ctype*>(data), work.data());
    else fft->backward(reinterpret_cast<ctype*>(data), work.data());
    return 0;
}
template<typename backend_tag, typename precision>
int exec1d_r2r(box3d<> const box, int dim, int dir, void *data){
    auto fft = make_executor<backend_tag>(nullptr, box, dim);
    std::vector<precision> work(fft->workspace_size());
    if (dir == 0) fft->forward(reinterpret_cast<precision*>(data), work.data());
    else fft->backward(reinterpret_cast<precision*>(data), work.data());
    return 0;
}
template<typename precision>
int exec1d_r2c(box3d<> const box, int dim, int dir, void *rdata, void *cdata){
    using ctype = std::complex<precision>;
    auto fft = make_executor_r2c<backend::stock>(nullptr, box, dim);
    std::vector<ctype> work(fft->workspace_size());
    if (dir == 0) fft->forward(reinterpret_cast<precision const*>(rdata), reinterpret_cast<ctype*>(cdata), work.data());
    else fft->backward(reinterpret_cast<ctype*>(cdata), reinterpret_cast<precision*>(rdata), work.data());
    return 0;
}

// distributed driver: every thread-rank builds the reference plan on its boxes and transforms its piece
struct dist_job {
    int kind;       // 0 c2c, 1 r2c, 2 cos, 3 sin, 4 cos1
    int prec;       // 0 float, 1 double
    int nranks;
    int const *inboxes, *outboxes; // nranks * 9
    int r2c_dir, dir, scaling;
    int use_reorder, algorithm, use_pencils, subranks;
    void **inputs;  // per rank local arrays
    void **outputs;
    long long *workspace_sizes; // per rank, out
};
dist_job const *active_job = nullptr;

template<typename backend_tag, typename in_type, typename out_type>
void run_c2c_like(dist_job const &job, int me){
    box3d<> const inbox = make_box(job.inboxes + 9 * me), outbox = make_box(job.outboxes + 9 * me);
    fft3d<backend_tag> fft(inbox, outbox, MPI_COMM_WORLD, make_options(job.use_reorder, job.algorithm, job.use_pencils, job.subranks));
    if (job.workspace_sizes) job.workspace_sizes[me] = static_cast<long long>(fft.size_workspace());
    if (job.inputs == nullptr) return;
    if (job.dir == 0) fft.forward(reinterpret_cast<in_type const*>(job.inputs[me]), reinterpret_cast<out_type*>(job.outputs[me]), to_scale(job.scaling));
    else fft.backward(reinterpret_cast<out_type const*>(job.inputs[me]), reinterpret_cast<in_type*>(job.outputs[me]), to_scale(job.scaling));
}
template<typename precision>
void run_r2c(dist_job const &job, int me){
    using ctype = std::complex<precision>;
    box3d<> const inbox = make_box(job.inboxes + 9 * me), outbox = make_box(job.outboxes + 9 * me);
    fft3d_r2c<backend::stock> fft(inbox, outbox, job.r2c_dir, MPI_COMM_WORLD, make_options(job.use_reorder, job.algorithm, job.use_pencils, job.subranks));
    if (job.workspace_sizes) job.workspace_sizes[me] = static_cast<long long>(fft.size_workspace());
    if (job.inputs == nullptr) return;
    if (job.dir == 0) fft.forward(reinterpret_cast<precision const*>(job.inputs[me]), reinterpret_cast<ctype*>(job.outputs[me]), to_scale(job.scaling));
    else fft.backward(reinterpret_cast<ctype const*>(job.inputs[me]), reinterpret_cast<precision*>(job.outputs[me]), to_scale(job.scaling));
}

int dist_rank_main(int, char**){
    dist_job const &job = *active_job;
    int const me = mpi::comm_rank(MPI_COMM_WORLD);
    try{
        switch(job.kind){
            case 0:
                if (job.prec == 0) run_c2c_like<backend::stock, std::complex<float>, std::complex<float>>(job, me);
                else               run_c2c_like<backend::stock, std::complex<double>, std::complex<double>>(job, me);
                break;
            case 1:
                if (job.prec == 0) run_r2c<float>(job, me); else run_r2c<double>(job, me);
                break;
            case 2:
                if (job.prec == 0) run_c2c_like<backend::stock_cos, float, float>(job, me);
                else               run_c2c_like<backend::stock_cos, double, double>(job, me);
                break;
            case 3:
                if (job.prec == 0) run_c2c_like<backend::stock_sin, float, float>(job, me);
                else               run_c2c_like<backend::stock_sin, double, double>(job, me);
                break;
            case 4:
                if (job.prec == 0) run_c2c_like<backend::stock_cos1, float, float>(job, me);
                else               run_c2c_like<backend::stock_cos1, double, double>(job, me);
                break;
            default: return 1;
        }
    }catch(std::exception &e){
        std::fprintf(stderr, "reference threw on rank %d: %s\n", me, e.what());
        return 2;
    }
    return 0;
}

} // namespace

extern "C" {

int ref_version(void){ return 10000 * Heffte_VERSION_MAJOR + 100 * Heffte_VERSION_MINOR + Heffte_VERSION_PATCH; }

/* test/test_fft3d.h:19-38 make_data(): minstd_rand(4242) -> U(0,1), as doubles (the caller casts) */
void ref_make_data(long long count, double *out){
    std::minstd_rand park_miller(4242);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    for(long long i=0; i<count; i++) out[i] = unif(park_miller);
}

/* one_dim_backend<stock>::executor on one box (include/heffte_backend_stock.h:446-535); dir 0 forward, 1 backward */
int ref_exec1d_c2c(int prec, int const *box9, int dim, int dir, void *data){
    box3d<> const box = make_box(box9);
    return (prec == 0) ? exec1d_c2c<backend::stock, float>(box, dim, dir, data) : exec1d_c2c<backend::stock, double>(box, dim, dir, data);
}
/* executor_r2c (include/heffte_backend_stock.h:548-626): forward real->complex (box.r2c(dim) layout), backward complex->real */
int ref_exec1d_r2c(int prec, int const *box9, int dim, int dir, void *rdata, void *cdata){
    box3d<> const box = make_box(box9);
    return (prec == 0) ? exec1d_r2c<float>(box, dim, dir, rdata, cdata) : exec1d_r2c<double>(box, dim, dir, rdata, cdata);
}
/* real2real_executor<stock, {cos,sin,cos1}> (include/heffte_r2r_executor.h:191-278); kind 2 cos, 3 sin, 4 cos1 */
int ref_exec1d_r2r(int prec, int kind, int const *box9, int dim, int dir, void *data){
    box3d<> const box = make_box(box9);
    if (kind == 2) return (prec == 0) ? exec1d_r2r<backend::stock_cos, float>(box, dim, dir, data) : exec1d_r2r<backend::stock_cos, double>(box, dim, dir, data);
    if (kind == 3) return (prec == 0) ? exec1d_r2r<backend::stock_sin, float>(box, dim, dir, data) : exec1d_r2r<backend::stock_sin, double>(box, dim, dir, data);
    if (kind == 4) return (prec == 0) ? exec1d_r2r<backend::stock_cos1, float>(box, dim, dir, data) : exec1d_r2r<backend::stock_cos1, double>(box, dim, dir, data);
    return 1;
}

/*
 * heffte::fft3d / fft3d_r2c <backend::stock> run on `nranks` thread-ranks (include/heffte_fft3d.h, heffte_fft3d_r2c.h).
 * inputs/outputs may be NULL to only query size_workspace() per rank.
 */
int ref_fft3d(int kind, int prec, int nranks, int const *inboxes, int const *outboxes, int r2c_dir, int dir, int scaling,
              int use_reorder, int algorithm, int use_pencils, int subranks,
              void **inputs, void **outputs, long long *workspace_sizes){
    dist_job job = {kind, prec, nranks, inboxes, outboxes, r2c_dir, dir, scaling, use_reorder, algorithm, use_pencils, subranks,
                    inputs, outputs, workspace_sizes};
    active_job = &job;
    int rc = shim_run(nranks, dist_rank_main, 0, nullptr);
    active_job = nullptr;
    return rc;
}

/* plan_operations (src/heffte_plan_logic.cpp:424-453): shapes[8][nranks][9] = in_shape[0..3], out_shape[0..3] */
int ref_plan_operations(int nranks, int const *inboxes, int const *outboxes, int r2c_dir,
                        int use_reorder, int algorithm, int use_pencils, int subranks, int my_rank,
                        int *shapes, int *fft_direction, long long *index_count){
    try{
        ioboxes<> boxes;
        for(int i=0; i<nranks; i++){
            boxes.in.push_back(make_box(inboxes + 9 * i));
            boxes.out.push_back(make_box(outboxes + 9 * i));
        }
        logic_plan3d<int> plan = plan_operations(boxes, r2c_dir, make_options(use_reorder, algorithm, use_pencils, subranks), my_rank);
        for(int s=0; s<4; s++){
            for(int i=0; i<nranks; i++){
                put_box(plan.in_shape[s][i],  shapes + (s * nranks + i) * 9);
                put_box(plan.out_shape[s][i], shapes + ((4 + s) * nranks + i) * 9);
            }
        }
        for(int i=0; i<3; i++) fft_direction[i] = plan.fft_direction[i];
        *index_count = plan.index_count;
    }catch(std::exception &e){
        std::fprintf(stderr, "reference plan_operations threw: %s\n", e.what());
        return 2;
    }
    return 0;
}

/* include/heffte_geometry.h:337-349, 643-691, 409-436 */
void ref_make_procgrid(int nprocs, int *grid2){ auto g = make_procgrid(nprocs); grid2[0] = g[0]; grid2[1] = g[1]; }
void ref_proc_setup_min_surface(int const *world9, int nprocs, int *grid3){
    auto g = proc_setup_min_surface(make_box(world9), nprocs);
    for(int i=0; i<3; i++) grid3[i] = g[i];
}
void ref_split_world(int const *world9, int const *grid3, int *boxes){
    auto list = split_world(make_box(world9), std::array<int, 3>{grid3[0], grid3[1], grid3[2]});
    for(size_t i=0; i<list.size(); i++) put_box(list[i], boxes + 9 * i);
}

/*
 * Overlap maps (src/heffte_reshape3d.cpp:125-206). transpose = 0 -> compute_overlap_map_direct_pack semantics are
 * file-local in the reference, so the direct case is obtained from the transpose map with identical orders
 * (then map = identity and the buffer strides equal the dense overlap strides).
 * plans: per entry 10 ints = size[3], line_stride, plane_stride, buff_line_stride, buff_plane_stride, map[3]
 */
int ref_overlap_map_transpose(int me, int nprocs, int const *destination9, int const *boxes,
                              int *proc, int *offset, int *sizes, int *plans){
    std::vector<box3d<>> list;
    for(int i=0; i<nprocs; i++) list.push_back(make_box(boxes + 9 * i));
    std::vector<int> vproc, voffset, vsizes;
    std::vector<pack_plan_3d<int>> vplans;
    compute_overlap_map_transpose_pack(me, nprocs, make_box(destination9), list, vproc, voffset, vsizes, vplans);
    for(size_t i=0; i<vproc.size(); i++){
        proc[i] = vproc[i]; offset[i] = voffset[i]; sizes[i] = vsizes[i];
        int *p = plans + 10 * i;
        for(int j=0; j<3; j++) p[j] = vplans[i].size[j];
        p[3] = vplans[i].line_stride; p[4] = vplans[i].plane_stride;
        p[5] = vplans[i].buff_line_stride; p[6] = vplans[i].buff_plane_stride;
        for(int j=0; j<3; j++) p[7+j] = vplans[i].map[j];
    }
    return static_cast<int>(vproc.size());
}

/* direct_packer<tag::cpu> / transpose_packer<tag::cpu> (include/heffte_pack3d.h:89-197); elem_bytes in {4, 8, 16} */
int ref_pack(int elem_bytes, int const *plan10, int mode, void const *src, void *dst){
    pack_plan_3d<int> plan = {{plan10[0], plan10[1], plan10[2]}, plan10[3], plan10[4], plan10[5], plan10[6], {plan10[7], plan10[8], plan10[9]}};
    auto run = [&](auto const *s, auto *d){
        if (mode == 0) direct_packer<tag::cpu>().pack(nullptr, plan, s, d);
        else if (mode == 1) direct_packer<tag::cpu>().unpack(nullptr, plan, s, d);
        else transpose_packer<tag::cpu>().unpack(nullptr, plan, s, d);
    };
    if (elem_bytes == 4) run(static_cast<float const*>(src), static_cast<float*>(dst));
    else if (elem_bytes == 8) run(static_cast<double const*>(src), static_cast<double*>(dst));
    else if (elem_bytes == 16) run(static_cast<std::complex<double> const*>(src), static_cast<std::complex<double>*>(dst));
    else return 1;
    return 0;
}

int ref_hardware_threads(void){ return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"
