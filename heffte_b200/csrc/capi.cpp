// C ABI of the b200 backend (include/heffte_b200.h): the reference's heffte_c.h entry points for the new backend id,
// plus communicator creation and host-only plan introspection.  Mirrors src/heffte_c.cpp:193-498 of the reference in
// behaviour: create returns 0 / 1 (invalid backend) / 2 (exception), destroy returns 3 on a corrupt handle.
#include "../../include/heffte_b200.h"

#include <cstdio>
#include <cstring>
#include <memory>

#include "transform.h"

namespace b200 {
void set_error(std::string const &message);
int fail(int code, std::string const &message);
}
using namespace b200;

struct heffte_comm_s {
    std::unique_ptr<communicator> impl;
};

namespace {

struct plan_state {
    std::unique_ptr<transform3d> fft;
    // device staging of the *_host entry points
    void *dev_in = nullptr, *dev_out = nullptr;
    size_t dev_in_bytes = 0, dev_out_bytes = 0;
    // complex staging of the real-data entry points (s2c / d2z / c2s / z2d) of a complex-to-complex plan
    void *promoted = nullptr;
    size_t promoted_bytes = 0;
    ~plan_state(){
        if (dev_in) cudaFree(dev_in);
        if (dev_out) cudaFree(dev_out);
        if (promoted) cudaFree(promoted);
    }
};

bool known_backend(int backend){ return backend >= Heffte_BACKEND_B200 and backend <= Heffte_BACKEND_B200_COS1; }

transform_kind kind_of(int backend, bool r2c){
    switch(backend){
        case Heffte_BACKEND_B200_COS:  return kind_cos;
        case Heffte_BACKEND_B200_SIN:  return kind_sin;
        case Heffte_BACKEND_B200_COS1: return kind_cos1;
        default: return r2c ? kind_r2c : kind_c2c;
    }
}

template<typename index>
box3 box_from(index const low[3], index const high[3], int const *order){
    box3 b({{static_cast<idx>(low[0]), static_cast<idx>(low[1]), static_cast<idx>(low[2])}},
           {{static_cast<idx>(high[0]), static_cast<idx>(high[1]), static_cast<idx>(high[2])}});
    if (order != nullptr) b.order = {{order[0], order[1], order[2]}};
    return b;
}
box3 box_from9(int const *nine){ return box_from(nine, nine + 3, nine + 6); }
void box_to9(box3 const &b, int *nine){
    for(int d=0; d<3; d++){ nine[d] = static_cast<int>(b.low[d]); nine[3+d] = static_cast<int>(b.high[d]); nine[6+d] = b.order[d]; }
}

plan_options options_from(int backend, heffte_plan_options const *o){
    plan_options p;
    p.use_reorder = (backend != Heffte_BACKEND_B200);   // reference defaults: cufft false, cufft_cos/sin true
    if (o != nullptr){
        p.use_reorder = (o->use_reorder != 0);
        p.algorithm = o->algorithm;
        // 0 slabs / 1 pencils: the caller's choice, executed as given; 2 (Heffte_B200_DECOMPOSITION_AUTO): the planner picks the
        // decomposition that moves less over NVLink (the reported sizes are those of the pencil plan, the reference's default)
        p.use_pencils = (o->use_pencils != 0);
        p.explicit_decomposition = (o->use_pencils == 0 or o->use_pencils == 1);
        p.use_gpu_aware = (o->use_gpu_aware != 0);
    }
    return p;
}

template<typename index>
int create_plan(int backend, void *stream, index const inbox_low[3], index const inbox_high[3], int const *inbox_order,
                index const outbox_low[3], index const outbox_high[3], int const *outbox_order,
                int r2c_direction, bool r2c, heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan, int subranks = -1){
    if (plan == nullptr) return 2;
    *plan = nullptr;
    if (not known_backend(backend)){ set_error("invalid backend id (this library implements Heffte_BACKEND_B200*)"); return 1; }
    if (comm == nullptr or not comm->impl){ set_error("null communicator"); return 2; }
    if (options != nullptr and (options->algorithm < 0 or options->algorithm > 3)){ set_error("invalid reshape algorithm"); return 2; }
    if (r2c and backend != Heffte_BACKEND_B200){ set_error("r2c plans use Heffte_BACKEND_B200"); return 1; }
    if (r2c and (r2c_direction < 0 or r2c_direction > 2)){ set_error("r2c_direction must be 0, 1 or 2"); return 2; }
    try{
        std::unique_ptr<plan_state> state(new plan_state());
        plan_options effective = options_from(backend, options);
        effective.subranks = subranks;
        state->fft.reset(new transform3d(kind_of(backend, r2c), box_from(inbox_low, inbox_high, inbox_order),
                                         box_from(outbox_low, outbox_high, outbox_order), r2c_direction,
                                         comm->impl.get(), effective, static_cast<cudaStream_t>(stream)));
        heffte_fft_plan *handle = new heffte_fft_plan;
        handle->backend_type = backend;
        handle->using_r2c = r2c ? 1 : 0;
        handle->fft = state.release();
        *plan = handle;
    }catch(std::exception &e){
        set_error(e.what());
        return 2;
    }
    return Heffte_SUCCESS;
}

plan_state* state_of(heffte_plan const plan){
    if (plan == nullptr or plan->fft == nullptr or not known_backend(plan->backend_type)) return nullptr;
    return static_cast<plan_state*>(plan->fft);
}

void run_or_report(heffte_plan const plan, int precision, int direction, void const *input, void *output, void *workspace, int scale){
    int rc = heffte_execute(plan, precision, direction, 1, input, output, workspace, scale);
    if (rc != 0) std::fprintf(stderr, "heffte(b200): transform failed (%d): %s\n", rc, heffte_last_error());
}

// The real-data entry points (heffte_forward_s2c / d2z, heffte_backward_c2s / z2d).  On an r2c plan they are the transform
// itself; on a complex-to-complex plan the reference promotes the real input with a zero imaginary part and drops the
// imaginary part of the result (fft3d::forward(real, complex) / backward(complex, real), include/heffte_fft3d.h:353-389 with
// the cuFFT executor's convert overloads, include/heffte_backend_cuda.h:527-545).
void run_real_side(heffte_plan const plan, int precision, int direction, void const *input, void *output, void *workspace, int scale){
    plan_state *s = state_of(plan);
    if (s == nullptr or s->fft->kind() != kind_c2c){ run_or_report(plan, precision, direction, input, output, workspace, scale); return; }
    transform3d &fft = *s->fft;
    size_t const count = static_cast<size_t>(fft.size_inbox());
    size_t const bytes = count * ((precision == B200_PREC_FLOAT) ? 8 : 16);
    int rc = 0;
    if (bytes > s->promoted_bytes){
        if (s->promoted){ cudaStreamSynchronize(fft.stream()); cudaFree(s->promoted); s->promoted = nullptr; s->promoted_bytes = 0; }
        if (cudaMalloc(&s->promoted, bytes) != cudaSuccess){ std::fprintf(stderr, "heffte(b200): cannot allocate the complex staging array\n"); return; }
        s->promoted_bytes = bytes;
    }
    if (direction == B200_FORWARD){
        rc = b200_convert_r2c(precision, static_cast<long long>(count), input, s->promoted, fft.stream());
        if (rc == 0) rc = heffte_execute(plan, precision, B200_FORWARD, 1, s->promoted, output, workspace, scale);
    }else{
        rc = heffte_execute(plan, precision, B200_BACKWARD, 1, input, s->promoted, workspace, scale);
        if (rc == 0) rc = b200_convert_c2r(precision, static_cast<long long>(count), s->promoted, output, fft.stream());
    }
    if (rc != 0) std::fprintf(stderr, "heffte(b200): transform failed (%d): %s\n", rc, heffte_last_error());
}

// communicator whose allgather answers from a fixed table: lets the real plan constructor run without any transport
struct table_context { std::vector<long long> table; };
int table_gather(void *context, const void *mine, void *all, size_t bytes){
    auto *t = static_cast<table_context*>(context);
    if (bytes == 18 * sizeof(long long)){        // the boxes of every rank
        std::memcpy(all, t->table.data(), t->table.size() * sizeof(long long));
        return 0;
    }
    // anything else (the plan signature): every rank of the table is this process
    size_t const ranks = t->table.size() / 18;
    for(size_t r=0; r<ranks; r++) std::memcpy(static_cast<char*>(all) + r * bytes, mine, bytes);
    return 0;
}

} // namespace

extern "C" {

const char* heffte_last_error(void){ return b200_last_error(); }

// ---- communicators ------------------------------------------------------------------------------------------------
int heffte_comm_create_self(heffte_comm *comm){
    if (comm == nullptr) return 2;
    *comm = new heffte_comm_s{std::unique_ptr<communicator>(make_self_communicator())};
    return 0;
}
int heffte_comm_nccl_unique_id(void *id128){
    std::string error;
    if (nccl_unique_id(id128, error) != 0) return fail(B200_ERR_NCCL, error);
    return 0;
}
int heffte_comm_create_nccl(int rank, int size, const void *id128, heffte_comm *comm){
    if (comm == nullptr or id128 == nullptr or rank < 0 or rank >= size) return fail(B200_ERR_INVALID, "bad arguments");
    std::string error;
    communicator *c = make_nccl_communicator(rank, size, id128, error);
    if (c == nullptr) return fail(B200_ERR_NCCL, error);
    *comm = new heffte_comm_s{std::unique_ptr<communicator>(c)};
    return 0;
}
int heffte_comm_create_callbacks(int rank, int size, heffte_allgather_fn gather, heffte_exchange_fn exchange, void *context, heffte_comm *comm){
    if (comm == nullptr or gather == nullptr or rank < 0 or rank >= size) return fail(B200_ERR_INVALID, "bad arguments");
    *comm = new heffte_comm_s{std::unique_ptr<communicator>(make_callback_communicator(rank, size, gather, exchange, context))};
    return 0;
}
int heffte_comm_create_threads(int size, const int *devices, heffte_comm *comms){
    if (comms == nullptr or size < 1) return fail(B200_ERR_INVALID, "bad arguments");
    std::string error;
    std::vector<communicator*> group = make_thread_communicators(size, devices, error);
    if (static_cast<int>(group.size()) != size) return fail(B200_ERR_INVALID, error);
    for(int r=0; r<size; r++) comms[r] = new heffte_comm_s{std::unique_ptr<communicator>(group[r])};
    return 0;
}
int heffte_comm_rank(heffte_comm comm){ return (comm and comm->impl) ? comm->impl->rank() : -1; }
int heffte_comm_size(heffte_comm comm){ return (comm and comm->impl) ? comm->impl->size() : -1; }
int heffte_comm_destroy(heffte_comm comm){ delete comm; return 0; }

// ---- plans -------------------------------------------------------------------------------------------------------------
int heffte_set_default_options(int backend, heffte_plan_options *options){
    if (not known_backend(backend) or options == nullptr) return 1;
    options->use_reorder = (backend == Heffte_BACKEND_B200) ? 0 : 1;
    options->algorithm = Heffte_RESHAPE_ALGORITHM_ALLTOALLV;
    options->use_pencils = 1;
    options->use_gpu_aware = 1;
    return Heffte_SUCCESS;
}

int heffte_plan_create(int backend, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                       int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                       heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan){
    return create_plan(backend, nullptr, inbox_low, inbox_high, inbox_order, outbox_low, outbox_high, outbox_order, -1, false, comm, options, plan);
}
int heffte_plan_create_r2c(int backend, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                           int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                           int r2c_direction, heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan){
    return create_plan(backend, nullptr, inbox_low, inbox_high, inbox_order, outbox_low, outbox_high, outbox_order, r2c_direction, true, comm, options, plan);
}
int heffte_plan_create_stream(int backend, void *cuda_stream, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                              int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                              int r2c_direction, heffte_comm const comm, heffte_plan_options const *options, heffte_plan *plan){
    return create_plan(backend, cuda_stream, inbox_low, inbox_high, inbox_order, outbox_low, outbox_high, outbox_order,
                       r2c_direction, r2c_direction >= 0, comm, options, plan);
}

int heffte_plan_create_subcomm(int backend, void *cuda_stream, int const inbox_low[3], int const inbox_high[3], int const *inbox_order,
                               int const outbox_low[3], int const outbox_high[3], int const *outbox_order,
                               int r2c_direction, heffte_comm const comm, heffte_plan_options const *options, int num_subranks, heffte_plan *plan){
    return create_plan(backend, cuda_stream, inbox_low, inbox_high, inbox_order, outbox_low, outbox_high, outbox_order,
                       r2c_direction, r2c_direction >= 0, comm, options, plan, (num_subranks > 0) ? num_subranks : -1);
}

int heffte_plan_create64(int backend, void *cuda_stream, long long const inbox_low[3], long long const inbox_high[3], int const *inbox_order,
                         long long const outbox_low[3], long long const outbox_high[3], int const *outbox_order,
                         int r2c_direction, heffte_comm const comm, heffte_plan_options const *options, int num_subranks, heffte_plan *plan){
    return create_plan(backend, cuda_stream, inbox_low, inbox_high, inbox_order, outbox_low, outbox_high, outbox_order,
                       r2c_direction, r2c_direction >= 0, comm, options, plan, (num_subranks > 0) ? num_subranks : -1);
}

int heffte_plan_destroy(heffte_plan plan){
    if (plan == nullptr) return Heffte_SUCCESS;
    if (not known_backend(plan->backend_type)) return 3;
    delete static_cast<plan_state*>(plan->fft);
    delete plan;
    return Heffte_SUCCESS;
}

long long heffte_size_inbox64(heffte_plan const plan){ auto *s = state_of(plan); return s ? s->fft->size_inbox() : -1; }
long long heffte_size_outbox64(heffte_plan const plan){ auto *s = state_of(plan); return s ? s->fft->size_outbox() : -1; }
long long heffte_size_workspace64(heffte_plan const plan){ auto *s = state_of(plan); return s ? s->fft->size_workspace() : -1; }
int heffte_size_inbox(heffte_plan const plan){ return static_cast<int>(heffte_size_inbox64(plan)); }
int heffte_size_outbox(heffte_plan const plan){ return static_cast<int>(heffte_size_outbox64(plan)); }
int heffte_size_workspace(heffte_plan const plan){ return static_cast<int>(heffte_size_workspace64(plan)); }
int heffte_get_backend(heffte_plan const plan){ return plan ? plan->backend_type : -1; }
int heffte_is_r2c(heffte_plan const plan){ return plan ? plan->using_r2c : -1; }
double heffte_get_scale_factor(heffte_plan const plan, int scale){ auto *s = state_of(plan); return s ? s->fft->scale_factor(scale) : 0.0; }

int heffte_b200_uses_peer_memory(heffte_plan const plan, int precision){
    plan_state *s = state_of(plan);
    if (s == nullptr or precision < 0 or precision > 1) return -1;
    return s->fft->uses_peer_memory(precision) ? 1 : 0;
}

int heffte_b200_stage_timing(heffte_plan const plan, int enable){
    plan_state *s = state_of(plan);
    if (s == nullptr) return -1;
    s->fft->enable_stage_timing(enable != 0);
    return 0;
}
int heffte_b200_stage_times(heffte_plan const plan, int max_entries, char *names, double *ms, long long *local_bytes, long long *sent_bytes){
    plan_state *s = state_of(plan);
    if (s == nullptr) return -1;
    if (cudaStreamSynchronize(s->fft->stream()) != cudaSuccess) return -1;
    auto records = s->fft->collect_stage_times();
    int n = 0;
    for(auto const &r : records){
        if (n >= max_entries) break;
        std::memcpy(names + 40 * n, r.name, 40);
        ms[n] = r.ms; local_bytes[n] = r.local_bytes; sent_bytes[n] = r.sent_bytes;
        n++;
    }
    return n;
}

int heffte_execute(heffte_plan const plan, int precision, int direction, int batch, void const *input, void *output, void *workspace, int scale){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    if (scale < 0 or scale > 2) return fail(B200_ERR_INVALID, "invalid scale");
    if (batch < 1) return fail(B200_ERR_INVALID, "batch must be positive");
    if (batch > 1 and workspace != nullptr){
        // the caller's workspace holds batch * size_workspace() elements; the stages run entry by entry on its first part
    }
    if (direction == B200_FORWARD) return s->fft->forward(precision, batch, input, output, workspace, scale);
    return s->fft->backward(precision, batch, input, output, workspace, scale);
}

int heffte_convolve(heffte_plan const plan, int precision, void const *input, void *output, void *workspace, void const *multiplier, int scale){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    if (scale < 0 or scale > 2) return fail(B200_ERR_INVALID, "invalid scale");
    return s->fft->convolve(precision, input, output, workspace, multiplier, scale);
}
int heffte_convolve_box(heffte_plan const plan, long long low[3], long long high[3], int order[3]){
    plan_state *s = state_of(plan);
    if (s == nullptr or low == nullptr or high == nullptr or order == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle or null argument");
    box3 const b = s->fft->convolve_box();
    for(int d=0; d<3; d++){ low[d] = b.low[d]; high[d] = b.high[d]; order[d] = b.order[d]; }
    return Heffte_SUCCESS;
}
int heffte_b200_prepare(heffte_plan const plan, int precision, int batch){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    return s->fft->prepare(precision, batch);
}

int heffte_b200_register_buffer(heffte_plan const plan, int precision, void *device_array, size_t bytes){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    return s->fft->register_buffer(precision, device_array, bytes);      // a null array (an empty box) still takes part in the collective
}

int heffte_b200_unregister_buffer(heffte_plan const plan, int precision, void *device_array){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    return s->fft->unregister_buffer(precision, device_array);
}

int heffte_execute_host(heffte_plan const plan, int precision, int direction, int batch, void const *host_input, void *host_output, int scale){
    plan_state *s = state_of(plan);
    if (s == nullptr) return fail(B200_ERR_INVALID, "invalid plan handle");
    transform3d &fft = *s->fft;
    size_t const real_bytes = (precision == B200_PREC_FLOAT) ? 4 : 8;
    bool const cplx_in = (fft.kind() == kind_c2c) or (fft.kind() == kind_r2c and direction == B200_BACKWARD);
    bool const cplx_out = (fft.kind() == kind_c2c) or (fft.kind() == kind_r2c and direction == B200_FORWARD);
    size_t const count_in = static_cast<size_t>((direction == B200_FORWARD) ? fft.size_inbox() : fft.size_outbox()) * batch;
    size_t const count_out = static_cast<size_t>((direction == B200_FORWARD) ? fft.size_outbox() : fft.size_inbox()) * batch;
    size_t const bytes_in = count_in * real_bytes * (cplx_in ? 2 : 1), bytes_out = count_out * real_bytes * (cplx_out ? 2 : 1);
    if (bytes_in > s->dev_in_bytes){
        if (s->dev_in) cudaFree(s->dev_in);
        s->dev_in = nullptr; s->dev_in_bytes = 0;
        if (cudaMalloc(&s->dev_in, bytes_in) != cudaSuccess) return fail(B200_ERR_CUDA, "cudaMalloc failed (device input staging)");
        s->dev_in_bytes = bytes_in;
    }
    if (bytes_out > s->dev_out_bytes){
        if (s->dev_out) cudaFree(s->dev_out);
        s->dev_out = nullptr; s->dev_out_bytes = 0;
        if (cudaMalloc(&s->dev_out, bytes_out) != cudaSuccess) return fail(B200_ERR_CUDA, "cudaMalloc failed (device output staging)");
        s->dev_out_bytes = bytes_out;
    }
    cudaStream_t stream = fft.stream();
    if (bytes_in > 0 and cudaMemcpyAsync(s->dev_in, host_input, bytes_in, cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return fail(B200_ERR_CUDA, "host to device copy failed");
    int rc = heffte_execute(plan, precision, direction, batch, s->dev_in, s->dev_out, nullptr, scale);
    if (rc) return rc;
    if (bytes_out > 0 and cudaMemcpyAsync(host_output, s->dev_out, bytes_out, cudaMemcpyDeviceToHost, stream) != cudaSuccess)
        return fail(B200_ERR_CUDA, "device to host copy failed");
    if (cudaStreamSynchronize(stream) != cudaSuccess) return fail(B200_ERR_CUDA, "stream synchronisation failed");
    return 0;
}

void heffte_forward_s2c(heffte_plan const plan, float const *input, void *output, int scale){ run_real_side(plan, B200_PREC_FLOAT, B200_FORWARD, input, output, nullptr, scale); }
void heffte_forward_c2c(heffte_plan const plan, void const *input, void *output, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_FORWARD, input, output, nullptr, scale); }
void heffte_forward_d2z(heffte_plan const plan, double const *input, void *output, int scale){ run_real_side(plan, B200_PREC_DOUBLE, B200_FORWARD, input, output, nullptr, scale); }
void heffte_forward_z2z(heffte_plan const plan, void const *input, void *output, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_FORWARD, input, output, nullptr, scale); }
void heffte_forward_s2c_buffered(heffte_plan const plan, float const *input, void *output, void *workspace, int scale){ run_real_side(plan, B200_PREC_FLOAT, B200_FORWARD, input, output, workspace, scale); }
void heffte_forward_c2c_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_FORWARD, input, output, workspace, scale); }
void heffte_forward_d2z_buffered(heffte_plan const plan, double const *input, void *output, void *workspace, int scale){ run_real_side(plan, B200_PREC_DOUBLE, B200_FORWARD, input, output, workspace, scale); }
void heffte_forward_z2z_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_FORWARD, input, output, workspace, scale); }

void heffte_backward_c2s(heffte_plan const plan, void const *input, float *output, int scale){ run_real_side(plan, B200_PREC_FLOAT, B200_BACKWARD, input, output, nullptr, scale); }
void heffte_backward_c2c(heffte_plan const plan, void const *input, void *output, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_BACKWARD, input, output, nullptr, scale); }
void heffte_backward_z2d(heffte_plan const plan, void const *input, double *output, int scale){ run_real_side(plan, B200_PREC_DOUBLE, B200_BACKWARD, input, output, nullptr, scale); }
void heffte_backward_z2z(heffte_plan const plan, void const *input, void *output, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_BACKWARD, input, output, nullptr, scale); }
void heffte_backward_c2s_buffered(heffte_plan const plan, void const *input, float *output, void *workspace, int scale){ run_real_side(plan, B200_PREC_FLOAT, B200_BACKWARD, input, output, workspace, scale); }
void heffte_backward_c2c_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_BACKWARD, input, output, workspace, scale); }
void heffte_backward_z2d_buffered(heffte_plan const plan, void const *input, double *output, void *workspace, int scale){ run_real_side(plan, B200_PREC_DOUBLE, B200_BACKWARD, input, output, workspace, scale); }
void heffte_backward_z2z_buffered(heffte_plan const plan, void const *input, void *output, void *workspace, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_BACKWARD, input, output, workspace, scale); }

void heffte_forward_s2s_buffered(heffte_plan const plan, float const *input, float *output, float *workspace, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_FORWARD, input, output, workspace, scale); }
void heffte_forward_d2d_buffered(heffte_plan const plan, double const *input, double *output, double *workspace, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_FORWARD, input, output, workspace, scale); }
void heffte_backward_s2s_buffered(heffte_plan const plan, float const *input, float *output, float *workspace, int scale){ run_or_report(plan, B200_PREC_FLOAT, B200_BACKWARD, input, output, workspace, scale); }
void heffte_backward_d2d_buffered(heffte_plan const plan, double const *input, double *output, double *workspace, int scale){ run_or_report(plan, B200_PREC_DOUBLE, B200_BACKWARD, input, output, workspace, scale); }

// ---- host-only introspection ---------------------------------------------------------------------------------------------
int heffte_b200_logic_plan(int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                           int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                           int *shapes_out, int *fft_direction, long long *index_count){
    try{
        shape ins, outs;
        for(int r=0; r<nranks; r++){ ins.push_back(box_from9(inboxes + 9 * r)); outs.push_back(box_from9(outboxes + 9 * r)); }
        plan_options o;
        o.use_reorder = use_reorder != 0; o.algorithm = algorithm; o.use_pencils = use_pencils != 0; o.explicit_decomposition = (use_pencils == 0 or use_pencils == 1); o.subranks = subranks;
        logic_plan lp = make_logic_plan(ins, outs, r2c_direction, o, rank);
        for(int s=0; s<4; s++)
            for(int r=0; r<nranks; r++){
                box_to9(lp.in_shape[s][r], shapes_out + (s * nranks + r) * 9);
                box_to9(lp.out_shape[s][r], shapes_out + ((4 + s) * nranks + r) * 9);
            }
        for(int d=0; d<3; d++) fft_direction[d] = lp.fft_direction[d];
        *index_count = lp.index_count;
    }catch(std::exception &e){
        return fail(B200_ERR_INVALID, e.what());
    }
    return 0;
}

int heffte_b200_execution_plan(int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                               int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                               int *shapes_out, int *fft_direction, int *swaps){
    try{
        shape ins, outs;
        for(int r=0; r<nranks; r++){ ins.push_back(box_from9(inboxes + 9 * r)); outs.push_back(box_from9(outboxes + 9 * r)); }
        plan_options o;
        o.use_reorder = use_reorder != 0; o.algorithm = algorithm; o.use_pencils = use_pencils != 0; o.explicit_decomposition = (use_pencils == 0 or use_pencils == 1); o.subranks = subranks;
        logic_plan lp = make_execution_plan(ins, outs, r2c_direction, o, rank, swaps);
        for(int s=0; s<4; s++)
            for(int r=0; r<nranks; r++){
                box_to9(lp.in_shape[s][r], shapes_out + (s * nranks + r) * 9);
                box_to9(lp.out_shape[s][r], shapes_out + ((4 + s) * nranks + r) * 9);
            }
        for(int d=0; d<3; d++) fft_direction[d] = lp.fft_direction[d];
    }catch(std::exception &e){
        return fail(B200_ERR_INVALID, e.what());
    }
    return 0;
}

void heffte_b200_make_procgrid(int nprocs, int *grid2){ auto g = grid2d(nprocs); grid2[0] = g[0]; grid2[1] = g[1]; }
void heffte_b200_proc_setup_min_surface(int const *world_box, int nprocs, int *grid3){
    auto g = grid_min_surface(box_from9(world_box), nprocs);
    for(int d=0; d<3; d++) grid3[d] = g[d];
}
void heffte_b200_split_world(int const *world_box, int const *grid3, int *boxes_out){
    shape cells = split(box_from9(world_box), {{grid3[0], grid3[1], grid3[2]}});
    for(size_t i=0; i<cells.size(); i++) box_to9(cells[i], boxes_out + 9 * i);
}

int heffte_b200_reshape_pieces(int nranks, int const *inboxes, int const *outboxes, int me, int receive, long long *pieces_out, int max_pieces){
    try{
        shape ins, outs;
        for(int r=0; r<nranks; r++){ ins.push_back(box_from9(inboxes + 9 * r)); outs.push_back(box_from9(outboxes + 9 * r)); }
        std::unique_ptr<communicator> none(make_self_communicator());
        reshape_op op(ins, outs, me, none.get());
        auto const &list = receive ? op.recv_list() : op.send_list();
        if (static_cast<int>(list.size()) > max_pieces) return -2;
        for(size_t i=0; i<list.size(); i++){
            long long *p = pieces_out + 14 * i;
            piece const &q = list[i];
            p[0] = q.peer; p[1] = q.offset; p[2] = q.size[0]; p[3] = q.size[1]; p[4] = q.size[2];
            p[5] = q.line; p[6] = q.plane; p[7] = q.buff_line; p[8] = q.buff_plane;
            p[9] = q.map[0]; p[10] = q.map[1]; p[11] = q.map[2]; p[12] = q.count; p[13] = q.buffer_offset;
        }
        return static_cast<int>(list.size());
    }catch(std::exception &e){
        fail(B200_ERR_INVALID, e.what());
        return -1;
    }
}

int heffte_b200_plan_sizes(int kind, int nranks, int const *inboxes, int const *outboxes, int r2c_direction,
                           int use_reorder, int algorithm, int use_pencils, int subranks, int rank,
                           long long *size_inbox, long long *size_outbox, long long *size_workspace){
    try{
        table_context context;
        for(int r=0; r<nranks; r++){
            for(int i=0; i<9; i++) context.table.push_back(inboxes[9 * r + i]);
            for(int i=0; i<9; i++) context.table.push_back(outboxes[9 * r + i]);
        }
        std::unique_ptr<communicator> comm(make_callback_communicator(rank, nranks, table_gather, nullptr, &context));
        plan_options o;
        o.use_reorder = use_reorder != 0; o.algorithm = algorithm; o.use_pencils = use_pencils != 0; o.explicit_decomposition = (use_pencils == 0 or use_pencils == 1); o.subranks = subranks;
        transform3d fft(static_cast<transform_kind>(kind), box_from9(inboxes + 9 * rank), box_from9(outboxes + 9 * rank), r2c_direction,
                        comm.get(), o, nullptr);
        *size_inbox = fft.size_inbox(); *size_outbox = fft.size_outbox(); *size_workspace = fft.size_workspace();
    }catch(std::exception &e){
        return fail(B200_ERR_INVALID, e.what());
    }
    return 0;
}

} // extern "C"
