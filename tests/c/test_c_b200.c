/*
 * The scenario of the reference's C test (test/test_c.c: 4x4x4 world on two ranks split along the slow dimension, inputs
 * 0..31, transforms s2c / c2c / d2z / z2z, forward then backward with full scaling, plus the r2c plan of :180-227) written
 * against include/heffte_b200.h in plain C99.  The two ranks are host threads of this process; each owns a CUDA stream.
 * Golden values: test/test_c.c:47-74 (rank 0: 992, -32+32i, -32, -32-32i, -128+128i, -128, -128-128i; rank 1: -512) and the
 * sizes asserted at :149-151 and :225-227 (inbox 32, outbox 32 / 16, workspace 96 / 88 with the reference's options).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "heffte_b200.h"
#include "heffte_b200_kernels.h"

static int failures = 0;
#define CHECK(condition) do{ if (!(condition)){ printf("FAILED at %s:%d: %s\n", __FILE__, __LINE__, #condition); failures++; } }while(0)

typedef struct { heffte_comm comm; int rank; } rank_args;

static void* to_device(void *stream, const void *host, size_t bytes){
    void *d = NULL;
    CHECK(b200_device_alloc(bytes, &d) == 0);
    CHECK(b200_copy_to_device(host, d, bytes, stream) == 0);
    CHECK(b200_stream_synchronize(stream) == 0);
    return d;
}
static void to_host(void *stream, const void *device, void *host, size_t bytes){
    CHECK(b200_copy_to_host(device, host, bytes, stream) == 0);
    CHECK(b200_stream_synchronize(stream) == 0);
}
static double max_diff(const double *x, const double *y, int n){
    double e = 0; int i;
    for(i=0; i<n; i++) if (fabs(x[i] - y[i]) > e) e = fabs(x[i] - y[i]);
    return e;
}
static double max_diff_f(const float *x, const float *y, int n){
    double e = 0; int i;
    for(i=0; i<n; i++) if (fabs((double) x[i] - (double) y[i]) > e) e = fabs((double) x[i] - (double) y[i]);
    return e;
}

static void* rank_body(void *p){
    rank_args *args = (rank_args*) p;
    int const me = args->rank;
    int const low[3] = {0, 0, 2 * me}, high[3] = {3, 3, 2 * me + 1};
    void *stream = NULL;
    heffte_plan plan = NULL;
    heffte_plan_options options;
    int i;

    CHECK(b200_stream_create(&stream) == 0);
    CHECK(heffte_set_default_options(Heffte_BACKEND_B200, &options) == 0);
    options.use_reorder = 1;        /* the options of the reference's CPU run: sizes below are the reference's */
    CHECK(heffte_plan_create_stream(Heffte_BACKEND_B200, stream, low, high, NULL, low, high, NULL, -1, args->comm, &options, &plan) == 0);
    CHECK(heffte_size_inbox(plan) == 32 && heffte_size_outbox(plan) == 32 && heffte_size_workspace(plan) == 96);
    CHECK(heffte_get_backend(plan) == Heffte_BACKEND_B200 && heffte_is_r2c(plan) == 0);

    {   /* expected spectrum of the index-valued input (interleaved complex) */
        double expect[64]; float expect_f[64];
        double zin[64], zout[64], zback[64]; float cin[64], cout[64], cback[64]; double din[32]; float sin_[32];
        void *d_zin, *d_din, *d_cin, *d_sin, *d_out, *d_back, *d_work;
        memset(expect, 0, sizeof(expect));
        if (me == 0){
            expect[0] = 992.0; expect[2] = -32.0; expect[3] = 32.0; expect[4] = -32.0; expect[6] = -32.0; expect[7] = -32.0;
            expect[8] = -128.0; expect[9] = 128.0; expect[16] = -128.0; expect[24] = -128.0; expect[25] = -128.0;
        }else expect[0] = -512.0;
        for(i=0; i<64; i++) expect_f[i] = (float) expect[i];
        memset(zin, 0, sizeof(zin)); memset(cin, 0, sizeof(cin));
        for(i=0; i<32; i++){ zin[2*i] = i; cin[2*i] = (float) i; din[i] = i; sin_[i] = (float) i; }

        /* every device array is allocated before the first transform and freed after the last one: cudaFree waits for the whole
         * device, and the other rank (a thread of this process on the same GPU) may already be spinning in a barrier of the
         * next transform that only this thread's next launches can release */
        d_zin = to_device(stream, zin, sizeof(zin)); d_din = to_device(stream, din, sizeof(din));
        d_cin = to_device(stream, cin, sizeof(cin)); d_sin = to_device(stream, sin_, sizeof(sin_));
        CHECK(b200_device_alloc(sizeof(zout), &d_out) == 0 && b200_device_alloc(sizeof(zback), &d_back) == 0);
        CHECK(b200_device_alloc(96 * 2 * sizeof(double), &d_work) == 0);
        /* z2z, buffered, forward then backward with full scaling */
        heffte_forward_z2z_buffered(plan, d_zin, d_out, d_work, Heffte_SCALE_NONE);
        to_host(stream, d_out, zout, sizeof(zout));
        CHECK(max_diff(zout, expect, 64) < 1e-11);
        heffte_backward_z2z_buffered(plan, d_out, d_back, d_work, Heffte_SCALE_FULL);
        to_host(stream, d_back, zback, sizeof(zback));
        CHECK(max_diff(zback, zin, 64) < 1e-11);
        /* d2z / z2d: real input of the complex plan */
        heffte_forward_d2z(plan, (double const*) d_din, d_out, Heffte_SCALE_NONE);
        to_host(stream, d_out, zout, sizeof(zout));
        CHECK(max_diff(zout, expect, 64) < 1e-11);
        heffte_backward_z2d(plan, d_out, (double*) d_back, Heffte_SCALE_FULL);
        to_host(stream, d_back, zback, 32 * sizeof(double));
        CHECK(max_diff(zback, din, 32) < 1e-11);
        /* c2c and s2c / c2s in single precision */
        heffte_forward_c2c(plan, d_cin, d_out, Heffte_SCALE_NONE);
        to_host(stream, d_out, cout, sizeof(cout));
        CHECK(max_diff_f(cout, expect_f, 64) < 1e-4);
        heffte_backward_c2c(plan, d_out, d_back, Heffte_SCALE_FULL);
        to_host(stream, d_back, cback, sizeof(cback));
        CHECK(max_diff_f(cback, cin, 64) < 1e-4);
        heffte_forward_s2c(plan, (float const*) d_sin, d_out, Heffte_SCALE_NONE);
        to_host(stream, d_out, cout, sizeof(cout));
        CHECK(max_diff_f(cout, expect_f, 64) < 1e-4);
        heffte_backward_c2s(plan, d_out, (float*) d_back, Heffte_SCALE_FULL);
        to_host(stream, d_back, cback, 32 * sizeof(float));
        CHECK(max_diff_f(cback, sin_, 32) < 1e-4);
        b200_device_free(d_zin); b200_device_free(d_din); b200_device_free(d_cin); b200_device_free(d_sin);
        b200_device_free(d_out); b200_device_free(d_back); b200_device_free(d_work);
    }
    CHECK(heffte_plan_destroy(plan) == 0);

    {   /* r2c along dimension 2: rank 0 keeps two of the three complex planes, rank 1 the last one */
        int const clow[3] = {0, 0, (me == 0) ? 0 : 2}, chigh[3] = {3, 3, (me == 0) ? 1 : 2};
        int const nout = (me == 0) ? 32 : 16;
        double din[32], dback[32]; void *d_in, *d_out, *d_back;
        CHECK(heffte_plan_create_stream(Heffte_BACKEND_B200, stream, low, high, NULL, clow, chigh, NULL, 2, args->comm, &options, &plan) == 0);
        CHECK(heffte_size_inbox(plan) == 32 && heffte_size_outbox(plan) == nout && heffte_size_workspace(plan) == ((me == 0) ? 96 : 88));
        CHECK(heffte_is_r2c(plan) == 1);
        for(i=0; i<32; i++) din[i] = i;
        d_in = to_device(stream, din, sizeof(din));
        CHECK(b200_device_alloc((size_t) nout * 2 * sizeof(double), &d_out) == 0 && b200_device_alloc(sizeof(dback), &d_back) == 0);
        heffte_forward_d2z(plan, (double const*) d_in, d_out, Heffte_SCALE_SYMMETRIC);
        heffte_backward_z2d(plan, d_out, (double*) d_back, Heffte_SCALE_SYMMETRIC);
        to_host(stream, d_back, dback, sizeof(dback));
        CHECK(max_diff(dback, din, 32) < 1e-11);
        b200_device_free(d_in); b200_device_free(d_out); b200_device_free(d_back);
        CHECK(heffte_plan_destroy(plan) == 0);
    }
    /* error conventions of src/heffte_c.cpp:232, 273-277 */
    CHECK(heffte_plan_create_stream(-7, stream, low, high, NULL, low, high, NULL, -1, args->comm, &options, &plan) == 1);
    CHECK(b200_stream_destroy(stream) == 0);
    return NULL;
}

int main(void){
    heffte_comm comms[2];
    rank_args args[2];
    pthread_t other;
    if (b200_device_count() < 1){ printf("no CUDA device: the b200 backend has no CPU fallback\n"); return 2; }
    if (heffte_comm_create_threads(2, NULL, comms) != 0){ printf("cannot create the communicators: %s\n", heffte_last_error()); return 1; }
    args[0].comm = comms[0]; args[0].rank = 0;
    args[1].comm = comms[1]; args[1].rank = 1;
    pthread_create(&other, NULL, rank_body, &args[1]);
    rank_body(&args[0]);
    pthread_join(other, NULL);
    heffte_comm_destroy(comms[0]); heffte_comm_destroy(comms[1]);
    printf(failures == 0 ? "test_c_b200: ok\n" : "test_c_b200: FAILED\n");
    return failures == 0 ? 0 : 1;
}
