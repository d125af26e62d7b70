// User program against the C++ front-end (include/heffte_b200.hpp), in the shape of the reference's examples
// (examples/heffte_example_gpu.cpp, heffte_example_r2c.cpp, heffte_example_r2r.cpp) and of test/test_c.c: a 4x4x4 world on
// two ranks split along the slow dimension, each rank filling its 32 entries with 0..31.  The two ranks are host threads of
// this process sharing the visible GPU(s).  Expected spectrum entries are the values asserted by test/test_c.c:47-74.
#include <cmath>
#include <cstdio>
#include <numeric>
#include <thread>

#include "heffte_b200.hpp"

using backend_tag = heffte::backend::b200;

static int failures = 0;

void compute_dft(heffte::comm const &comm){
    int const me = comm.rank();
    heffte::box3d<> const left_box  = {{0, 0, 0}, {3, 3, 1}};
    heffte::box3d<> const right_box = {{0, 0, 2}, {3, 3, 3}};
    heffte::box3d<> const my_box = (me == 0) ? left_box : right_box;
    // ranks that are threads of one process: one CUDA stream per rank (the barrier between the ranks is stream ordered)
    heffte::gpu::stream stream;

    heffte::fft3d<backend_tag> fft(stream.get(), my_box, my_box, comm);
    if (fft.size_inbox() != 32 or fft.size_outbox() != 32 or fft.size_workspace() < 32){ std::printf("rank %d: wrong sizes\n", me); failures++; return; }

    std::vector<std::complex<double>> input(fft.size_inbox());
    std::iota(input.begin(), input.end(), 0);
    heffte::gpu::vector<std::complex<double>> gpu_input = heffte::gpu::transfer().load(stream.get(), input);
    heffte::gpu::vector<std::complex<double>> gpu_output(fft.size_outbox());
    heffte::fft3d<backend_tag>::buffer_container<std::complex<double>> workspace(fft.size_workspace());

    // optional, collective: the other rank then stores its part of the spectrum straight into this array (several GPUs: over NVLink)
    bool const registered = fft.register_buffer(gpu_output.data(), gpu_output.size());
    if (registered != fft.uses_peer_memory()){ std::printf("rank %d: register_buffer answered %d\n", me, (int) registered); failures++; }
    fft.forward(gpu_input.data(), gpu_output.data(), workspace.data());
    std::vector<std::complex<double>> spectrum = heffte::gpu::transfer::unload(stream.get(), gpu_output);
    // test/test_c.c:47-74
    std::vector<std::complex<double>> expect(32, 0.0);
    if (me == 0){
        expect[0] = {992.0, 0.0}; expect[1] = {-32.0, 32.0}; expect[2] = {-32.0, 0.0}; expect[3] = {-32.0, -32.0};
        expect[4] = {-128.0, 128.0}; expect[8] = {-128.0, 0.0}; expect[12] = {-128.0, -128.0};
    }else expect[0] = {-512.0, 0.0};
    double err = 0.0;
    for(size_t i=0; i<32; i++) err = std::max(err, std::abs(spectrum[i] - expect[i]));

    heffte::gpu::vector<std::complex<double>> gpu_inverse = fft.backward(gpu_output, heffte::scale::full);
    std::vector<std::complex<double>> inverse = heffte::gpu::transfer::unload(stream.get(), gpu_inverse);
    for(size_t i=0; i<input.size(); i++) err = std::max(err, std::abs(inverse[i] - input[i]));

    // real-to-complex along dimension 2 (test/test_c.c:180-227): rank 0 keeps planes k = 0..1 of the 3 complex planes, rank 1 keeps k = 2
    heffte::box3d<> const cbox = (me == 0) ? heffte::box3d<>({0, 0, 0}, {3, 3, 1}) : heffte::box3d<>({0, 0, 2}, {3, 3, 2});
    heffte::fft3d_r2c<backend_tag> rfft(stream.get(), my_box, cbox, 2, comm);
    if (rfft.size_outbox() != (me == 0 ? 32u : 16u)){ std::printf("rank %d: wrong r2c sizes\n", me); failures++; }
    std::vector<float> rinput(32);
    std::iota(rinput.begin(), rinput.end(), 0.0f);
    auto gpu_rin = heffte::gpu::transfer::load(stream.get(), rinput);
    auto gpu_rout = rfft.forward(gpu_rin);
    auto gpu_rback = rfft.backward(gpu_rout, heffte::scale::full);
    std::vector<float> rback = heffte::gpu::transfer::unload(stream.get(), gpu_rback);
    double rerr = 0.0;
    for(size_t i=0; i<32; i++) rerr = std::max(rerr, std::abs(double(rback[i]) - double(rinput[i])));

    // cosine transform (examples/heffte_example_r2r.cpp): forward with full scaling, then backward, returns the input
    heffte::fft3d<heffte::backend::b200_cos> cfft(stream.get(), my_box, my_box, comm);
    std::vector<double> dinput(32);
    std::iota(dinput.begin(), dinput.end(), 1.0);
    auto gpu_din = heffte::gpu::transfer::load(stream.get(), dinput);
    auto gpu_dct = cfft.forward(gpu_din, heffte::scale::full);
    auto gpu_dback = cfft.backward(gpu_dct);
    std::vector<double> dback = heffte::gpu::transfer::unload(stream.get(), gpu_dback);
    double derr = 0.0;
    for(size_t i=0; i<32; i++) derr = std::max(derr, std::abs(dback[i] - dinput[i]));

    std::printf("rank %d computed error: c2c %.3e  r2c %.3e  cos %.3e\n", me, err, rerr, derr);
    if (not (err < 1e-11 and rerr < 1e-4 and derr < 1e-11)) failures++;
}

int main(){
    if (heffte::gpu::device_count() < 1){ std::printf("no CUDA device: the b200 backend has no CPU fallback\n"); return 2; }
    std::vector<heffte::comm> ranks = heffte::comm::threads(2);
    std::thread other([&]{ compute_dft(ranks[1]); });
    compute_dft(ranks[0]);
    other.join();
    // single rank, caller stream left at the default, options object of the reference
    {
        heffte::comm self = heffte::comm::self();
        heffte::box3d<> const world = {{0, 0, 0}, {15, 7, 9}};
        heffte::plan_options options = heffte::default_options<backend_tag>();
        options.use_reorder = true;
        options.algorithm = heffte::reshape_algorithm::p2p_plined;
        options.use_pencils = false;
        heffte::fft3d<backend_tag> fft(world, world, self, options);
        std::vector<float> x(fft.size_inbox(), 1.0f);
        auto gx = heffte::gpu::transfer::load(x);
        heffte::gpu::vector<std::complex<float>> gy(fft.size_outbox());
        fft.forward(gx.data(), gy.data(), heffte::scale::symmetric);     // real input of a complex plan
        auto y = heffte::gpu::transfer::unload(gy);
        double const expect0 = std::sqrt(double(world.count()));         // constant field: all energy in the zero mode
        if (std::abs(y[0] - std::complex<float>(float(expect0), 0.0f)) > 1e-3 or std::abs(y[1]) > 1e-3){ std::printf("single rank: wrong spectrum\n"); failures++; }
    }
    // factories, aliases, 64-bit boxes and the std::vector overloads (include/heffte_fft3d.h:417-447, 703-763; test/test_longlong.cpp)
    {
        heffte::comm self = heffte::comm::self();
        heffte::box3d<long long> const world = {{0, 0, 0}, {7, 5, 11}};
        auto fft = heffte::make_fft3d<backend_tag>(world, world, self);
        std::vector<std::complex<double>> x(fft.size_inbox());
        for(size_t i=0; i<x.size(); i++) x[i] = {double(i % 7), double(i % 3)};
        std::vector<std::complex<double>> y = fft.forward(x);                       // host vectors in, host vectors out
        std::vector<std::complex<double>> back = fft.backward(y, heffte::scale::full);
        double err = 0.0;
        for(size_t i=0; i<x.size(); i++) err = std::max(err, std::abs(back[i] - x[i]));
        std::complex<double> sum = 0.0;
        for(auto const &v : x) sum += v;
        err = std::max(err, std::abs(y[0] - sum));
        heffte::rtransform<heffte::backend::b200_sin, long long> dst(world, world, self);
        std::vector<double> r(dst.size_inbox());
        std::iota(r.begin(), r.end(), 1.0);
        std::vector<double> rb = dst.backward(dst.forward(r, heffte::scale::full));
        for(size_t i=0; i<r.size(); i++) err = std::max(err, std::abs(rb[i] - r[i]));
        heffte::box3d<> const plane = {{0, 0, 0}, {15, 11, 0}};
        heffte::fft2d<backend_tag> fft2(plane, plane, self);
        std::vector<std::complex<float>> p(fft2.size_inbox(), {1.0f, 0.0f});
        std::vector<std::complex<float>> q = fft2.forward(p);
        err = std::max(err, double(std::abs(q[0] - std::complex<float>(192.0f, 0.0f))) + double(std::abs(q[5])));
        std::vector<float> preal = fft2.backward_real(q, heffte::scale::full);
        err = std::max(err, double(std::abs(preal[7] - 1.0f)));
        if (not (err < 1e-4)){ std::printf("factories / vectors / 64-bit boxes: error %.3e\n", err); failures++; }
    }
    std::printf(failures == 0 ? "example_b200: ok\n" : "example_b200: FAILED\n");
    return failures == 0 ? 0 : 1;
}
