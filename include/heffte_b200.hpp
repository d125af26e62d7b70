/*
 * heffte_b200.hpp -- C++ front-end of the B200 backend: the user-facing classes of heFFTe for the new backend tags,
 *
 *      heffte::fft3d<heffte::backend::b200>            complex-to-complex (and real input / real output overloads)
 *      heffte::fft3d_r2c<heffte::backend::b200>        real-to-complex
 *      heffte::fft3d<heffte::backend::b200_cos>        DCT-II/III   (also b200_sin: DST-II/III, b200_cos1: DCT-I)
 *
 * with the reference's vocabulary: box3d (inclusive low/high + order), plan_options, scale::none/full/symmetric,
 * reshape_algorithm, size_inbox()/size_outbox()/size_workspace(), forward()/backward() with optional caller workspace
 * and batch, get_scale_factor(), gpu::vector / gpu::transfer helpers.  Signatures follow icl-utk-edu/heffte v2.4.1
 * (include/heffte_fft3d.h:270-572, include/heffte_fft3d_r2c.h:76-380, include/heffte_geometry.h:67-133,
 * include/heffte_plan_logic.h:48-176, include/heffte_backend_vector.h:52-157); the implementation is new: every call
 * goes through the C ABI of libheffte_b200.so (include/heffte_b200.h), so this header needs neither nvcc nor CUDA headers.
 *
 * Differences from the reference, all forced by the platform:
 *   - the communicator is a heffte_comm handle (NCCL over NVLink, host threads, or single rank) where the reference takes
 *     an MPI_Comm: the image has no MPI.  heffte::comm is a small RAII owner for it.
 *   - all data pointers are DEVICE pointers (like the reference's cufft backend).
 *   - errors of the C layer become std::runtime_error (the reference throws from cuda::check_error,
 *     include/heffte_backend_cuda.h:49-60).
 * A program written against the reference compiles against this header after replacing the backend tag and the
 * communicator argument; tests/cpp/ holds such programs.  Everything lives in namespace heffte_b200; `heffte` is an alias of
 * it unless the reference's heffte.h has been included before (see the end of this file), so both can share a translation unit.
 */
#ifndef HEFFTE_B200_HPP
#define HEFFTE_B200_HPP

#include <array>
#include <complex>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "heffte_b200.h"

namespace heffte_b200 {

// ---- geometry: include/heffte_geometry.h:67-133 ----------------------------------------------------------------------
template<typename index = int>
struct box3d {
    box3d(std::array<index, 3> clow, std::array<index, 3> chigh) : box3d(clow, chigh, {0, 1, 2}) {}
    box3d(std::array<index, 3> clow, std::array<index, 3> chigh, std::array<int, 3> corder)
        : low(clow), high(chigh), size({chigh[0] - clow[0] + 1, chigh[1] - clow[1] + 1, chigh[2] - clow[2] + 1}), order(corder) {}
    bool empty() const { return size[0] <= 0 or size[1] <= 0 or size[2] <= 0; }
    long long count() const { return empty() ? 0 : static_cast<long long>(size[0]) * size[1] * size[2]; }
    index osize(int dimension) const { return size[order[dimension]]; }
    //! intersection with another box, keeps the order of this box
    box3d collide(box3d const &other) const {
        if (empty() or other.empty()) return box3d({0, 0, 0}, {-1, -1, -1}, order);
        return box3d({std::max(low[0], other.low[0]), std::max(low[1], other.low[1]), std::max(low[2], other.low[2])},
                     {std::min(high[0], other.high[0]), std::min(high[1], other.high[1]), std::min(high[2], other.high[2])}, order);
    }
    //! box of the non-redundant complex coefficients of a real transform along `dimension`
    box3d r2c(int dimension) const {
        if (empty()) return *this;
        std::array<index, 3> h = high;
        h[dimension] = low[dimension] + size[dimension] / 2;
        return box3d(low, h, order);
    }
    bool operator == (box3d const &o) const { return low == o.low and high == o.high; }
    bool operator != (box3d const &o) const { return not (*this == o); }
    std::array<index, 3> low, high, size;
    std::array<int, 3> order;
};

enum class scale { none = Heffte_SCALE_NONE, full = Heffte_SCALE_FULL, symmetric = Heffte_SCALE_SYMMETRIC };

// include/heffte_plan_logic.h:48-57 (same numeric values)
enum class reshape_algorithm { alltoallv = 0, alltoall = 3, p2p_plined = 1, p2p = 2 };

namespace backend {
    struct b200 {};       //!< c2c and r2c transforms on B200 GPUs
    struct b200_cos {};   //!< DCT-II forward / DCT-III backward  (reference tag cufft_cos)
    struct b200_sin {};   //!< DST-II forward / DST-III backward  (reference tag cufft_sin)
    struct b200_cos1 {};  //!< DCT-I                              (reference tag cufft_cos1)

    template<typename tag> struct is_enabled : std::false_type {};
    template<> struct is_enabled<b200> : std::true_type {};
    template<> struct is_enabled<b200_cos> : std::true_type {};
    template<> struct is_enabled<b200_sin> : std::true_type {};
    template<> struct is_enabled<b200_cos1> : std::true_type {};

    template<typename tag> struct c_id {};
    template<> struct c_id<b200> { static constexpr int value = Heffte_BACKEND_B200; };
    template<> struct c_id<b200_cos> { static constexpr int value = Heffte_BACKEND_B200_COS; };
    template<> struct c_id<b200_sin> { static constexpr int value = Heffte_BACKEND_B200_SIN; };
    template<> struct c_id<b200_cos1> { static constexpr int value = Heffte_BACKEND_B200_COS1; };

    //! false for the real-to-real tags (include/heffte_common.h:439, 463-543)
    template<typename tag> struct uses_fft_types : std::true_type {};
    template<> struct uses_fft_types<b200_cos> : std::false_type {};
    template<> struct uses_fft_types<b200_sin> : std::false_type {};
    template<> struct uses_fft_types<b200_cos1> : std::false_type {};

    template<typename tag> inline std::string name();
    template<> inline std::string name<b200>(){ return "b200"; }
    template<> inline std::string name<b200_cos>(){ return "b200-cos-type-II"; }
    template<> inline std::string name<b200_sin>(){ return "b200-sin-type-II"; }
    template<> inline std::string name<b200_cos1>(){ return "b200-cos-type-I"; }
}

//! reference include/heffte_backend_cuda.h:854-877: the FFT backend does not reorder by default, the r2r ones do
template<typename backend_tag> struct default_plan_options { static const bool use_reorder = not std::is_same<backend_tag, backend::b200>::value; };

//! use_pencils of plan_options: reads and assigns like the reference's bool; an ASSIGNED value is executed as given (pencils or
//! slabs), an untouched one lets the planner pick the decomposition that moves the fewest bytes over NVLink
struct decomposition_choice {
    decomposition_choice(bool pencils = true, bool chosen_by_caller = false) : value(pencils), chosen(chosen_by_caller) {}
    decomposition_choice& operator = (bool pencils){ value = pencils; chosen = true; return *this; }
    operator bool () const { return value; }
    int c_value() const { return chosen ? (value ? 1 : 0) : Heffte_B200_DECOMPOSITION_AUTO; }
    bool value, chosen;
};

// include/heffte_plan_logic.h:131-176
struct plan_options {
    template<typename backend_tag> plan_options(backend_tag const)
        : use_reorder(default_plan_options<backend_tag>::use_reorder), algorithm(reshape_algorithm::alltoallv), use_pencils(true), use_gpu_aware(true) {}
    plan_options(bool reorder, reshape_algorithm alg, bool pencils) : use_reorder(reorder), algorithm(alg), use_pencils(pencils, true), use_gpu_aware(true) {}
    bool use_reorder;
    reshape_algorithm algorithm;
    decomposition_choice use_pencils;
    bool use_gpu_aware;
    //! include/heffte_plan_logic.h:100-129: hold the intermediate stages on the first num_subranks ranks only
    void use_subcomm(int num_subranks){ num_sub = num_subranks; }
    int get_subranks() const { return num_sub; }
private:
    int num_sub = -1;
};
template<typename backend_tag> inline plan_options default_options(){ return plan_options(backend_tag()); }

namespace b200_detail {
    inline void check(int code, const char *what){
        if (code != 0) throw std::runtime_error(std::string(what) + ": " + heffte_last_error());
    }
    template<typename T> struct is_complex : std::false_type {};
    template<typename T> struct is_complex<std::complex<T>> : std::true_type {};
    template<typename T> struct precision_of { static constexpr int value = std::is_same<T, float>::value ? B200_PREC_FLOAT : B200_PREC_DOUBLE; };
    template<typename T> struct precision_of<std::complex<T>> { static constexpr int value = precision_of<T>::value; };
    template<typename T> struct real_of { using type = T; };
    template<typename T> struct real_of<std::complex<T>> { using type = T; };
}

// ---- communicator owner (stands where user code holds an MPI_Comm) -----------------------------------------------------
class comm {
public:
    comm() : handle(nullptr) {}
    explicit comm(heffte_comm adopted) : handle(adopted) {}
    comm(comm const&) = delete;
    comm& operator = (comm const&) = delete;
    comm(comm &&other) noexcept : handle(other.handle){ other.handle = nullptr; }
    comm& operator = (comm &&other) noexcept { std::swap(handle, other.handle); return *this; }
    ~comm(){ if (handle) heffte_comm_destroy(handle); }
    //! single rank
    static comm self(){ heffte_comm h = nullptr; b200_detail::check(heffte_comm_create_self(&h), "heffte_comm_create_self"); return comm(h); }
    //! one process per GPU over NCCL; id128 comes from nccl_unique_id() on rank 0 and is shipped to every rank by the caller
    static comm nccl(int rank, int size, const void *id128){
        heffte_comm h = nullptr; b200_detail::check(heffte_comm_create_nccl(rank, size, id128, &h), "heffte_comm_create_nccl"); return comm(h);
    }
    static std::array<char, 128> nccl_unique_id(){
        std::array<char, 128> id{}; b200_detail::check(heffte_comm_nccl_unique_id(id.data()), "heffte_comm_nccl_unique_id"); return id;
    }
    //! `size` ranks in this process, one host thread each, rank r on CUDA device devices[r]
    static std::vector<comm> threads(int size, std::vector<int> const &devices = {}){
        std::vector<heffte_comm> raw(size, nullptr);
        b200_detail::check(heffte_comm_create_threads(size, devices.empty() ? nullptr : devices.data(), raw.data()), "heffte_comm_create_threads");
        std::vector<comm> out;
        for(auto h : raw) out.emplace_back(h);
        return out;
    }
    int rank() const { return heffte_comm_rank(handle); }
    int size() const { return heffte_comm_size(handle); }
    heffte_comm get() const { return handle; }
private:
    heffte_comm handle;
};

// ---- device containers: include/heffte_backend_vector.h:52-157, include/heffte_backend_data_transfer.h:29-183 ------------
namespace gpu {
    template<typename T>
    class vector {
    public:
        using value_type = T;
        explicit vector(size_t count = 0, void *cuda_stream = nullptr) : stream(cuda_stream), num(count), ptr(nullptr){
            void *p = nullptr;
            b200_detail::check(b200_device_alloc(count * sizeof(T), &p), "b200_device_alloc");
            ptr = static_cast<T*>(p);
        }
        vector(vector const &other) : vector(other.num, other.stream){
            b200_detail::check(b200_copy_on_device(other.ptr, ptr, num * sizeof(T), stream), "b200_copy_on_device");
        }
        vector(vector &&other) noexcept : stream(other.stream), num(other.num), ptr(other.ptr){ other.num = 0; other.ptr = nullptr; }
        vector& operator = (vector other){ std::swap(stream, other.stream); std::swap(num, other.num); std::swap(ptr, other.ptr); return *this; }
        ~vector(){ if (ptr) b200_device_free(ptr); }
        T* data(){ return ptr; }
        T const* data() const { return ptr; }
        size_t size() const { return num; }
        bool empty() const { return num == 0; }
    private:
        void *stream;
        size_t num;
        T *ptr;
    };

    struct transfer {
        template<typename T> static vector<T> load(void *stream, std::vector<T> const &host){
            vector<T> result(host.size(), stream);
            b200_detail::check(b200_copy_to_device(host.data(), result.data(), host.size() * sizeof(T), stream), "b200_copy_to_device");
            b200_detail::check(b200_stream_synchronize(stream), "b200_stream_synchronize");
            return result;
        }
        template<typename T> static vector<T> load(std::vector<T> const &host){ return load(nullptr, host); }
        template<typename T> static std::vector<T> unload(void *stream, vector<T> const &device){
            std::vector<T> result(device.size());
            b200_detail::check(b200_copy_to_host(device.data(), result.data(), device.size() * sizeof(T), stream), "b200_copy_to_host");
            b200_detail::check(b200_stream_synchronize(stream), "b200_stream_synchronize");
            return result;
        }
        template<typename T> static std::vector<T> unload(vector<T> const &device){ return unload(nullptr, device); }
    };
    //! RAII non-blocking CUDA stream for callers that do not include the CUDA headers; ranks that are host threads of one
    //! process (comm::threads) must give every rank its own stream
    class stream {
    public:
        stream() : handle(nullptr){ b200_detail::check(b200_stream_create(&handle), "b200_stream_create"); }
        stream(stream const&) = delete;
        stream& operator = (stream const&) = delete;
        ~stream(){ if (handle) b200_stream_destroy(handle); }
        void* get() const { return handle; }
        void synchronize() const { b200_detail::check(b200_stream_synchronize(handle), "b200_stream_synchronize"); }
    private:
        void *handle;
    };
    inline int device_count(){ return b200_device_count(); }
    inline void device_set(int device){ b200_detail::check(b200_device_set(device), "b200_device_set"); }
    inline void synchronize_default_stream(){ b200_detail::check(b200_stream_synchronize(nullptr), "b200_stream_synchronize"); }
}

// ---- shared implementation of the two plan classes ----------------------------------------------------------------------
namespace b200_detail {
    template<typename index>
    class plan_base {
    public:
        plan_base(plan_base const&) = delete;
        plan_base& operator = (plan_base const&) = delete;
        plan_base(plan_base &&other) noexcept : plan(other.plan), cstream(other.cstream){ other.plan = nullptr; }
        //! the CUDA stream of the plan (cudaStream_t as void*, null = default stream)
        void* stream() const { return cstream; }
        ~plan_base(){ if (plan) heffte_plan_destroy(plan); }
        //! number of entries of the input / output / workspace arrays (in units of the respective element type)
        size_t size_inbox() const { return static_cast<size_t>(heffte_size_inbox64(plan)); }
        size_t size_outbox() const { return static_cast<size_t>(heffte_size_outbox64(plan)); }
        size_t size_workspace() const { return static_cast<size_t>(heffte_size_workspace64(plan)); }
        double get_scale_factor(scale scaling) const { return heffte_get_scale_factor(plan, static_cast<int>(scaling)); }
        //! true when the reshapes of this plan run through peer memory (NVLink stores fused into the FFT kernels)
        bool uses_peer_memory(int precision = B200_PREC_DOUBLE) const { return heffte_b200_uses_peer_memory(plan, precision) == 1; }
        //! \brief Collective: registers a device array this plan will be asked to write (see heffte_b200_register_buffer); true when registered.
        template<typename T> bool register_buffer(T *array, size_t num_entries){
            return heffte_b200_register_buffer(plan, b200_detail::precision_of<T>::value, array, num_entries * sizeof(T)) == 0;
        }
        template<typename T> void unregister_buffer(T *array){
            heffte_b200_unregister_buffer(plan, b200_detail::precision_of<T>::value, array);
        }
    protected:
        plan_base(int backend_id, void *stream, box3d<index> const &inbox, box3d<index> const &outbox, int r2c_direction, comm const &c, plan_options const &o)
            : plan(nullptr), cstream(stream){
            // 64-bit coordinates all the way down (box3d<long long>, reference test/test_longlong.cpp)
            long long const lo_in[3] = {static_cast<long long>(inbox.low[0]), static_cast<long long>(inbox.low[1]), static_cast<long long>(inbox.low[2])};
            long long const hi_in[3] = {static_cast<long long>(inbox.high[0]), static_cast<long long>(inbox.high[1]), static_cast<long long>(inbox.high[2])};
            long long const lo_out[3] = {static_cast<long long>(outbox.low[0]), static_cast<long long>(outbox.low[1]), static_cast<long long>(outbox.low[2])};
            long long const hi_out[3] = {static_cast<long long>(outbox.high[0]), static_cast<long long>(outbox.high[1]), static_cast<long long>(outbox.high[2])};
            heffte_plan_options opts{o.use_reorder ? 1 : 0, static_cast<int>(o.algorithm), o.use_pencils.c_value(), o.use_gpu_aware ? 1 : 0};
            int const code = heffte_plan_create64(backend_id, stream, lo_in, hi_in, inbox.order.data(), lo_out, hi_out, outbox.order.data(),
                                                  r2c_direction, c.get(), &opts, o.get_subranks(), &plan);
            if (code != 0) throw std::runtime_error(std::string("heffte::fft3d (b200) plan creation failed: ") + heffte_last_error());
        }
        void execute(int precision, int direction, int batch, void const *input, void *output, void *workspace, scale scaling) const {
            check(heffte_execute(plan, precision, direction, batch, input, output, workspace, static_cast<int>(scaling)), "heffte::fft3d (b200) transform");
        }
        heffte_plan plan;
        void *cstream;
    };
}

/*
 * heffte::fft3d<backend_tag, index>: include/heffte_fft3d.h:270-572.
 * b200:      forward(complex|real in, complex out), backward(complex in, complex|real out); in-place allowed for complex/complex.
 * b200_cos/sin/cos1: real in, real out.
 */
template<typename backend_tag, typename index = int>
class fft3d : public b200_detail::plan_base<index> {
    static_assert(backend::is_enabled<backend_tag>::value, "heffte_b200.hpp provides the backend::b200* tags only");
    using base = b200_detail::plan_base<index>;
    static constexpr bool is_fft = backend::uses_fft_types<backend_tag>::value;
public:
    using backend_type = backend_tag;
    template<typename T> using buffer_container = gpu::vector<T>;

    fft3d(box3d<index> const inbox, box3d<index> const outbox, comm const &c, plan_options const options = default_options<backend_tag>())
        : base(backend::c_id<backend_tag>::value, nullptr, inbox, outbox, -1, c, options) {}
    //! plan bound to a caller-owned CUDA stream (cudaStream_t passed as void*), include/heffte_fft3d.h:287-297
    fft3d(void *cuda_stream, box3d<index> const inbox, box3d<index> const outbox, comm const &c, plan_options const options = default_options<backend_tag>())
        : base(backend::c_id<backend_tag>::value, cuda_stream, inbox, outbox, -1, c, options) {}

    template<typename input_type, typename output_type>
    void forward(input_type const input[], output_type output[], scale scaling = scale::none) const { forward(1, input, output, static_cast<output_type*>(nullptr), scaling); }
    template<typename input_type, typename output_type>
    void forward(input_type const input[], output_type output[], output_type workspace[], scale scaling = scale::none) const { forward(1, input, output, workspace, scaling); }
    template<typename input_type, typename output_type>
    void forward(int batch_size, input_type const input[], output_type output[], scale scaling = scale::none) const { forward(batch_size, input, output, static_cast<output_type*>(nullptr), scaling); }
    template<typename input_type, typename output_type>
    void forward(int batch_size, input_type const input[], output_type output[], output_type workspace[], scale scaling = scale::none) const {
        check_types<input_type, output_type>();
        constexpr int prec = b200_detail::precision_of<output_type>::value;
        if (is_fft and not b200_detail::is_complex<input_type>::value){
            // real input of a complex plan: promote with a zero imaginary part (reference cufft executor, heffte_backend_cuda.h:527-536)
            gpu::vector<output_type> promoted(batch_size * this->size_inbox());
            b200_detail::check(b200_convert_r2c(prec, static_cast<long long>(promoted.size()), input, promoted.data(), this->stream()), "b200_convert_r2c");
            this->execute(prec, B200_FORWARD, batch_size, promoted.data(), output, workspace, scaling);
            b200_detail::check(b200_stream_synchronize(this->stream()), "b200_stream_synchronize");
        }else{
            this->execute(prec, B200_FORWARD, batch_size, input, output, workspace, scaling);
        }
    }

    template<typename input_type, typename output_type>
    void backward(input_type const input[], output_type output[], scale scaling = scale::none) const { backward(1, input, output, static_cast<input_type*>(nullptr), scaling); }
    template<typename input_type, typename output_type>
    void backward(input_type const input[], output_type output[], input_type workspace[], scale scaling = scale::none) const { backward(1, input, output, workspace, scaling); }
    template<typename input_type, typename output_type>
    void backward(int batch_size, input_type const input[], output_type output[], scale scaling = scale::none) const { backward(batch_size, input, output, static_cast<input_type*>(nullptr), scaling); }
    template<typename input_type, typename output_type>
    void backward(int batch_size, input_type const input[], output_type output[], input_type workspace[], scale scaling = scale::none) const {
        check_types<output_type, input_type>();
        constexpr int prec = b200_detail::precision_of<input_type>::value;
        if (is_fft and not b200_detail::is_complex<output_type>::value){
            gpu::vector<input_type> full(batch_size * this->size_inbox());
            this->execute(prec, B200_BACKWARD, batch_size, input, full.data(), workspace, scaling);
            b200_detail::check(b200_convert_c2r(prec, static_cast<long long>(full.size()), full.data(), output, this->stream()), "b200_convert_c2r");
            b200_detail::check(b200_stream_synchronize(this->stream()), "b200_stream_synchronize");
        }else{
            this->execute(prec, B200_BACKWARD, batch_size, input, output, workspace, scaling);
        }
    }

    //! container variants (include/heffte_fft3d.h:417-447, 517-542)
    template<typename T> gpu::vector<typename std::conditional<is_fft, std::complex<typename b200_detail::real_of<T>::type>, T>::type>
    forward(gpu::vector<T> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_inbox()) throw std::invalid_argument("The input vector is smaller than size_inbox(), i.e., not enough entries provided to fill the inbox.");
        gpu::vector<typename std::conditional<is_fft, std::complex<typename b200_detail::real_of<T>::type>, T>::type> output(this->size_outbox());
        forward(input.data(), output.data(), scaling);
        return output;
    }
    template<typename T> gpu::vector<T> backward(gpu::vector<T> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_outbox()) throw std::invalid_argument("The input vector is smaller than size_outbox(), i.e., not enough entries provided to fill the outbox.");
        gpu::vector<T> output(this->size_inbox());
        backward(input.data(), output.data(), scaling);
        return output;
    }
    //! std::vector variants (include/heffte_fft3d.h:417-447, 517-542): HOST data, staged through device vectors (H2D, transform, D2H)
    template<typename T> std::vector<typename std::conditional<is_fft, std::complex<typename b200_detail::real_of<T>::type>, T>::type>
    forward(std::vector<T> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_inbox()) throw std::invalid_argument("The input vector is smaller than size_inbox(), i.e., not enough entries provided to fill the inbox.");
        return gpu::transfer::unload(this->stream(), forward(gpu::transfer::load(this->stream(), input), scaling));
    }
    template<typename T> std::vector<T> backward(std::vector<T> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_outbox()) throw std::invalid_argument("The input vector is smaller than size_outbox(), i.e., not enough entries provided to fill the outbox.");
        return gpu::transfer::unload(this->stream(), backward(gpu::transfer::load(this->stream(), input), scaling));
    }
    //! complex spectrum back to a REAL field (include/heffte_fft3d.h:534-542)
    template<typename real> gpu::vector<real> backward_real(gpu::vector<std::complex<real>> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_outbox()) throw std::invalid_argument("The input vector is smaller than size_outbox(), i.e., not enough entries provided to fill the outbox.");
        gpu::vector<real> output(this->size_inbox());
        backward(input.data(), output.data(), scaling);
        return output;
    }
    template<typename real> std::vector<real> backward_real(std::vector<std::complex<real>> const &input, scale scaling = scale::none) const {
        return gpu::transfer::unload(this->stream(), backward_real(gpu::transfer::load(this->stream(), input), scaling));
    }
private:
    template<typename spatial_type, typename spectral_type> static void check_types(){
        using real = typename b200_detail::real_of<spectral_type>::type;
        static_assert(std::is_same<real, float>::value or std::is_same<real, double>::value, "heffte::fft3d works with float and double precision");
        static_assert(std::is_same<typename b200_detail::real_of<spatial_type>::type, real>::value, "input and output must have the same precision");
        static_assert(not is_fft or b200_detail::is_complex<spectral_type>::value, "the transformed data of a complex plan is complex (include/heffte_fft3d.h:355)");
        static_assert(is_fft or (not b200_detail::is_complex<spectral_type>::value and not b200_detail::is_complex<spatial_type>::value),
                      "the cosine / sine transforms work with real data");
    }
};

/*
 * heffte::fft3d_r2c<backend_tag, index>: include/heffte_fft3d_r2c.h:76-380.  forward(real in, complex out), backward(complex in, real out);
 * outbox is a box of the world shortened along r2c_direction (box3d::r2c).
 */
template<typename backend_tag, typename index = int>
class fft3d_r2c : public b200_detail::plan_base<index> {
    static_assert(std::is_same<backend_tag, backend::b200>::value, "fft3d_r2c uses backend::b200");
    using base = b200_detail::plan_base<index>;
public:
    using backend_type = backend_tag;
    fft3d_r2c(box3d<index> const inbox, box3d<index> const outbox, int r2c_direction, comm const &c, plan_options const options = default_options<backend_tag>())
        : base(Heffte_BACKEND_B200, nullptr, inbox, outbox, checked(r2c_direction), c, options) {}
    fft3d_r2c(void *cuda_stream, box3d<index> const inbox, box3d<index> const outbox, int r2c_direction, comm const &c,
              plan_options const options = default_options<backend_tag>())
        : base(Heffte_BACKEND_B200, cuda_stream, inbox, outbox, checked(r2c_direction), c, options) {}

    template<typename real>
    void forward(real const input[], std::complex<real> output[], scale scaling = scale::none) const { forward(1, input, output, static_cast<std::complex<real>*>(nullptr), scaling); }
    template<typename real>
    void forward(real const input[], std::complex<real> output[], std::complex<real> workspace[], scale scaling = scale::none) const { forward(1, input, output, workspace, scaling); }
    template<typename real>
    void forward(int batch_size, real const input[], std::complex<real> output[], std::complex<real> workspace[], scale scaling = scale::none) const {
        this->execute(b200_detail::precision_of<real>::value, B200_FORWARD, batch_size, input, output, workspace, scaling);
    }
    template<typename real>
    void backward(std::complex<real> const input[], real output[], scale scaling = scale::none) const { backward(1, input, output, static_cast<std::complex<real>*>(nullptr), scaling); }
    template<typename real>
    void backward(std::complex<real> const input[], real output[], std::complex<real> workspace[], scale scaling = scale::none) const { backward(1, input, output, workspace, scaling); }
    template<typename real>
    void backward(int batch_size, std::complex<real> const input[], real output[], std::complex<real> workspace[], scale scaling = scale::none) const {
        this->execute(b200_detail::precision_of<real>::value, B200_BACKWARD, batch_size, input, output, workspace, scaling);
    }
    template<typename real> gpu::vector<std::complex<real>> forward(gpu::vector<real> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_inbox()) throw std::invalid_argument("The input vector is smaller than size_inbox(), i.e., not enough entries provided to fill the inbox.");
        gpu::vector<std::complex<real>> output(this->size_outbox());
        forward(input.data(), output.data(), scaling);
        return output;
    }
    template<typename real> gpu::vector<real> backward(gpu::vector<std::complex<real>> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_outbox()) throw std::invalid_argument("The input vector is smaller than size_outbox(), i.e., not enough entries provided to fill the outbox.");
        gpu::vector<real> output(this->size_inbox());
        backward(input.data(), output.data(), scaling);
        return output;
    }
    //! std::vector variants: HOST data, staged through device vectors
    template<typename real> std::vector<std::complex<real>> forward(std::vector<real> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_inbox()) throw std::invalid_argument("The input vector is smaller than size_inbox(), i.e., not enough entries provided to fill the inbox.");
        return gpu::transfer::unload(this->stream(), forward(gpu::transfer::load(this->stream(), input), scaling));
    }
    template<typename real> std::vector<real> backward(std::vector<std::complex<real>> const &input, scale scaling = scale::none) const {
        if (input.size() < this->size_outbox()) throw std::invalid_argument("The input vector is smaller than size_outbox(), i.e., not enough entries provided to fill the outbox.");
        return gpu::transfer::unload(this->stream(), backward(gpu::transfer::load(this->stream(), input), scaling));
    }
private:
    static int checked(int r2c_direction){
        if (r2c_direction < 0 or r2c_direction > 2) throw std::runtime_error("fft3d_r2c: r2c_direction must be 0, 1 or 2");
        return r2c_direction;
    }
};

// ---- aliases and factories of include/heffte_fft3d.h:703-763, include/heffte_fft3d_r2c.h:383-420 -----------------------------
//! two-dimensional transforms are three-dimensional plans on boxes of extent 1 along the third axis
template<typename backend_tag, typename index = int> using fft2d = fft3d<backend_tag, index>;
template<typename backend_tag, typename index = int> using fft2d_r2c = fft3d_r2c<backend_tag, index>;
//! real-to-real transforms: the cosine / sine tags
template<typename backend_tag, typename index = int> using rtransform = fft3d<backend_tag, index>;
template<typename backend_tag, typename index>
fft3d<backend_tag, index> make_fft3d(box3d<index> const inbox, box3d<index> const outbox, comm const &c, plan_options const options = default_options<backend_tag>()){
    static_assert(backend::is_enabled<backend_tag>::value, "the requested backend is not enabled");
    return fft3d<backend_tag, index>(inbox, outbox, c, options);
}
template<typename backend_tag, typename index>
fft3d_r2c<backend_tag, index> make_fft3d_r2c(box3d<index> const inbox, box3d<index> const outbox, int r2c_direction, comm const &c,
                                             plan_options const options = default_options<backend_tag>()){
    static_assert(backend::is_enabled<backend_tag>::value, "the requested backend is not enabled");
    return fft3d_r2c<backend_tag, index>(inbox, outbox, r2c_direction, c, options);
}

//! include/heffte_geometry.h:643-691 / 409-436 through the library (same answers as the reference, tests/test_plan_logic.py)
inline std::array<int, 3> proc_setup_min_surface(box3d<> const &world, int num_procs){
    int const nine[9] = {world.low[0], world.low[1], world.low[2], world.high[0], world.high[1], world.high[2], world.order[0], world.order[1], world.order[2]};
    std::array<int, 3> grid{};
    heffte_b200_proc_setup_min_surface(nine, num_procs, grid.data());
    return grid;
}
inline std::vector<box3d<>> split_world(box3d<> const &world, std::array<int, 3> const &proc_grid){
    int const nine[9] = {world.low[0], world.low[1], world.low[2], world.high[0], world.high[1], world.high[2], world.order[0], world.order[1], world.order[2]};
    size_t const n = static_cast<size_t>(proc_grid[0]) * proc_grid[1] * proc_grid[2];
    std::vector<int> raw(9 * n);
    heffte_b200_split_world(nine, proc_grid.data(), raw.data());
    std::vector<box3d<>> out;
    for(size_t i=0; i<n; i++){
        int const *b = raw.data() + 9 * i;
        out.push_back(box3d<>({b[0], b[1], b[2]}, {b[3], b[4], b[5]}, {b[6], b[7], b[8]}));
    }
    return out;
}

} // namespace heffte_b200

/*
 * Stand-alone use: the classes answer to the reference's spelling, heffte::fft3d<heffte::backend::b200>.
 * Next to the reference's own heffte.h (include it FIRST) the alias is left out and this front-end stays in its own
 * namespace: heffte_b200::fft3d<heffte_b200::backend::b200> is the plan-level (fused reshapes over NVLink) entry point,
 * while heffte::fft3d<heffte::backend::b200> is the reference's template over include/heffte_backend_b200.h.
 */
#if !defined(HEFFTE_H) && !defined(HEFFTE_COMMON_H) && !defined(HEFFTE_B200_NO_ALIAS)
namespace heffte = heffte_b200;
#endif

#endif
