/*
 * heffte_backend_b200.h -- the b200 backend of heFFTe as a TEMPLATE PLUG-IN: the header a maintainer adds next to
 * include/heffte_backend_cuda.h so that the reference's own heffte::fft3d<backend::b200>, fft3d_r2c<backend::b200> and
 * fft3d<backend::b200_cos / b200_sin / b200_cos1> compile and run on the hand-written sm_100a kernels of
 * libheffte_b200.so instead of cuFFT and the kernels of src/heffte_backend_cuda.cu.
 *
 * It provides exactly the contract the reference's templates consume:
 *   backend tags + is_enabled / name / uses_fft_types / buffer_traits      (reference include/heffte_common.h:95-215, 225-264, 439-543;
 *                                                                             include/heffte_backend_cuda.h:179-326)
 *   backend::device_instance<tag::gpu>, default_backend<tag::gpu>,
 *   backend::data_manipulator<tag::gpu>                                     (include/heffte_backend_cuda.h:201-284)
 *   b200_executor / b200_executor_r2c / b200_executor_r2r<kind>             (executor_base, include/heffte_common.h:561-595;
 *                                                                             cufft_executor :436-572, cufft_executor_r2c :631-751;
 *                                                                             real2real_executor include/heffte_r2r_executor.h:191-278)
 *   one_dim_backend<tag>                                                    (include/heffte_backend_cuda.h:759-794)
 *   direct_packer<tag::gpu>, transpose_packer<tag::gpu>                     (:800-829)
 *   data_scaling::apply, default_plan_options<tag>                          (:831-877)
 *   gpu::device_count / device_set / synchronize_default_stream             (include/heffte_backend_vector.h:163-177; src/heffte_backend_cuda.cu:16-34)
 *
 * Everything is implemented over the C ABI of include/heffte_b200_kernels.h: no CUDA header is needed, a cudaStream_t
 * travels as void*.  Include order: after heffte_backend_vector.h, before heffte_backend_data_transfer.h (the position of
 * heffte_backend_cuda.h inside include/heffte_backends.h); enabled by Heffte_ENABLE_B200 (which implies Heffte_ENABLE_GPU).
 * INTEGRATION.md lists the reference-side edits; integration/ applies them to a scratch copy of the reference and builds the
 * reference's OWN test programs against this header (tests/test_z_reference_plugin_gpu.py runs them on the GPU).
 *
 * In this mode the reshapes are the reference's (pack -> MPI -> unpack through the packers below); the fused
 * reshape-in-the-store data plane over NVLink peer memory is reached through the plan-level entry points
 * (include/heffte_b200.h, heffte::b200::fft3d in include/heffte_b200.hpp), which a b200-aware heffte::fft3d would call.
 */
#ifndef HEFFTE_BACKEND_B200_H
#define HEFFTE_BACKEND_B200_H

#ifdef Heffte_ENABLE_B200

#include <array>
#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>

#include "heffte_b200_kernels.h"

//! \brief Forward declaration of the CUDA stream object, as in the CUDA headers (cudaStream_t = CUstream_st*): no CUDA header is needed.
struct CUstream_st;

namespace heffte {

/*! \brief Helpers of the b200 backend (the role of namespace heffte::cuda in the reference). */
namespace b200 {
    //! \brief Converts a b200_* return code into the exception the reference throws from cuda::check_error (heffte_backend_cuda.h:49-60).
    inline void check_error(int status, const char *function_name){
        if (status != B200_SUCCESS)
            throw std::runtime_error(std::string(function_name) + " failed with message: " + b200_last_error());
    }
    //! \brief The stream type: a cudaStream_t (the C ABI takes it as void*; the reference's transfer helpers overload on void* for the CPU).
    using stream_t = ::CUstream_st*;

    template<typename T> struct precision_of{};
    template<> struct precision_of<float>{ static constexpr int value = B200_PREC_FLOAT; };
    template<> struct precision_of<double>{ static constexpr int value = B200_PREC_DOUBLE; };
    template<> struct precision_of<std::complex<float>>{ static constexpr int value = B200_PREC_FLOAT; };
    template<> struct precision_of<std::complex<double>>{ static constexpr int value = B200_PREC_DOUBLE; };

    //! \brief real -> complex with zero imaginary part (reference cuda::convert, src/heffte_backend_cuda.cu:352-361).
    template<typename precision_type, typename index>
    void convert(stream_t stream, index num_entries, precision_type const source[], std::complex<precision_type> destination[]){
        check_error(b200_convert_r2c(precision_of<precision_type>::value, static_cast<long long>(num_entries), source, destination, stream), "b200_convert_r2c()");
    }
    //! \brief complex -> real, drops the imaginary part.
    template<typename precision_type, typename index>
    void convert(stream_t stream, index num_entries, std::complex<precision_type> const source[], precision_type destination[]){
        check_error(b200_convert_c2r(precision_of<precision_type>::value, static_cast<long long>(num_entries), source, destination, stream), "b200_convert_c2r()");
    }
}

namespace backend {
    //! \brief Type-tag of the b200 backend: hand-written sm_100a FFT kernels (complex-to-complex and real-to-complex).
    struct b200{};
    //! \brief Cosine transform (DCT-II forward, DCT-III backward) computed inside the b200 FFT kernels.
    struct b200_cos{};
    //! \brief Sine transform (DST-II forward, DST-III backward).
    struct b200_sin{};
    //! \brief Cosine transform of type I.
    struct b200_cos1{};

    template<> struct is_enabled<b200> : std::true_type{};
    template<> struct is_enabled<b200_cos> : std::true_type{};
    template<> struct is_enabled<b200_sin> : std::true_type{};
    template<> struct is_enabled<b200_cos1> : std::true_type{};

    template<> inline std::string name<b200>(){ return "b200"; }
    template<> inline std::string name<b200_cos>(){ return "b200-cos-type-II"; }
    template<> inline std::string name<b200_sin>(){ return "b200-sin-type-II"; }
    template<> inline std::string name<b200_cos1>(){ return "b200-cos-type-I"; }

    template<> struct uses_fft_types<b200_cos> : std::false_type{};
    template<> struct uses_fft_types<b200_sin> : std::false_type{};
    template<> struct uses_fft_types<b200_cos1> : std::false_type{};

    /*! \brief The stream holder (reference heffte_backend_cuda.h:201-215). */
    template<>
    struct device_instance<tag::gpu>{
        device_instance(heffte::b200::stream_t new_stream = nullptr) : _stream(new_stream){}
        heffte::b200::stream_t stream(){ return _stream; }
        heffte::b200::stream_t stream() const{ return _stream; }
        void synchronize_device() const{ heffte::b200::check_error(b200_stream_synchronize(_stream), "device sync"); }
        mutable heffte::b200::stream_t _stream;
        using stream_type = heffte::b200::stream_t;
    };

    template<> struct default_backend<tag::gpu>{ using type = b200; };

    /*! \brief Device memory and copies (reference heffte_backend_cuda.h:230-284). */
    template<> struct data_manipulator<tag::gpu>{
        using stream_type = heffte::b200::stream_t;
        using backend_device = backend::device_instance<tag::gpu>;
        template<typename scalar_type>
        static scalar_type* allocate(stream_type, size_t num_entries){
            void *new_data = nullptr;
            heffte::b200::check_error(b200_device_alloc(num_entries * sizeof(scalar_type), &new_data), "b200_device_alloc()");
            return reinterpret_cast<scalar_type*>(new_data);
        }
        template<typename scalar_type>
        static void free(stream_type, scalar_type *pntr){
            if (pntr == nullptr) return;
            heffte::b200::check_error(b200_device_free(pntr), "b200_device_free()");
        }
        template<typename scalar_type>
        static void copy_n(stream_type stream, scalar_type const source[], size_t num_entries, scalar_type destination[]){
            heffte::b200::check_error(b200_copy_on_device(source, destination, num_entries * sizeof(scalar_type), stream), "data_manipulator::copy_n()");
            if (stream == nullptr) heffte::b200::check_error(b200_stream_synchronize(nullptr), "data_manipulator::copy_n()");
        }
        template<typename scalar_type>
        static void copy_n(stream_type stream, std::complex<scalar_type> const source[], size_t num_entries, scalar_type destination[]){
            heffte::b200::convert(stream, static_cast<long long>(num_entries), source, destination);
        }
        template<typename scalar_type>
        static void copy_n(stream_type stream, scalar_type const source[], size_t num_entries, std::complex<scalar_type> destination[]){
            heffte::b200::convert(stream, static_cast<long long>(num_entries), source, destination);
        }
        template<typename scalar_type>
        static void copy_device_to_host(stream_type stream, scalar_type const source[], size_t num_entries, scalar_type destination[]){
            heffte::b200::check_error(b200_copy_to_host(source, destination, num_entries * sizeof(scalar_type), stream), "device_to_host (b200)");
        }
        template<typename scalar_type>
        static void copy_device_to_device(stream_type stream, scalar_type const source[], size_t num_entries, scalar_type destination[]){
            heffte::b200::check_error(b200_copy_on_device(source, destination, num_entries * sizeof(scalar_type), stream), "device_to_device (b200)");
        }
        template<typename scalar_type>
        static void copy_host_to_device(stream_type stream, scalar_type const source[], size_t num_entries, scalar_type destination[]){
            heffte::b200::check_error(b200_copy_to_device(source, destination, num_entries * sizeof(scalar_type), stream), "host_to_device (b200)");
        }
    };

    #define HEFFTE_B200_BUFFER_TRAITS(tag_name) \
    template<> struct buffer_traits<tag_name>{ \
        using location = tag::gpu; \
        template<typename T> using container = heffte::gpu::device_vector<T, data_manipulator<tag::gpu>>; \
    };
    HEFFTE_B200_BUFFER_TRAITS(b200)
    HEFFTE_B200_BUFFER_TRAITS(b200_cos)
    HEFFTE_B200_BUFFER_TRAITS(b200_sin)
    HEFFTE_B200_BUFFER_TRAITS(b200_cos1)
    #undef HEFFTE_B200_BUFFER_TRAITS
}

namespace b200 {
    /*!
     * \brief Geometry of the batch of lines of a box that run along a dimension: the two other axes
     * are kept apart so that the middle-axis transform is ONE launch (the reference loops `blocks` cuFFT calls,
     * heffte_backend_cuda.h:452, 496-499).
     */
    template<typename index>
    void line_layout(box3d<index> const &box, int dimension, b200_line_geom &geom, long long &count_a, long long &count_b){
        long long const strides[3] = {1, static_cast<long long>(box.osize(0)), static_cast<long long>(box.osize(0)) * static_cast<long long>(box.osize(1))};
        int const pos = box.find_order(dimension);
        int const a_pos = (pos == 0) ? 1 : 0, b_pos = (pos == 2) ? 1 : 2;
        geom.stride = strides[pos]; geom.stride_a = strides[a_pos]; geom.stride_b = strides[b_pos];
        count_a = box.osize(a_pos); count_b = box.osize(b_pos);
    }
    //! \brief Owning handle of a batched 1-D plan of libheffte_b200.so.
    struct plan_deleter{ void operator()(b200_fft1d_plan_s *p) const{ b200_fft1d_destroy(p); } };
    using plan_pointer = std::unique_ptr<b200_fft1d_plan_s, plan_deleter>;

    inline plan_pointer make_plan(b200_fft1d_desc const &desc){
        b200_fft1d_plan raw = nullptr;
        check_error(b200_fft1d_create(&desc, &raw), "b200_fft1d_create()");
        return plan_pointer(raw);
    }
    //! \brief Plan for the lines of `box` along `dimension`; `cbox` is the shortened complex box of an r2c transform.
    template<typename index>
    plan_pointer make_plan(int precision, int kind, box3d<index> const &box, int dimension, box3d<index> const *cbox = nullptr){
        b200_fft1d_desc d{};
        d.precision = precision; d.kind = kind; d.n = box.size[dimension];
        line_layout(box, dimension, d.in, d.count_a, d.count_b);
        d.out = d.in;
        if (cbox != nullptr){ long long ca, cb; line_layout(*cbox, dimension, d.out, ca, cb); }
        return make_plan(d);
    }
}

/*!
 * \brief Executor of the b200 backend: batched 1-D transforms of a box along one, two or all three dimensions, in place.
 *
 * Stands where heffte::cufft_executor stands (heffte_backend_cuda.h:436-572).  `kind` selects complex-to-complex
 * (B200_C2C) or one of the real-to-real transforms (B200_COS / B200_SIN / B200_COS1), which run INSIDE the FFT kernel
 * (Makhoul's N-point algorithm) instead of the reference's pre/post-processing kernels around a 4N-point r2c FFT
 * (heffte_r2r_executor.h:191-278): no workspace, strided lines allowed.  Plans are created on first use per precision and
 * are `mutable` like the reference's lazy cuFFT plans (:555-571).
 */
template<int kind>
class b200_executor_kind : public executor_base{
public:
    using executor_base::forward;
    using executor_base::backward;
    using executor_base::complex_size;
    //! \brief One dimension.
    template<typename index>
    b200_executor_kind(b200::stream_t active_stream, box3d<index> const box, int dimension) :
        stream(active_stream), total_size(static_cast<int>(box.count())), num_passes(1){
        setup(0, box, dimension);
    }
    //! \brief Two dimensions (slab decomposition): two passes, the faster dimension first.
    template<typename index>
    b200_executor_kind(b200::stream_t active_stream, box3d<index> const box, int dir1, int dir2) :
        stream(active_stream), total_size(static_cast<int>(box.count())), num_passes(2){
        bool const first = box.find_order(dir1) < box.find_order(dir2);
        setup(0, box, first ? dir1 : dir2);
        setup(1, box, first ? dir2 : dir1);
    }
    //! \brief All three dimensions (single rank).
    template<typename index>
    b200_executor_kind(b200::stream_t active_stream, box3d<index> const box) :
        stream(active_stream), total_size(static_cast<int>(box.count())), num_passes(3){
        for(int i=0; i<3; i++) setup(i, box, box.order[i]);
    }

    void forward(std::complex<float> data[], std::complex<float>*) const override{ run(B200_PREC_FLOAT, B200_FORWARD, data); }
    void backward(std::complex<float> data[], std::complex<float>*) const override{ run(B200_PREC_FLOAT, B200_BACKWARD, data); }
    void forward(std::complex<double> data[], std::complex<double>*) const override{ run(B200_PREC_DOUBLE, B200_FORWARD, data); }
    void backward(std::complex<double> data[], std::complex<double>*) const override{ run(B200_PREC_DOUBLE, B200_BACKWARD, data); }
    void forward(float data[], float*) const override{ run(B200_PREC_FLOAT, B200_FORWARD, data); }
    void backward(float data[], float*) const override{ run(B200_PREC_FLOAT, B200_BACKWARD, data); }
    void forward(double data[], double*) const override{ run(B200_PREC_DOUBLE, B200_FORWARD, data); }
    void backward(double data[], double*) const override{ run(B200_PREC_DOUBLE, B200_BACKWARD, data); }

    //! \brief Real input of a complex plan: convert, then transform (reference :526-545).
    void forward(float const indata[], std::complex<float> outdata[], std::complex<float> *workspace) const override{
        b200::convert(stream, total_size, indata, outdata);
        forward(outdata, workspace);
    }
    void backward(std::complex<float> indata[], float outdata[], std::complex<float> *workspace) const override{
        backward(indata, workspace);
        b200::convert(stream, total_size, indata, outdata);
    }
    void forward(double const indata[], std::complex<double> outdata[], std::complex<double> *workspace) const override{
        b200::convert(stream, total_size, indata, outdata);
        forward(outdata, workspace);
    }
    void backward(std::complex<double> indata[], double outdata[], std::complex<double> *workspace) const override{
        backward(indata, workspace);
        b200::convert(stream, total_size, indata, outdata);
    }

    int box_size() const override{ return total_size; }
    size_t workspace_size() const override{ return 0; }

private:
    template<typename index>
    void setup(int pass, box3d<index> const &box, int dimension){
        b200_fft1d_desc &d = desc[pass];
        d = b200_fft1d_desc{};
        d.kind = kind; d.n = box.size[dimension];
        b200::line_layout(box, dimension, d.in, d.count_a, d.count_b);
        d.out = d.in;
    }
    void run(int precision, int direction, void *data) const{
        for(int i=0; i<num_passes; i++){
            int const pass = (direction == B200_FORWARD) ? i : num_passes - 1 - i;
            b200::plan_pointer &plan = plans[precision][pass];
            if (not plan){
                b200_fft1d_desc d = desc[pass];
                d.precision = precision;
                plan = b200::make_plan(d);
            }
            b200::check_error(b200_fft1d_execute(plan.get(), direction, data, data, 1.0, stream), "b200_fft1d_execute()");
        }
    }

    mutable b200::stream_t stream;
    int total_size, num_passes;
    b200_fft1d_desc desc[3];
    mutable b200::plan_pointer plans[2][3];
};

//! \brief The complex-to-complex executor of backend::b200.
using b200_executor = b200_executor_kind<B200_C2C>;

/*!
 * \brief Real-to-complex executor with shortening of the data (stands where heffte::cufft_executor_r2c stands, :631-751).
 *
 * The real box has box_size() entries, the complex result sits in box.r2c(dimension) with complex_size() entries; the two
 * arrays may not alias.  No realignment copies: lines that are not aligned to a complex number take the generic kernel.
 */
class b200_executor_r2c : public executor_base{
public:
    using executor_base::forward;
    using executor_base::backward;
    template<typename index>
    b200_executor_r2c(b200::stream_t active_stream, box3d<index> const box, int dimension) :
        stream(active_stream), rsize(static_cast<int>(box.count())), csize(static_cast<int>(box.r2c(dimension).count())){
        desc = b200_fft1d_desc{};
        desc.kind = B200_R2C; desc.n = box.size[dimension];
        b200::line_layout(box, dimension, desc.in, desc.count_a, desc.count_b);
        box3d<index> const cbox = box.r2c(dimension);
        long long ca, cb;
        b200::line_layout(cbox, dimension, desc.out, ca, cb);
    }
    void forward(float const indata[], std::complex<float> outdata[], std::complex<float>*) const override{ run(B200_PREC_FLOAT, B200_FORWARD, indata, outdata); }
    void backward(std::complex<float> indata[], float outdata[], std::complex<float>*) const override{ run(B200_PREC_FLOAT, B200_BACKWARD, indata, outdata); }
    void forward(double const indata[], std::complex<double> outdata[], std::complex<double>*) const override{ run(B200_PREC_DOUBLE, B200_FORWARD, indata, outdata); }
    void backward(std::complex<double> indata[], double outdata[], std::complex<double>*) const override{ run(B200_PREC_DOUBLE, B200_BACKWARD, indata, outdata); }
    int box_size() const override{ return rsize; }
    int complex_size() const override{ return csize; }
    size_t workspace_size() const override{ return 0; }
private:
    void run(int precision, int direction, const void *in, void *out) const{
        b200::plan_pointer &plan = plans[precision];
        if (not plan){
            b200_fft1d_desc d = desc;
            d.precision = precision;
            plan = b200::make_plan(d);
        }
        b200::check_error(b200_fft1d_execute(plan.get(), direction, in, out, 1.0, stream), "b200_fft1d_execute() r2c");
    }
    mutable b200::stream_t stream;
    int rsize, csize;
    b200_fft1d_desc desc;
    mutable b200::plan_pointer plans[2];
};

template<> struct one_dim_backend<backend::b200>{
    using executor = b200_executor;
    using executor_r2c = b200_executor_r2c;
};
template<> struct one_dim_backend<backend::b200_cos>{
    using executor = b200_executor_kind<B200_COS>;
    using executor_r2c = void;
};
template<> struct one_dim_backend<backend::b200_sin>{
    using executor = b200_executor_kind<B200_SIN>;
    using executor_r2c = void;
};
template<> struct one_dim_backend<backend::b200_cos1>{
    using executor = b200_executor_kind<B200_COS1>;
    using executor_r2c = void;
};

/*! \brief Sub-box copy between a strided box and a dense buffer (reference :800-811). */
template<> struct direct_packer<tag::gpu>{
    template<typename scalar_type, typename index>
    void pack(b200::stream_t stream, pack_plan_3d<index> const &plan, scalar_type const data[], scalar_type buffer[]) const{
        b200::check_error(b200_direct_pack(static_cast<int>(sizeof(scalar_type)), plan.size[0], plan.size[1], plan.size[2], plan.line_stride, plan.plane_stride,
                                           data, buffer, stream), "b200_direct_pack()");
    }
    template<typename scalar_type, typename index>
    void unpack(b200::stream_t stream, pack_plan_3d<index> const &plan, scalar_type const buffer[], scalar_type data[]) const{
        b200::check_error(b200_direct_unpack(static_cast<int>(sizeof(scalar_type)), plan.size[0], plan.size[1], plan.size[2], plan.line_stride, plan.plane_stride,
                                             buffer, data, stream), "b200_direct_unpack()");
    }
};

/*! \brief Unpack with axis permutation through a shared-memory tile (reference :817-829). */
template<> struct transpose_packer<tag::gpu>{
    template<typename scalar_type, typename index>
    void pack(b200::stream_t stream, pack_plan_3d<index> const &plan, scalar_type const data[], scalar_type buffer[]) const{
        direct_packer<tag::gpu>().pack(stream, plan, data, buffer);
    }
    template<typename scalar_type, typename index>
    void unpack(b200::stream_t stream, pack_plan_3d<index> const &plan, scalar_type const buffer[], scalar_type data[]) const{
        b200::check_error(b200_transpose_unpack(static_cast<int>(sizeof(scalar_type)), plan.size[0], plan.size[1], plan.size[2], plan.line_stride, plan.plane_stride,
                                                plan.buff_line_stride, plan.buff_plane_stride, plan.map[0], plan.map[1], plan.map[2], buffer, data, stream),
                          "b200_transpose_unpack()");
    }
};

namespace data_scaling {
    /*! \brief data[i] *= scale_factor over num_entries reals (reference :831-848). */
    template<typename scalar_type, typename index>
    void apply(b200::stream_t stream, index num_entries, scalar_type *data, double scale_factor){
        b200::check_error(b200_scale(b200::precision_of<scalar_type>::value, static_cast<long long>(num_entries), data, scale_factor, stream), "b200_scale()");
    }
    template<typename precision_type, typename index>
    void apply(b200::stream_t stream, index num_entries, std::complex<precision_type> *data, double scale_factor){
        apply<precision_type>(stream, 2 * num_entries, reinterpret_cast<precision_type*>(data), scale_factor);
    }
}

/*!
 * \brief Default options (reference :854-877).  The strided kernels run at the HBM rate of the contiguous ones, so no
 * backend of this family needs the reorder -- not even the cosine / sine ones, for which the reference forces it
 * (include/heffte_plan_logic.h:206-224).
 */
template<> struct default_plan_options<backend::b200>{ static const bool use_reorder = false; };
template<> struct default_plan_options<backend::b200_cos>{ static const bool use_reorder = false; };
template<> struct default_plan_options<backend::b200_sin>{ static const bool use_reorder = false; };
template<> struct default_plan_options<backend::b200_cos1>{ static const bool use_reorder = false; };

namespace gpu {
    // declared in include/heffte_backend_vector.h:163-177, defined by the backend's source file in the reference
    // (src/heffte_backend_cuda.cu:16-34); header-only here
    inline int device_count(){ return b200_device_count(); }
    inline void device_set(int active_device){ heffte::b200::check_error(b200_device_set(active_device), "b200_device_set()"); }
    inline void synchronize_default_stream(){ heffte::b200::check_error(b200_stream_synchronize(nullptr), "b200_stream_synchronize()"); }
}

}

#endif   // Heffte_ENABLE_B200

#endif   /* HEFFTE_BACKEND_B200_H */
