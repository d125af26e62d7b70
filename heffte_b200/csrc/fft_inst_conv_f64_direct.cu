// One slice of the fused spectral-operator kernel instantiations (double, plain store); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_conv_f64_direct(int n, fft_args const &a, cuda_launcher &L){ return dispatch_strided_conv<double, false>(n, a, L); }
}
