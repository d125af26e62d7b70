// One slice of the fused spectral-operator kernel instantiations (float, plain store); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_conv_f32_direct(int n, fft_args const &a, cuda_launcher &L){ return dispatch_strided_conv<float, false>(n, a, L); }
}
