// One slice of the fused spectral-operator kernel instantiations (double, fused reshape store); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_conv_f64_scatter(int n, fft_args const &a, cuda_launcher &L){ return dispatch_strided_conv<double, true>(n, a, L); }
}
