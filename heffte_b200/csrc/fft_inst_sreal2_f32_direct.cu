// One slice of the second-generation strided real kernel instantiations (float); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_sreal2_f32(int kind, int n, fft_args const &a, cuda_launcher &L){ return dispatch_strided_real2<float>(kind, n, a, L); }
int run_creal2_f32(int kind, int n, fft_args const &a, cuda_launcher &L){ return dispatch_contig_real2<float>(kind, n, a, L); }
}
