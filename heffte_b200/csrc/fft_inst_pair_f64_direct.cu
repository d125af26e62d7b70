// One slice of the paired-kernel instantiations (double, plain store); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_pair_f64_direct(int n, bool contig_first, pair_args const &p, cuda_launcher &L){ return dispatch_pair<double, false>(n, contig_first, p, L); }
}
