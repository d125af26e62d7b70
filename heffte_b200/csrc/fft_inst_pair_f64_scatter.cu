// One slice of the paired-kernel instantiations (double, fused reshape store); see fft_dispatch.cuh.
#include "fft_dispatch.cuh"
#include "runtime.h"

namespace b200 {
int run_pair_f64_scatter(int n, bool contig_first, pair_args const &p, cuda_launcher &L){ return dispatch_pair<double, true>(n, contig_first, p, L); }
}
