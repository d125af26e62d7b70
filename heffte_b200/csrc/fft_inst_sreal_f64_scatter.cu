// One slice of the power-of-two FFT kernel instantiations (real-data strided kernel, double, fused-reshape store); see fft_inst_real.inc.
#define B200_INST_NAME run_sreal_f64_scatter
#define B200_INST_DISPATCH dispatch_strided_real
#define B200_INST_TYPE double
#define B200_INST_SCATTER true
#include "fft_inst_real.inc"
