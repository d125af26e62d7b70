// One slice of the power-of-two FFT kernel instantiations (real-data contig kernel, float, fused-reshape store); see fft_inst_real.inc.
#define B200_INST_NAME run_real_f32_scatter
#define B200_INST_DISPATCH dispatch_contig_real
#define B200_INST_TYPE float
#define B200_INST_SCATTER true
#include "fft_inst_real.inc"
