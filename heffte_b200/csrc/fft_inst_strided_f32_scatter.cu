// One slice of the power-of-two FFT kernel instantiations (strided kernel, float, fused-reshape (scatter) store); see fft_inst.inc.
#define B200_INST_NAME run_strided_f32_scatter
#define B200_INST_DISPATCH dispatch_strided
#define B200_INST_TYPE float
#define B200_INST_SCATTER true
#include "fft_inst.inc"
