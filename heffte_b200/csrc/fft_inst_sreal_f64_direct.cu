// One slice of the power-of-two FFT kernel instantiations (real-data strided kernel, double, plain store); see fft_inst_real.inc.
#define B200_INST_NAME run_sreal_f64_direct
#define B200_INST_DISPATCH dispatch_strided_real
#define B200_INST_TYPE double
#define B200_INST_SCATTER false
#include "fft_inst_real.inc"
