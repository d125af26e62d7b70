// One slice of the power-of-two FFT kernel instantiations (strided kernel, float, plain strided store); see fft_inst.inc.
#define B200_INST_NAME run_strided_f32_direct
#define B200_INST_DISPATCH dispatch_strided
#define B200_INST_TYPE float
#define B200_INST_SCATTER false
#include "fft_inst.inc"
