// One slice of the power-of-two FFT kernel instantiations (contig kernel, float, plain strided store); see fft_inst.inc.
#define B200_INST_NAME run_contig_f32_direct
#define B200_INST_DISPATCH dispatch_contig
#define B200_INST_TYPE float
#define B200_INST_SCATTER false
#include "fft_inst.inc"
