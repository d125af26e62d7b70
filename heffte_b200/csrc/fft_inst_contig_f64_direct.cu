// One slice of the power-of-two FFT kernel instantiations (contig kernel, double, plain strided store); see fft_inst.inc.
#define B200_INST_NAME run_contig_f64_direct
#define B200_INST_DISPATCH dispatch_contig
#define B200_INST_TYPE double
#define B200_INST_SCATTER false
#include "fft_inst.inc"
