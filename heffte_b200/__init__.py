"""
heffte_b200 -- B200-native implementation of heFFTe's distributed 3-D FFT hot path.
The compute path is libheffte_b200.so (hand-written sm_100a CUDA + NCCL); importing this package never falls back
to a CPU implementation.
"""
from . import heffte  # noqa: F401
from .heffte import (backend, scale, reshape_algorithm, box3d, plan_options, fft3d, fft3d_r2c,  # noqa: F401
                     comm_self, comm_threads, comm_from_torch, comm_from_callbacks, heffte_input_error)
