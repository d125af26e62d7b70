"""
Known-answer vectors that the reference's own test-suite asserts for the hot path (numeric facts transcribed from
icl-utk-edu/heffte v2.4.1; the file:line of each is given).  They pin the oracle (tests/test_oracle.py) and the CUDA path.
"""
import numpy as np

# test/test_units_nompi.cpp:306-331 -- 1-D DCT-II / DST-II (unnormalised, REDFT10 / RODFT10) of small vectors
DCT2_1234 = np.array([20.0, -6.3086440598, 0.0, -0.4483415292])
DST2_1234 = np.array([13.0656296488, -5.6568542495, 5.4119610015, -4.0])

# test/test_cos.cpp:39-54 -- 3-D transforms of the 2x3x4 box filled with 1..24 (forward, no scaling)
COS_2x3x4 = np.array([2.4e+03, -6.7882250993908571e+01, -2.2170250336881628e+02, 0.0, 0.0, 0.0, -9.0844474461089760e+02, 0.0, 0.0, 0.0,
                      0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -6.4561180200187039e+01, 0.0, 0.0, 0.0, 0.0, 0.0])
SIN_2x3x4 = np.array([7.3910362600902943e+02, -4.1810014876044050e+01, -1.0241320258448191e+02, 0.0, 3.6955181300451477e+02,
                      -2.0905007438022025e+01, -3.8400000000000006e+02, 0.0, 0.0, 0.0, -1.9200000000000003e+02, 0.0,
                      3.0614674589207186e+02, -1.7318275204678301e+01, -4.2420937476555700e+01, 0.0, 1.5307337294603599e+02,
                      -8.6591376023391504e+00, -2.7152900397563417e+02, 0.0, 0.0, 0.0, -1.3576450198781720e+02, 0.0])
COS1_2x3x4 = np.array([600.0, -24.0, -48.0, 0.0, 0.0, 0.0, -192.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -48.0, 0.0,
                       0.0, 0.0, 0.0, 0.0])

# test/test_c.c:27-74, 118-133 -- 4x4x4 world on two ranks split along the slow dimension (rank 0: k = 0..1, rank 1: k = 2..3);
# EACH rank fills its 32 local entries with 0..31.  Non-zero entries of the c2c spectrum per rank (local index -> value).
C_TEST_4x4x4_RANK0 = {0: 992.0 + 0j, 1: -32.0 + 32.0j, 2: -32.0 + 0j, 3: -32.0 - 32.0j, 4: -128.0 + 128.0j, 8: -128.0 + 0j, 12: -128.0 - 128.0j}
C_TEST_4x4x4_RANK1 = {0: -512.0 + 0j}
# test/test_c.c:149-151, 225-227 -- sizes reported by the 2-rank plans of that test
C_TEST_SIZES = {"c2c": [dict(inbox=32, outbox=32, workspace=96), dict(inbox=32, outbox=32, workspace=96)],
                "r2c": [dict(inbox=32, outbox=32, workspace=96), dict(inbox=32, outbox=16, workspace=88)]}  # per rank, r2c_direction = 2

# test/test_units_nompi.cpp:595-636 -- local transposes of the 2x3x4 box holding 1..24 to orders (1,2,0), (2,1,0), (0,2,1)
TRANSPOSE_2x3x4 = {
    (1, 2, 0): [1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24],
    (2, 1, 0): [1, 7, 13, 19, 3, 9, 15, 21, 5, 11, 17, 23, 2, 8, 14, 20, 4, 10, 16, 22, 6, 12, 18, 24],
    (0, 2, 1): [1, 2, 7, 8, 13, 14, 19, 20, 3, 4, 9, 10, 15, 16, 21, 22, 5, 6, 11, 12, 17, 18, 23, 24],
}

# test/test_units_nompi.cpp:25-46 -- processor grids asserted by the reference
PROCGRID = {20: [4, 5], 17: [1, 17], 6561: [81, 81], 323: [17, 19], 128: [8, 16]}
