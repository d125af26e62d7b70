// Plain-data description of a fused reshape, shared by the host builder (scatter_build.h) and the kernels (fft_device.cuh).
#pragma once

namespace b200 {

struct scatter_cell {
    long long base;          // address of the destination element of local (k, a, b) = (0, 0, 0), as an integer
    long long sk, sa, sb;    // strides of the destination box, in elements of the destination type
};
constexpr int scatter_max_cuts = 8;
constexpr int scatter_max_cells = 64;
struct scatter_map {
    int nk, na, nb, ncells;
    int cut_k[scatter_max_cuts], cut_a[scatter_max_cuts], cut_b[scatter_max_cuts];   // first index of every cell (cut[0] = 0)
    unsigned long long local_mask;   // bit c: cell c lands in the memory of the rank that writes it (may be re-based per call)
    unsigned long long reserved;
    scatter_cell cell[scatter_max_cells];                                            // index (ck * na + ca) * nb + cb
};
constexpr int scatter_header_bytes = 16 + 3 * 4 * scatter_max_cuts + 16;   // 128: multiple of 16


} // namespace b200
