// Host construction of scatter maps (fft_device.cuh): given my box, the axis the transform runs along and the boxes every
// rank owns after the following reshape, cut my box into cells that each land in ONE destination box and record where.
// This is the plan-time half of the fused reshape; it carries the same information as the reference's overlap maps
// (compute_overlap_map_direct_pack / _transpose_pack, src/heffte_reshape3d.cpp:125-206) -- who gets which sub-box at
// which offset with which strides -- but resolved per element on the device instead of per message on the host.
// Pure host code, no CUDA calls: unit-tested on the CPU (tests/test_scatter.py).
#pragma once

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "scatter_map.h"
#include "geometry.h"

namespace b200 {

// stride (in elements) of dimension id `dim` inside the memory layout of `box`
inline idx box_stride(box3 const &box, int dim){
    int const pos = box.position_of(dim);
    return (pos == 0) ? 1 : ((pos == 1) ? box.osize(0) : box.osize(0) * box.osize(1));
}

// k_pos: position (0 fast, 1 mid, 2 slow) inside `mine` of the axis the device code calls k (the transform axis, or 0 for a
// plain copy); a is the faster of the two remaining positions, b the other.
// dest[r] / dest_base[r]: box of rank r after the reshape and the address (valid on THIS device) of its first element.
// Returns false with a reason when the box cannot be expressed (too many cells, uncovered region).
// owners (optional, scatter_max_cells entries): the rank every cell lands on
// self_rank (optional): the rank that writes; the cells it keeps are flagged in map.local_mask
inline bool build_scatter_map(box3 const &mine, int k_pos, std::vector<box3> const &dest, std::vector<void*> const &dest_base,
                              int elem_bytes, scatter_map &map, std::string &why, int *owners = nullptr, int self_rank = -1){
    map = scatter_map{};
    map.nk = map.na = map.nb = 1;
    if (mine.empty()){ map.ncells = 0; return true; }
    int const a_pos = (k_pos == 0) ? 1 : 0;
    int const b_pos = (k_pos == 2) ? 1 : 2;
    int const pos_of[3] = {k_pos, a_pos, b_pos};
    int dim_of[3];
    for(int i=0; i<3; i++) dim_of[i] = mine.order[pos_of[i]];

    // cut points of every axis, in local coordinates
    std::vector<idx> cuts[3];
    std::vector<int> touching;
    for(size_t r=0; r<dest.size(); r++){
        box3 ov = mine.overlap(dest[r]);
        if (ov.empty()) continue;
        touching.push_back(static_cast<int>(r));
        for(int i=0; i<3; i++){
            int const d = dim_of[i];
            cuts[i].push_back(ov.low[d] - mine.low[d]);
            cuts[i].push_back(ov.high[d] + 1 - mine.low[d]);
        }
    }
    int counts[3];
    for(int i=0; i<3; i++){
        std::sort(cuts[i].begin(), cuts[i].end());
        cuts[i].erase(std::unique(cuts[i].begin(), cuts[i].end()), cuts[i].end());
        idx const extent = mine.size(dim_of[i]);
        if (cuts[i].empty() or cuts[i].front() != 0 or cuts[i].back() != extent){ why = "the destination boxes do not cover my box"; return false; }
        counts[i] = static_cast<int>(cuts[i].size()) - 1;
        if (counts[i] > scatter_max_cuts){ why = "too many cells along one axis"; return false; }
        if (extent > 2147483647LL){ why = "box extent exceeds 32 bits"; return false; }
    }
    if (counts[0] * counts[1] * counts[2] > scatter_max_cells){ why = "too many cells"; return false; }
    map.nk = counts[0]; map.na = counts[1]; map.nb = counts[2];
    map.ncells = counts[0] * counts[1] * counts[2];
    for(int c=0; c<scatter_max_cuts; c++){
        map.cut_k[c] = (c < counts[0]) ? static_cast<int>(cuts[0][c]) : 2147483647;
        map.cut_a[c] = (c < counts[1]) ? static_cast<int>(cuts[1][c]) : 2147483647;
        map.cut_b[c] = (c < counts[2]) ? static_cast<int>(cuts[2][c]) : 2147483647;
    }

    for(int ck=0; ck<counts[0]; ck++) for(int ca=0; ca<counts[1]; ca++) for(int cb=0; cb<counts[2]; cb++){
        std::array<idx, 3> corner{{0, 0, 0}};   // global coordinates of the low corner of the cell
        corner[dim_of[0]] = mine.low[dim_of[0]] + cuts[0][ck];
        corner[dim_of[1]] = mine.low[dim_of[1]] + cuts[1][ca];
        corner[dim_of[2]] = mine.low[dim_of[2]] + cuts[2][cb];
        int owner = -1;
        for(int r : touching){
            box3 const &b = dest[r];
            bool inside = true;
            for(int d=0; d<3; d++) inside = inside and corner[d] >= b.low[d] and corner[d] <= b.high[d];
            if (inside){ owner = r; break; }
        }
        if (owner < 0){ why = "a cell of my box belongs to no destination box"; return false; }
        box3 const &b = dest[owner];
        scatter_cell &cell = map.cell[(ck * counts[1] + ca) * counts[2] + cb];
        if (owners) owners[(ck * counts[1] + ca) * counts[2] + cb] = owner;
        if (owner == self_rank) map.local_mask |= 1ULL << ((ck * counts[1] + ca) * counts[2] + cb);
        // destination element of local (k, a, b): sum_d (mine.low[d] + local_d - b.low[d]) * stride_b(d)
        idx shift = 0;
        for(int d=0; d<3; d++) shift += (mine.low[d] - b.low[d]) * box_stride(b, d);
        cell.base = static_cast<long long>(reinterpret_cast<intptr_t>(dest_base[owner])) + shift * elem_bytes;
        cell.sk = box_stride(b, dim_of[0]);
        cell.sa = box_stride(b, dim_of[1]);
        cell.sb = box_stride(b, dim_of[2]);
    }
    return true;
}

// true when every destination box written by `map` keeps the k axis of the source unit-stride or keeps the line axis
// unit-stride: used only for reporting (coalescing quality), never for correctness
inline bool scatter_is_coalesced(scatter_map const &map, bool lines_are_fast){
    for(int c=0; c<map.ncells; c++){
        if (lines_are_fast){ if (map.cell[c].sa != 1) return false; }
        else if (map.cell[c].sk != 1) return false;
    }
    return true;
}

} // namespace b200
