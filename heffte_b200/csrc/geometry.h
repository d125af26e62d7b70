// Integer geometry of the distributed transform: boxes, processor grids, world splitting.
// Behaviour (not code) follows the reference so that every rank ends up with exactly the same intermediate
// boxes as icl-utk-edu/heffte would give it: include/heffte_geometry.h:67-133 (box3d), :302-349 (factor pairs, 2-D
// grid), :362-396 (constrained grid), :409-436 (split_world), :489-524 (maximize_overlap), :568-629 (pencils/slabs),
// :643-691 (min-surface grid).  Parity is checked box-by-box against the reference in tests/test_plan_logic.py.
#pragma once

#include <algorithm>
#include <array>
#include <limits>
#include <stdexcept>
#include <vector>

namespace b200 {

using idx = long long;

struct box3 {
    std::array<idx, 3> low{{0, 0, 0}}, high{{-1, -1, -1}};
    std::array<int, 3> order{{0, 1, 2}};

    box3() = default;
    box3(std::array<idx, 3> l, std::array<idx, 3> h, std::array<int, 3> o = {{0, 1, 2}}) : low(l), high(h), order(o) {}

    idx size(int d) const { return high[d] - low[d] + 1; }
    idx osize(int d) const { return size(order[d]); }          // 0 fast, 1 mid, 2 slow
    bool empty() const { return size(0) <= 0 or size(1) <= 0 or size(2) <= 0; }
    idx count() const { return empty() ? 0 : size(0) * size(1) * size(2); }
    bool same_extent(box3 const &o) const { return low == o.low and high == o.high; }
    bool same_order(box3 const &o) const { return order == o.order; }
    int position_of(int dim) const { return (order[0] == dim) ? 0 : ((order[1] == dim) ? 1 : 2); }
    bool is2d() const { return size(0) == 1 or size(1) == 1 or size(2) == 1; }

    static box3 nothing(std::array<int, 3> o = {{0, 1, 2}}){ return box3({{0, 0, 0}}, {{-1, -1, -1}}, o); }

    // intersection, keeps the order of *this
    box3 overlap(box3 const &o) const {
        if (empty() or o.empty()) return nothing();
        box3 r = *this;
        for(int d=0; d<3; d++){ r.low[d] = std::max(low[d], o.low[d]); r.high[d] = std::min(high[d], o.high[d]); }
        return r;
    }
    // half-spectrum box of a real-to-complex transform along `dim`
    box3 halved(int dim) const {
        if (empty()) return nothing();
        box3 r = *this;
        r.high[dim] = low[dim] + size(dim) / 2;
        return r;
    }
    box3 reordered(std::array<int, 3> o) const { box3 r = *this; r.order = o; return r; }
    // linear position of a global index inside this box
    idx offset_of(std::array<idx, 3> const &point) const {
        return (point[order[2]] - low[order[2]]) * osize(0) * osize(1) + (point[order[1]] - low[order[1]]) * osize(0) + (point[order[0]] - low[order[0]]);
    }
};

using shape = std::vector<box3>;   // one box per rank

// which ranks take part in the intermediate stages (sub-communicator option)
struct rank_subset {
    int my_rank = -1;
    int active = 0;                 // number of working ranks (0: everybody)
    std::vector<int> slot;          // slot[r] = position of rank r among the working ranks, -1 if idle
    bool everybody() const { return slot.empty(); }
    void use_first(size_t all_ranks, int working){
        active = working;
        slot.assign(all_ranks, -1);
        for(int i=0; i<working; i++) slot[i] = i;
    }
};

inline box3 bounding_box(shape const &boxes){
    box3 w(boxes[0].low, boxes[0].high);
    for(auto const &b : boxes)
        for(int d=0; d<3; d++){ w.low[d] = std::min(w.low[d], b.low[d]); w.high[d] = std::max(w.high[d], b.high[d]); }
    return w;
}

inline void check_world(shape const &boxes, box3 const &world){
    idx total = 0;
    for(auto const &b : boxes) total += b.count();
    if (total < world.count()) throw std::invalid_argument("The provided input boxes do not fill the world box!");
    for(int d=0; d<3; d++) if (world.low[d] != 0) throw std::invalid_argument("Global box indexing must start from 0!");
    for(size_t i=0; i<boxes.size(); i++)
        for(size_t j=0; j<boxes.size(); j++)
            if (i != j and not boxes[i].overlap(boxes[j]).empty()) throw std::invalid_argument("Input boxes cannot overlap!");
}

inline bool spans(box3 const &world, shape const &boxes, int dim){
    for(auto const &b : boxes) if (b.size(dim) != world.size(dim)) return false;
    return true;
}
inline bool spans(box3 const &world, shape const &boxes, std::vector<int> const &dims){
    for(int d : dims) if (not spans(world, boxes, d)) return false;
    return true;
}
inline bool spans2(box3 const &world, shape const &boxes, int d1, int d2){
    for(auto const &b : boxes) if (b.size(d1) != world.size(d1) or b.size(d2) != world.size(d2)) return false;
    return true;
}
inline shape with_order(shape const &boxes, std::array<int, 3> o){
    shape r; r.reserve(boxes.size());
    for(auto const &b : boxes) r.push_back(b.reordered(o));
    return r;
}
inline bool extents_match(shape const &a, shape const &b){
    if (a.size() != b.size()) return false;
    for(size_t i=0; i<a.size(); i++) if (not a[i].same_extent(b[i])) return false;
    return true;
}

// ---- processor grids -------------------------------------------------------------------------------------
inline std::vector<std::array<int, 2>> factor_pairs(int n){
    std::vector<std::array<int, 2>> r;
    for(int i=1; i<=n; i++) if (n % i == 0) r.push_back({{i, n / i}});
    if (n == 1) r.push_back({{1, 1}});
    return r;
}
inline int grid_cost(std::array<int, 2> const &g){ return g[0] * g[1] + g[0] + g[1]; }

// pair of factors with the smallest g0*g1 + g0 + g1, first one wins ties
inline std::array<int, 2> grid2d(int nprocs){
    auto pairs = factor_pairs(nprocs);
    std::array<int, 2> best = pairs.front();
    for(auto const &p : pairs) if (grid_cost(p) < grid_cost(best)) best = p;
    return best;
}

// 3-D grid that is 1 along `flat_dim`, prefers `wanted`, falls back to other factor pairs when a dimension has
// fewer indexes than processors
inline std::array<int, 3> grid_flat(box3 const &world, int flat_dim, std::array<int, 2> wanted){
    auto lift = [&](std::array<int, 2> const &g){
        return (flat_dim == 0) ? std::array<int, 3>{{1, g[0], g[1]}} : ((flat_dim == 1) ? std::array<int, 3>{{g[0], 1, g[1]}} : std::array<int, 3>{{g[0], g[1], 1}});
    };
    auto fits = [&](std::array<int, 3> const &g){ for(int d=0; d<3; d++) if (g[d] > world.size(d)) return false; return true; };
    std::array<int, 3> result = lift(wanted);
    if (fits(result)) return result;
    auto pairs = factor_pairs(wanted[0] * wanted[1]);
    int bound = grid_cost(pairs.front());
    for(auto const &p : pairs) bound = std::max(bound, grid_cost(p));
    for(auto const &p : pairs){
        if (fits(lift(p)) and grid_cost(p) <= bound){ result = lift(p); bound = grid_cost(p); }
    }
    if (not fits(result)) throw std::runtime_error("Cannot split the given number of indexes into the given set of ranks: too few indexes.");
    return result;
}

// near-equal split of the world over a grid; ranks are numbered with grid dimension 0 fastest
inline shape split(box3 const &world, std::array<int, 3> const &grid, rank_subset const &subset = rank_subset()){
    auto cut = [&](int d, idx i){ return world.low[d] + i * (world.size(d) / grid[d]) + std::min<idx>(i, world.size(d) % grid[d]); };
    shape cells; cells.reserve((size_t)grid[0] * grid[1] * grid[2]);
    for(idx k=0; k<grid[2]; k++)
        for(idx j=0; j<grid[1]; j++)
            for(idx i=0; i<grid[0]; i++)
                cells.push_back(box3({{cut(0, i), cut(1, j), cut(2, k)}}, {{cut(0, i+1) - 1, cut(1, j+1) - 1, cut(2, k+1) - 1}}, world.order));
    if (subset.everybody()) return cells;
    shape spread;
    for(size_t r=0; r<subset.slot.size(); r++)
        spread.push_back((subset.slot[r] == -1) ? box3::nothing(world.order) : cells[subset.slot[r]]);
    return spread;
}

// greedy assignment of new boxes to ranks so that each rank keeps as much of its old data as possible
inline shape keep_local(shape const &fresh, shape const &old, std::array<int, 3> order, rank_subset const &subset){
    shape result; result.reserve(fresh.size());
    std::vector<bool> used(fresh.size(), false);
    if (not subset.everybody())
        for(size_t i=0; i<subset.slot.size(); i++) if (subset.slot[i] == -1) used[i] = true;
    for(size_t i=0; i<fresh.size(); i++){
        if (not subset.everybody() and subset.slot[i] == -1){ result.push_back(fresh[i].reordered(order)); continue; }
        int best_overlap = -1;
        size_t pick = fresh.size();
        for(size_t j=0; j<fresh.size(); j++){
            int shared = static_cast<int>(old[i].overlap(fresh[j]).count());  // int on purpose: same arithmetic as the reference
            if (not used[j] and shared > best_overlap){ best_overlap = shared; pick = j; }
        }
        if (pick >= fresh.size()) throw std::runtime_error("internal error: no box left to assign");
        used[pick] = true;
        result.push_back(fresh[pick].reordered(order));
    }
    return result;
}

inline long long count_links(shape const &a, shape const &b){
    long long n = 0;
    for(auto const &x : a) for(auto const &y : b) if (not x.overlap(y).empty()) n++;
    return n;
}

inline shape pencils(box3 const &world, std::array<int, 2> grid, int dim, shape const &source, std::array<int, 3> order,
                     rank_subset const &subset = rank_subset()){
    if (spans(world, source, dim)) return with_order(source, order);
    shape a = keep_local(split(world, grid_flat(world, dim, grid), subset), source, order, subset);
    shape b = keep_local(split(world, grid_flat(world, dim, {{grid[1], grid[0]}}), subset), source, order, subset);
    return (count_links(b, source) < count_links(a, source)) ? b : a;
}

inline shape slabs(box3 const &world, int nslabs, int dim1, int dim2, shape const &source, std::array<int, 3> order,
                   rank_subset const &subset){
    int const cut_dim = 3 - dim1 - dim2;   // the dimension that is neither dim1 nor dim2
    std::array<int, 3> grid{{1, 1, 1}};
    grid[cut_dim] = nslabs;
    return keep_local(split(world, grid, subset), source, order, subset);
}

// 3-D grid minimising the surface of the bricks (benchmarks/speed3d.h uses it for the in/out boxes)
inline std::array<int, 3> grid_min_surface(box3 const &world, int nprocs){
    if (nprocs == 1) return {{1, 1, 1}};
    std::array<idx, 3> n{{world.size(0), world.size(1), world.size(2)}};
    std::array<idx, 3> best{{1, 1, 1}};
    idx best_surface = std::numeric_limits<idx>::max();
    int const imax = static_cast<int>(std::min<idx>(nprocs, n[0]));
    for(int i=1; i<=imax; i++){
        if (nprocs % i != 0) continue;
        int const jmax = static_cast<int>(std::min<idx>(nprocs / i, n[1]));
        for(int j=1; j<=jmax; j++){
            if (jmax % j != 0) continue;        // (sic) the reference tests divisibility of jmax, kept for identical grids
            int const k = nprocs / (i * j);
            if (k > n[2] or i * j * k != nprocs) continue;
            std::array<idx, 3> cell{{n[0] / i, n[1] / j, n[2] / k}};
            idx surface = cell[0] * cell[1] + cell[1] * cell[2] + cell[2] * cell[0];
            if (surface < best_surface){ best_surface = surface; best = {{i, j, k}}; }
        }
    }
    return {{static_cast<int>(best[0]), static_cast<int>(best[1]), static_cast<int>(best[2])}};
}

} // namespace b200
