#include "plan_logic.h"
#include "scatter_build.h"

#include <algorithm>
#include <cstdlib>

namespace b200 {

namespace {

// smallest dimension id not present in `taken`
int unused_dim(std::array<int, 3> const &taken){
    for(int d=0; d<3; d++) if (taken[0] != d and taken[1] != d and taken[2] != d) return d;
    return -1;
}

// bring `dim` to the front of an order by swapping it with whatever is there
std::array<int, 3> lead_with(std::array<int, 3> order, int dim){
    for(int i=1; i<3; i++) if (order[0] != dim and order[i] == dim) std::swap(order[0], order[i]);
    return order;
}

shape halve_all(shape const &boxes, int r2c_direction){
    if (r2c_direction == -1) return boxes;
    shape r; r.reserve(boxes.size());
    for(auto const &b : boxes) r.push_back(b.halved(r2c_direction));
    return r;
}

struct planner {
    box3 world_in, world_out;
    shape const &inboxes, &outboxes;
    int r2c;
    plan_options const &opt;
    rank_subset const &subset;
    std::array<int, 2> grid;
    int nworking;

    planner(shape const &in, shape const &out, int r2c_direction, plan_options const &o, rank_subset const &s)
        : world_in(bounding_box(in)), world_out(bounding_box(out)), inboxes(in), outboxes(out), r2c(r2c_direction), opt(o), subset(s){
        nworking = subset.everybody() ? static_cast<int>(in.size()) : subset.active;
        grid = grid2d(nworking);
    }

    // geometry for an FFT along `dim`: jump straight to `target` when it already has full lines in all of `need`
    shape stage_for(box3 const &world, int dim, shape const &current, box3 const &target_world, std::vector<int> const &need, shape const &target) const {
        bool const target_ok = spans(target_world, target, need);
        if (opt.use_reorder){
            if (target_ok) return (target[0].order[0] == dim) ? target : with_order(target, lead_with(target.front().order, dim));
            return pencils(world, grid, dim, current, lead_with(current.front().order, dim), subset);
        }
        return target_ok ? target : pencils(world, grid, dim, current, world.order, subset);
    }

    // first stage of an r2c transform: the output boxes are already shortened, so stretch them back before comparing
    shape first_stage(int dim) const {
        if (r2c != -1 and spans(world_out, outboxes, std::vector<int>{0, 1, 2})){
            shape stretched;
            for(auto const &b : outboxes){
                box3 s = b;
                s.low[r2c] = world_in.low[r2c];
                s.high[r2c] = world_in.high[r2c];
                stretched.push_back(s);
            }
            return stage_for(world_in, dim, inboxes, world_in, {0, 1, 2}, stretched);
        }
        return stage_for(world_in, dim, inboxes, world_out, {0, 1, 2}, outboxes);
    }

    shape slab_order(shape const &s, int dim) const {
        if (not opt.use_reorder or s.front().order[0] == dim) return s;
        return with_order(s, lead_with(s.front().order, dim));
    }

    logic_plan finish(shape const &s0, shape const &after0, shape const &s1, shape const &s2, std::array<int, 3> dirs) const {
        logic_plan p;
        p.in_shape[0] = inboxes; p.in_shape[1] = after0; p.in_shape[2] = s1; p.in_shape[3] = s2;
        p.out_shape[0] = s0;     p.out_shape[1] = s1;    p.out_shape[2] = s2; p.out_shape[3] = outboxes;
        p.fft_sizes = {{world_in.size(0), world_in.size(1), world_in.size(2)}};
        p.fft_direction = dirs;
        p.index_count = world_in.count();
        p.options = opt;
        p.rank = subset.my_rank;
        return p;
    }

    logic_plan by_pencils() const {
        std::array<int, 3> dir{{-1, -1, -1}};
        // first direction: forced by r2c, else one in which the input already holds full lines
        if (r2c != -1) dir[0] = r2c;
        else for(int d=0; d<3 and dir[0] == -1; d++) if (spans(world_in, inboxes, d)) dir[0] = d;
        // input that is a slab: its second full direction becomes the second transform
        if (dir[0] != -1 and spans(world_in, inboxes, dir[0]))
            for(int d=0; d<3 and dir[1] == -1; d++) if (d != dir[0] and spans(world_out, inboxes, d)) dir[1] = d;
        // last direction: one in which the output holds full lines
        for(int d=0; d<3 and dir[2] == -1; d++) if (d != dir[0] and d != dir[1] and spans(world_out, outboxes, d)) dir[2] = d;
        if (dir[0] == -1) dir[0] = unused_dim(dir);

        shape s0 = first_stage(dir[0]);
        shape after0 = halve_all(s0, r2c);

        if (dir[1] == -1){
            // (the reference walks d = 0,1,2: a direction with full lines wins, the fallback is taken after d = 0)
            for(int d=0; d<3; d++){
                if (d != dir[0] and d != dir[2] and spans(world_out, after0, d)){ dir[1] = d; break; }
                if (dir[1] == -1) dir[1] = unused_dim(dir);
            }
        }
        if (dir[2] == -1) dir[2] = unused_dim(dir);

        shape s1 = stage_for(world_out, dir[1], after0, world_out, {dir[1], dir[2]}, outboxes);
        shape s2 = stage_for(world_out, dir[2], s1, world_out, {dir[2]}, outboxes);
        return finish(s0, after0, s1, s2, dir);
    }

    logic_plan by_slabs() const {
        if (world_in.is2d()) return by_pencils();
        std::array<int, 3> dir{{-1, -1, -1}};

        // (1) the input is already a slab
        if (r2c == -1){
            for(int a=0; a<3 and dir[0] == -1; a++)
                for(int b=0; b<3; b++)
                    if (a != b and spans2(world_in, inboxes, a, b)){ dir[0] = a; dir[1] = b; break; }
        }else{
            for(int b=0; b<3; b++)
                if (b != r2c and spans2(world_in, inboxes, r2c, b)){ dir[0] = r2c; dir[1] = b; break; }
        }
        if (dir[0] != -1 and dir[1] != -1){
            dir[2] = unused_dim(dir);
            shape s0 = slab_order(inboxes, dir[0]);
            shape after0 = halve_all(s0, r2c);
            shape s1 = slab_order(after0, dir[1]);
            shape s2 = stage_for(world_out, dir[2], s1, world_out, {dir[2]}, outboxes);
            return finish(s0, after0, s1, s2, dir);
        }

        // (2) the output is a slab
        for(int a=0; a<3 and dir[0] == -1; a++)
            for(int b=0; b<3; b++)
                if (a != b and a != r2c and b != r2c and spans2(world_out, outboxes, a, b)){ dir[1] = a; dir[2] = b; break; }
        if (dir[1] != -1 and dir[2] != -1){
            dir[0] = unused_dim(dir);
            shape s0 = stage_for(world_in, dir[0], inboxes, world_out, {0, 1, 2}, outboxes);
            shape after0 = halve_all(s0, r2c);
            shape s1 = slab_order(outboxes, dir[1]);
            shape s2 = slab_order(s1, dir[2]);
            return finish(s0, after0, s1, s2, dir);
        }

        // (3) the input holds full lines in some direction
        if (r2c == -1){
            for(int d=0; d<3 and dir[0] == -1; d++) if (spans(world_in, inboxes, d)) dir[0] = d;
        }else if (spans(world_in, inboxes, r2c)){
            dir[0] = r2c;
        }
        if (dir[0] != -1){
            dir[1] = unused_dim(dir);
            dir[2] = unused_dim(dir);
            shape s0 = stage_for(world_in, dir[0], inboxes, world_out, {0, 1, 2}, outboxes);
            shape after0 = halve_all(s0, r2c);
            shape sl = slabs(world_out, nworking, dir[1], dir[2], after0, world_out.order, subset);
            shape s1 = slab_order(sl, dir[1]);
            shape s2 = slab_order(sl, dir[2]);
            return finish(s0, after0, s1, s2, dir);
        }

        // (4) bricks in, maybe full lines out: brick -> slab -> pencil (-> brick)
        for(int d=0; d<3 and dir[2] == -1; d++) if (d != r2c and spans(world_out, outboxes, d)) dir[2] = d;
        if (dir[2] != -1){
            dir[0] = (r2c == -1) ? unused_dim(dir) : r2c;
            dir[1] = unused_dim(dir);
        }else{
            dir[0] = (r2c != -1) ? r2c : 0;
            dir[1] = unused_dim(dir);
            dir[2] = unused_dim(dir);
        }
        shape sl = slabs(world_in, nworking, dir[0], dir[1], inboxes, world_in.order, subset);
        shape s0 = slab_order(sl, dir[0]);
        shape after0 = halve_all(s0, r2c);
        shape s1 = slab_order(after0, dir[1]);
        shape s2 = stage_for(world_out, dir[2], s1, world_out, {dir[2]}, outboxes);
        return finish(s0, after0, s1, s2, dir);
    }
};

bool one_order(shape const &s){
    for(auto const &b : s) if (not b.same_order(s.front())) return false;
    return true;
}

} // namespace

logic_plan make_logic_plan(shape const &inboxes, shape const &outboxes, int r2c_direction, plan_options const &options, int rank){
    if (inboxes.empty() or inboxes.size() != outboxes.size()) throw std::invalid_argument("need one input and one output box per rank");
    rank_subset subset;
    subset.my_rank = rank;
    if (options.subranks > 0 and static_cast<size_t>(options.subranks) < inboxes.size())
        subset.use_first(inboxes.size(), options.subranks);

    box3 const world_in = bounding_box(inboxes), world_out = bounding_box(outboxes);
    check_world(inboxes, world_in);
    check_world(outboxes, world_out);
    if (r2c_direction == -1){
        if (not world_in.same_extent(world_out)) throw std::invalid_argument("input and output boxes span different index sets");
    }else{
        if (not world_in.halved(r2c_direction).same_extent(world_out)) throw std::invalid_argument("output boxes do not span the half-complex index set");
    }
    if (not one_order(inboxes) or not one_order(outboxes)) throw std::invalid_argument("all boxes of one side must share the same order");

    planner p(inboxes, outboxes, r2c_direction, options, subset);
    return options.use_pencils ? p.by_pencils() : p.by_slabs();
}

namespace {
// maximum-weight perfect matching of ranks (rows) to boxes (columns): Hungarian algorithm with potentials, O(n^3)
std::vector<int> best_assignment(std::vector<double> const &weight, int n){
    double top = 0;
    for(double w : weight) top = std::max(top, w);
    auto cost = [&](int r, int c){ return top - weight[static_cast<size_t>(r) * n + c]; };
    double const inf = 1e300;
    std::vector<double> u(n + 1, 0.0), v(n + 1, 0.0);
    std::vector<int> match(n + 1, 0), way(n + 1, 0);      // match[c] = row assigned to column c (1-based)
    for(int i=1; i<=n; i++){
        match[0] = i;
        int j0 = 0;
        std::vector<double> minv(n + 1, inf);
        std::vector<char> used(n + 1, 0);
        do{
            used[j0] = 1;
            int const i0 = match[j0];
            int j1 = 0;
            double delta = inf;
            for(int j=1; j<=n; j++) if (not used[j]){
                double const cur = cost(i0 - 1, j - 1) - u[i0] - v[j];
                if (cur < minv[j]){ minv[j] = cur; way[j] = j0; }
                if (minv[j] < delta){ delta = minv[j]; j1 = j; }
            }
            for(int j=0; j<=n; j++){
                if (used[j]){ u[match[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            j0 = j1;
        }while(match[j0] != 0);
        do{ int const j1 = way[j0]; match[j0] = match[j1]; j0 = j1; }while(j0);
    }
    std::vector<int> box_of_rank(n, 0);
    for(int j=1; j<=n; j++) box_of_rank[match[j] - 1] = j - 1;
    return box_of_rank;
}
}

int balance_traffic(logic_plan &plan, int r2c_direction){
    int const n = static_cast<int>(plan.in_shape[0].size());
    if (n < 2 or n > 64) return 0;
    if (plan.options.subranks > 0 and plan.options.subranks < n) return 0;     // idle ranks must stay idle
    auto weight_of = [&](int i){ return (i == 0 and r2c_direction != -1) ? 1.0 : 2.0; };   // reals of the working precision per element
    // cost of a plan: sum over the reshapes of the busiest link end (out or in), then the total volume
    auto cost = [&](logic_plan const &p, double &busiest, double &volume){
        busiest = 0; volume = 0;
        for(int i=0; i<4; i++){
            std::vector<double> out(n, 0.0), in(n, 0.0);
            for(int r=0; r<n; r++) for(int q=0; q<n; q++) if (q != r){
                double const w = weight_of(i) * static_cast<double>(p.in_shape[i][r].overlap(p.out_shape[i][q]).count());
                out[r] += w; in[q] += w;
            }
            double worst = 0;
            for(int r=0; r<n; r++){ worst = std::max(worst, std::max(out[r], in[r])); volume += out[r]; }
            busiest += worst;
        }
    };
    auto better = [](double busy, double volume, double best_busy, double best_volume){
        double const eps = 1e-9 * (best_busy + best_volume + 1.0);
        return busy < best_busy - eps or (busy <= best_busy + eps and volume < best_volume - eps);
    };
    // the boxes of stage t are the outputs of reshape t and the inputs of reshape t+1: give every rank the box that shares the
    // most with what the rank holds before (reshape t) and / or after (reshape t+1), the other stages fixed
    auto assign_stage = [&](logic_plan &p, int t, bool look_back, bool look_ahead){
        std::vector<double> w(static_cast<size_t>(n) * n, 0.0);
        for(int r=0; r<n; r++) for(int c=0; c<n; c++){
            double share = 0;
            if (look_back)  share += weight_of(t) * static_cast<double>(p.in_shape[t][r].overlap(p.out_shape[t][c]).count());
            if (look_ahead) share += weight_of(t+1) * static_cast<double>(p.in_shape[t+1][c].overlap(p.out_shape[t+1][r]).count());
            w[static_cast<size_t>(r) * n + c] = share + ((r == c) ? 1e-3 : 0.0);     // ties keep the reference's choice
        }
        std::vector<int> const pick = best_assignment(w, n);
        shape const outs = p.out_shape[t], ins = p.in_shape[t+1];
        for(int r=0; r<n; r++){ p.out_shape[t][r] = outs[pick[r]]; p.in_shape[t+1][r] = ins[pick[r]]; }
    };
    // coordinate descent from a starting plan: one stage at a time against both neighbours, strict improvements only
    auto descend = [&](logic_plan &p){
        double busy, volume;
        cost(p, busy, volume);
        for(int sweep=0; sweep<4; sweep++){
            bool improved = false;
            for(int t=2; t>=0; t--){
                logic_plan trial = p;
                assign_stage(trial, t, true, true);
                double b, v;
                cost(trial, b, v);
                if (better(b, v, busy, volume)){ p = trial; busy = b; volume = v; improved = true; }
            }
            if (not improved) break;
        }
    };
    double best_busy, best_volume;
    cost(plan, best_busy, best_volume);
    int changes = 0;
    // candidates: descent from the reference's plan; a backward chain (last stage placed next to the output boxes, then each
    // earlier stage next to both neighbours); a forward chain (first stage next to the input boxes, ...)
    logic_plan candidate[3] = {plan, plan, plan};
    assign_stage(candidate[1], 2, false, true); assign_stage(candidate[1], 1, true, true); assign_stage(candidate[1], 0, true, true);
    assign_stage(candidate[2], 0, true, false); assign_stage(candidate[2], 1, true, true); assign_stage(candidate[2], 2, true, true);
    for(auto &c : candidate){
        descend(c);
        double busy, volume;
        cost(c, busy, volume);
        if (better(busy, volume, best_busy, best_volume)){ plan = c; best_busy = busy; best_volume = volume; changes++; }
    }
    // polish: pairwise swaps that lower the busiest end (small groups only: O(n^4) box overlaps per sweep)
    if (n <= 16){
        for(int sweep=0; sweep<4; sweep++){
            bool improved = false;
            for(int t=2; t>=0; t--) for(int p=0; p<n; p++) for(int q=p+1; q<n; q++){
                std::swap(plan.out_shape[t][p], plan.out_shape[t][q]);
                std::swap(plan.in_shape[t+1][p], plan.in_shape[t+1][q]);
                double busy, volume;
                cost(plan, busy, volume);
                if (better(busy, volume, best_busy, best_volume)){ best_busy = busy; best_volume = volume; improved = true; changes++; }
                else{
                    std::swap(plan.out_shape[t][p], plan.out_shape[t][q]);
                    std::swap(plan.in_shape[t+1][p], plan.in_shape[t+1][q]);
                }
            }
            if (not improved) break;
        }
    }
    return changes;
}

// Estimated cost of executing a plan with the fused reshapes, in bytes that cross NVLink at the busiest GPU (reals of the
// working precision count 1): every reshape that moves data costs what its busiest rank sends or receives, but never less
// than the HBM traffic of the transform fused in front of it; a transform with nothing to fuse into, and the final copy into
// the caller's array, cost their HBM traffic.  HBM bytes are converted with the ratio of the two rates (~770 GB/s : ~6.5 TB/s).
double execution_cost(logic_plan const &p, int r2c_direction){
    int const n = static_cast<int>(p.in_shape[0].size());
    double const hbm_to_nvlink = 0.12;
    auto weight_of = [&](int i){ return (i == 0 and r2c_direction != -1) ? 1.0 : 2.0; };
    double total = 0;
    for(int i=0; i<4; i++){
        bool const moves = not (extents_match(p.in_shape[i], p.out_shape[i]) and p.in_shape[i][0].same_order(p.out_shape[i][0]));
        double busiest = 0, largest = 0;
        std::vector<double> out(n, 0.0), in(n, 0.0);
        for(int r=0; r<n; r++){
            largest = std::max(largest, weight_of(i) * static_cast<double>(p.in_shape[i][r].count()));
            for(int q=0; q<n; q++) if (q != r){
                double const w = weight_of(i) * static_cast<double>(p.in_shape[i][r].overlap(p.out_shape[i][q]).count());
                out[r] += w; in[q] += w;
            }
        }
        for(int r=0; r<n; r++) busiest = std::max(busiest, std::max(out[r], in[r]));
        double const pass = hbm_to_nvlink * 2.0 * largest;            // read + write of the box in local memory
        if (moves) total += std::max(busiest, pass);
        else if (i > 0) total += pass;                                // transform i-1 runs alone
        if (i == 3){
            if (moves) total += pass;                                 // copy from the arena into the caller's array
            // the third transform runs before reshape 3 whether or not it moves: counted above in both branches
        }
    }
    return total;
}

// true when every reshape of the plan, in both directions and on every rank, can be expressed as a scatter map (at most
// scatter_max_cuts cells per axis): otherwise the plan falls back, collectively, to pack / exchange / unpack
bool fits_scatter_maps(logic_plan const &p){
    int const n = static_cast<int>(p.in_shape[0].size());
    std::vector<void*> bases(n, nullptr);
    scatter_map map;
    std::string why;
    for(int i=0; i<4; i++){
        if (extents_match(p.in_shape[i], p.out_shape[i]) and p.in_shape[i][0].same_order(p.out_shape[i][0])) continue;
        for(int r=0; r<n; r++){
            if (not build_scatter_map(p.in_shape[i][r], 0, p.out_shape[i], bases, 1, map, why)) return false;
            if (not build_scatter_map(p.out_shape[i][r], 0, p.in_shape[i], bases, 1, map, why)) return false;
        }
    }
    return true;
}

logic_plan make_execution_plan(shape const &inboxes, shape const &outboxes, int r2c_direction, plan_options const &options, int rank, int *swaps){
    if (swaps) *swaps = 0;
    const char *keep = std::getenv("HEFFTE_B200_REFERENCE_PLAN");
    if (keep != nullptr and keep[0] != '0') return make_logic_plan(inboxes, outboxes, r2c_direction, options, rank);
    // both decompositions are planned (pencils: three exchanges of part of the data; slabs: two exchanges of more of it) and
    // the cheaper one by execution_cost() runs; the caller's use_pencils decides ties.  HEFFTE_B200_DECOMPOSITION=pencils|slabs
    // forces one of them.
    const char *forced = std::getenv("HEFFTE_B200_DECOMPOSITION");
    logic_plan best;
    double best_cost = -1;
    int best_swaps = 0;
    for(int attempt=0; attempt<2; attempt++){
        plan_options exec_options = options;
        exec_options.use_reorder = false;
        if (attempt == 1) exec_options.use_pencils = not options.use_pencils;
        if (forced != nullptr and (forced[0] == 'p' or forced[0] == 's')){
            if (attempt == 1) break;
            exec_options.use_pencils = (forced[0] == 'p');
        }else if (options.explicit_decomposition and attempt == 1) break;      // the caller's choice runs
        if (attempt == 1 and options.subranks > 0) break;     // sub-communicator plans keep the caller's decomposition
        try{
            logic_plan plan = make_logic_plan(inboxes, outboxes, r2c_direction, exec_options, rank);
            int const n = balance_traffic(plan, r2c_direction);
            double c = execution_cost(plan, r2c_direction);
            if (not fits_scatter_maps(plan)) c *= 4.0;      // the exchange path: about four times slower (profiles/r01_multi_8gpu)
            if (best_cost < 0 or c < best_cost * (1.0 - 1e-9)){ best = plan; best_cost = c; best_swaps = n + ((attempt == 1) ? 1 : 0); }
        }catch(std::exception &){
            if (attempt == 0) throw;
        }
    }
    best.options.use_pencils = options.use_pencils;
    if (swaps) *swaps = best_swaps;
    return best;
}

std::vector<std::array<int, 3>> stage_grids(logic_plan const &plan){
    std::vector<std::array<int, 3>> grids;
    for(int stage=0; stage<5; stage++){
        shape const &s = (stage < 4) ? plan.in_shape[stage] : plan.out_shape[3];
        std::vector<std::array<idx, 2>> seen[3];
        for(auto const &b : s)
            for(int d=0; d<3; d++){
                std::array<idx, 2> range{{b.low[d], b.high[d]}};
                if (range == std::array<idx, 2>{{0, -1}}) continue;
                if (std::find(seen[d].begin(), seen[d].end(), range) == seen[d].end()) seen[d].push_back(range);
            }
        grids.push_back({{static_cast<int>(seen[0].size()), static_cast<int>(seen[1].size()), static_cast<int>(seen[2].size())}});
    }
    return grids;
}

} // namespace b200
