// Logic plan of the distributed transform: which boxes every rank holds before/after each of the four reshapes and
// along which dimension each 1-D FFT stage runs.  The decisions reproduce the reference planner
// (src/heffte_plan_logic.cpp:163-254 pencils, :273-422 slabs, :424-453 entry point; options in
// include/heffte_plan_logic.h:48-57, 131-176) so that a b200 plan moves exactly the same sub-boxes between the same
// ranks; tests/test_plan_logic.py compares every box with the reference's plan_operations().
#pragma once

#include "geometry.h"

namespace b200 {

// values match heffte::reshape_algorithm (include/heffte_plan_logic.h:48-57)
enum reshape_algorithm : int { alg_alltoallv = 0, alg_p2p_plined = 1, alg_p2p = 2, alg_alltoall = 3 };

struct plan_options {
    bool use_reorder = false;       // default of the GPU backends (reference include/heffte_backend_cuda.h:854-857)
    int algorithm = alg_alltoallv;
    bool use_pencils = true;
    bool explicit_decomposition = false;   // the caller chose pencils / slabs: the executed plan keeps it (else the cheaper one runs)
    bool use_gpu_aware = true;
    int subranks = -1;
};

struct logic_plan {
    shape in_shape[4], out_shape[4];
    std::array<idx, 3> fft_sizes{{0, 0, 0}};
    std::array<int, 3> fft_direction{{-1, -1, -1}};
    idx index_count = 0;
    plan_options options;
    int rank = 0;
};

// r2c_direction = -1 for complex-to-complex and real-to-real transforms
logic_plan make_logic_plan(shape const &inboxes, shape const &outboxes, int r2c_direction, plan_options const &options, int rank);

// Execution-plan refinement (not in the reference): the boxes of the three intermediate stages are permuted among the ranks
// so that the busiest GPU of every reshape sends and receives as little as possible.  The reference assigns pencil j of
// the last stage to rank j whatever the output bricks are (src/heffte_plan_logic.cpp:163-254); at 512^3 on 8 ranks that
// makes 6 of the 8 ranks ship their WHOLE pencil in the last reshape (268 MB instead of 134 MB).  Results are unchanged:
// only which rank works on which pencil.  A plan that is already balanced is returned untouched (strict improvements only).
// Returns the number of box swaps applied.
int balance_traffic(logic_plan &plan, int r2c_direction);

// The plan a b200 transform EXECUTES: the reference's plan for the same options with two refinements that do not change
// any result -- (1) no reorder of the intermediate boxes: the strided kernels run at the same HBM rate as the contiguous
// ones, and without a transposition every store of a fused reshape stays a full 128-byte row on the far side of NVLink;
// (2) balance_traffic(); (3) the decomposition -- pencils or slabs -- that moves the least over NVLink at the busiest GPU
// (execution_cost(); a caller that sets use_pencils explicitly -- plan_options::explicit_decomposition -- gets exactly that
// decomposition; HEFFTE_B200_DECOMPOSITION=pencils|slabs forces one for experiments).
// HEFFTE_B200_REFERENCE_PLAN=1 in the environment returns the reference's plan unchanged.
// The sizes a plan REPORTS (size_workspace) always come from the reference's plan.
double execution_cost(logic_plan const &plan, int r2c_direction);
bool fits_scatter_maps(logic_plan const &plan);
logic_plan make_execution_plan(shape const &inboxes, shape const &outboxes, int r2c_direction, plan_options const &options, int rank,
                               int *swaps = nullptr);

// process-grid extents of the five stages (benchmark printout, reference src/heffte_plan_logic.cpp:460-487)
std::vector<std::array<int, 3>> stage_grids(logic_plan const &plan);

} // namespace b200
