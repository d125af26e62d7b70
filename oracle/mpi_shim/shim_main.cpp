/*
 * TEST INFRASTRUCTURE ONLY.  Entry point for reference programs (benchmarks/speed3d_*.cpp, test/*.cpp)
 * compiled with -Dmain=shim_rank_main: runs the program's own main() on SHIM_NP thread-ranks.
 */
#include "mpi.h"
int shim_rank_main(int argc, char **argv);
#undef main
int main(int argc, char **argv){
    return shim_run(shim_world_size_from_env(), shim_rank_main, argc, argv);
}
