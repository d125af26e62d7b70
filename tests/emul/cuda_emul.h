// TEST INFRASTRUCTURE ONLY -- a tiny CPU emulation of the CUDA execution model (one std::thread per CUDA
// thread of a block, std::barrier for __syncthreads) so that tests/ can run the product's kernel SOURCE on the
// CPU-only development container.  Never compiled into the product library.
#pragma once
#include <barrier>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <cstring>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)
#define __restrict__

struct float2  { float x, y; };
struct double2 { double x, y; };
struct float4  { float x, y, z, w; };
struct uint4   { unsigned x, y, z, w; };
struct dim3    { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emul {
struct block_state {
    std::vector<unsigned char> smem;
    std::unique_ptr<std::barrier<>> bar;
};
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local block_state *t_block = nullptr;

template<typename kernel_t, typename args_t>
void launch(kernel_t kernel, dim3 grid, dim3 block, size_t smem_bytes, args_t args){
    for(unsigned by = 0; by < grid.y; by++)
    for(unsigned bx = 0; bx < grid.x; bx++){
        block_state state;
        state.smem.assign(smem_bytes + 64, 0);
        unsigned const nthreads = block.x * block.y;
        state.bar.reset(new std::barrier<>(nthreads));
        std::vector<std::thread> pool;
        for(unsigned ty = 0; ty < block.y; ty++)
        for(unsigned tx = 0; tx < block.x; tx++){
            pool.emplace_back([&, tx, ty]{
                t_threadIdx = dim3(tx, ty); t_blockIdx = dim3(bx, by); t_blockDim = block; t_gridDim = grid;
                t_block = &state;
                kernel(args);
                state.bar->arrive_and_drop();
            });
        }
        for(auto &t : pool) t.join();
    }
}
}

#define threadIdx (emul::t_threadIdx)
#define blockIdx  (emul::t_blockIdx)
#define blockDim  (emul::t_blockDim)
#define gridDim   (emul::t_gridDim)
#define B200_DYN_SMEM(name) unsigned char *name = emul::t_block->smem.data()
inline void __syncthreads(){ emul::t_block->bar->arrive_and_wait(); }
template<typename T> inline T __ldg(const T *p){ return *p; }
