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
struct int4    { int x, y, z, w; };
struct dim3    { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emul {
struct block_state {
    std::vector<unsigned char> smem;
    std::unique_ptr<std::barrier<>> bar;
};
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local block_state *t_block = nullptr;

// One pool of host threads per launch (one per CUDA thread of a block); the pool walks the blocks of the grid in order.
// Every block gets its own __syncthreads barrier (threads that leave the kernel early drop out of it, like on the GPU) and
// the pool meets on a second barrier between blocks so shared memory can be reused.
template<typename kernel_t, typename args_t>
void launch(kernel_t kernel, dim3 grid, dim3 block, size_t smem_bytes, args_t args){
    unsigned const nthreads = block.x * block.y;
    size_t const nblocks = static_cast<size_t>(grid.x) * grid.y * grid.z;
    if (nblocks == 0 or nthreads == 0) return;
    std::vector<block_state> states(nblocks);
    std::barrier<> between(nthreads);
    std::vector<std::thread> pool;
    for(unsigned ty = 0; ty < block.y; ty++)
    for(unsigned tx = 0; tx < block.x; tx++){
        pool.emplace_back([&, tx, ty]{
            size_t index = 0;
            for(unsigned bz = 0; bz < grid.z; bz++)
            for(unsigned by = 0; by < grid.y; by++)
            for(unsigned bx = 0; bx < grid.x; bx++, index++){
                block_state &state = states[index];
                if (tx == 0 and ty == 0){
                    state.smem.assign(smem_bytes + 64, 0);
                    state.bar.reset(new std::barrier<>(nthreads));
                }
                between.arrive_and_wait();
                t_threadIdx = dim3(tx, ty); t_blockIdx = dim3(bx, by, bz); t_blockDim = block; t_gridDim = grid;
                t_block = &state;
                kernel(args);
                state.bar->arrive_and_drop();
                between.arrive_and_wait();
                if (tx == 0 and ty == 0){ state.smem.clear(); state.smem.shrink_to_fit(); }
            }
        });
    }
    for(auto &t : pool) t.join();
}
} // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx  (emul::t_blockIdx)
#define blockDim  (emul::t_blockDim)
#define gridDim   (emul::t_gridDim)
#define B200_DYN_SMEM(name) unsigned char *name = emul::t_block->smem.data()
inline void __syncthreads(){ emul::t_block->bar->arrive_and_wait(); }
template<typename T> inline T __ldg(const T *p){ return *p; }

// ---- a synchronous stand-in for the few CUDA runtime calls the host side of the product makes -------------------------------
// "Device memory" is host memory, streams execute immediately in the calling thread, events are no-ops.  This lets the
// WHOLE library (planner, reshape schedules, peer-memory mode with ranks as threads) run in the CPU-only test-suite.
#include <cstdlib>
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaSuccess = 0, cudaErrorEmulation = 999, cudaErrorPeerAccessAlreadyEnabled = 704 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaMalloc(void **p, size_t bytes){ *p = std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : 2; }
template<typename T> inline cudaError_t cudaMalloc(T **p, size_t bytes){ return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t cudaFree(void *p){ std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, int){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, int, cudaStream_t){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpyPeerAsync(void *dst, int, const void *src, int, size_t bytes, cudaStream_t){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemset(void *dst, int value, size_t bytes){ std::memset(dst, value, bytes); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t){ return cudaSuccess; }
enum { cudaStreamNonBlocking = 1 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned){ static char token[64]; static int next = 0; *s = token + (next++ % 64); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t){ return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize(){ return cudaSuccess; }
inline cudaError_t cudaGetLastError(){ return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError(){ return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t){ return "emulated CUDA runtime"; }
inline cudaError_t cudaGetDeviceCount(int *n){ *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d){ *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int){ return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int){ *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned){ return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned){ *e = reinterpret_cast<void*>(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t){ return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t){ return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned){ return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*){ return cudaErrorEmulation; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned){ return cudaErrorEmulation; }
inline cudaError_t cudaIpcCloseMemHandle(void*){ return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(const void*, int, int){ return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e){ *e = reinterpret_cast<void*>(1); return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t){ *ms = 0; return cudaSuccess; }
