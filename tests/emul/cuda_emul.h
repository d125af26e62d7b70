// TEST INFRASTRUCTURE ONLY -- a tiny CPU emulation of the CUDA execution model (one std::thread per CUDA
// thread of a block, std::barrier for __syncthreads) so that tests/ can run the product's kernel SOURCE on the
// CPU-only development container.  Never compiled into the product library.
#pragma once
#include <barrier>
#include <condition_variable>
#include <mutex>
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

// A pool of host threads per calling thread (one worker per CUDA thread of a block), kept alive between launches: creating
// the threads again for every emulated launch used to dominate the run time of the CPU test-suite.
class worker_pool {
public:
    ~worker_pool(){
        { std::lock_guard<std::mutex> lock(guard); stopping = true; generation++; }
        wake.notify_all();
        for(auto &t : workers) t.join();
    }
    // runs job(0) ... job(n-1) concurrently, returns when all of them have finished
    void run(unsigned n, std::function<void(unsigned)> const &job){
        {
            std::lock_guard<std::mutex> lock(guard);
            while(workers.size() < n){
                unsigned const id = static_cast<unsigned>(workers.size());
                unsigned long long const seen = generation;
                workers.emplace_back([this, id, seen]{ loop(id, seen); });
            }
            current = &job; active = n; remaining = n; generation++;
        }
        wake.notify_all();
        std::unique_lock<std::mutex> lock(guard);
        done.wait(lock, [&]{ return remaining == 0; });
        current = nullptr;
    }
private:
    void loop(unsigned id, unsigned long long seen){
        for(;;){
            std::function<void(unsigned)> const *job = nullptr;
            {
                std::unique_lock<std::mutex> lock(guard);
                wake.wait(lock, [&]{ return generation != seen; });
                seen = generation;
                if (stopping) return;
                if (id < active) job = current;
            }
            if (job == nullptr) continue;
            (*job)(id);
            std::lock_guard<std::mutex> lock(guard);
            if (--remaining == 0) done.notify_all();
        }
    }
    std::vector<std::thread> workers;
    std::mutex guard;
    std::condition_variable wake, done;
    std::function<void(unsigned)> const *current = nullptr;
    unsigned active = 0, remaining = 0;
    unsigned long long generation = 0;
    bool stopping = false;
};
inline worker_pool& pool(){ static thread_local worker_pool p; return p; }

// The pool walks the blocks of the grid in order.  Every block gets its own __syncthreads barrier (threads that leave the
// kernel early drop out of it, like on the GPU) and the pool meets on a second barrier between blocks so that shared memory
// can be reused.
template<typename kernel_t, typename args_t>
void launch(kernel_t kernel, dim3 grid, dim3 block, size_t smem_bytes, args_t args){
    unsigned const nthreads = block.x * block.y;
    size_t const nblocks = static_cast<size_t>(grid.x) * grid.y * grid.z;
    if (nblocks == 0 or nthreads == 0) return;
    std::vector<block_state> states(nblocks);
    std::barrier<> between(nthreads);
    pool().run(nthreads, [&](unsigned id){
        unsigned const tx = id % block.x, ty = id / block.x;
        size_t index = 0;
        for(unsigned bz = 0; bz < grid.z; bz++)
        for(unsigned by = 0; by < grid.y; by++)
        for(unsigned bx = 0; bx < grid.x; bx++, index++){
            block_state &state = states[index];
            if (id == 0){
                state.smem.assign(smem_bytes + 64, 0);
                state.bar.reset(new std::barrier<>(nthreads));
            }
            between.arrive_and_wait();
            t_threadIdx = dim3(tx, ty); t_blockIdx = dim3(bx, by, bz); t_blockDim = block; t_gridDim = grid;
            t_block = &state;
            kernel(args);
            state.bar->arrive_and_drop();
            between.arrive_and_wait();
            if (id == 0){ state.smem.clear(); state.smem.shrink_to_fit(); }
        }
    });
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
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaMalloc(void **p, size_t bytes){ *p = std::malloc(bytes ? bytes : 1); return *p ? cudaSuccess : 2; }
template<typename T> inline cudaError_t cudaMalloc(T **p, size_t bytes){ return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t cudaFree(void *p){ std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, int){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, int, cudaStream_t){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemcpyPeerAsync(void *dst, int, const void *src, int, size_t bytes, cudaStream_t){ std::memmove(dst, src, bytes); return cudaSuccess; }
inline cudaError_t cudaMemset(void *dst, int value, size_t bytes){ std::memset(dst, value, bytes); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t){ std::memset(dst, value, bytes); return cudaSuccess; }
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
