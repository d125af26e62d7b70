// Developer micro-benchmark (not part of the product), 2 GPUs of one box in one process: how fast do remote stores into a peer
// GPU go (per-thread 16-byte stores against TMA bulk stores from shared memory), how many CTAs does the link need, and does an
// HBM-bound local kernel on a second stream overlap with them.  Run under gpurun --gpus 2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{ cudaError_t e = (x); if (e != cudaSuccess){ printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int TILE = 64 * 1024;      // bytes per CTA iteration

// local -> shared (cp.async, 16 bytes per thread) -> peer with per-thread 16-byte stores
__global__ void __launch_bounds__(256) remote_copy_st(const int4 *src, int4 *dst, long long tiles){
    extern __shared__ __align__(128) unsigned char smem[];
    int4 *sm = reinterpret_cast<int4*>(smem);
    for(long long t = blockIdx.x; t < tiles; t += gridDim.x){
        const int4 *s = src + t * (TILE / 16);
        for(int i = threadIdx.x; i < TILE / 16; i += 256){
            unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(sm + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(a), "l"(s + i));
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        int4 *d = dst + t * (TILE / 16);
        for(int i = threadIdx.x; i < TILE / 16; i += 256) d[i] = sm[i];
        __syncthreads();
    }
}

// local -> shared by one bulk copy, shared -> peer by bulk stores of `row` bytes each (row = 128: the rows of an FFT tile)
__global__ void __launch_bounds__(256) remote_copy_bulk(const char *src, char *dst, long long tiles, int row){
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_addr = static_cast<unsigned>(__cvta_generic_to_shared(&bar));
    const unsigned sm_addr = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    if (threadIdx.x == 0){
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned phase = 0;
    for(long long t = blockIdx.x; t < tiles; t += gridDim.x){
        if (threadIdx.x == 0){
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar_addr), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         :: "r"(sm_addr), "l"(src + t * TILE), "r"(TILE), "r"(bar_addr) : "memory");
        }
        // everybody waits for the tile
        unsigned done = 0;
        while(!done){
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar_addr), "r"(phase) : "memory");
        }
        phase ^= 1;
        // (an FFT would work on the tile here: generic-proxy writes, then the proxy fence)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        const int rows = TILE / row;
        for(int r = threadIdx.x; r < rows; r += 256){
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                         :: "l"(dst + t * TILE + (long long) r * row), "r"(sm_addr + r * row), "r"(row) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");     // the tile may be overwritten
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

__global__ void __launch_bounds__(256) local_copy(const int4 *src, int4 *dst, long long n){
    const long long step = (long long) gridDim.x * blockDim.x;
    for(long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) dst[i] = src[i];
}

int main(){
    int ndev = 0; CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2){ printf("needs 2 GPUs\n"); return 0; }
    const long long remote_bytes = 512LL << 20, local_bytes = 1LL << 30;
    char *peer; CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, remote_bytes)); CK(cudaMemset(peer, 0, remote_bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    char *src, *la, *lb; CK(cudaMalloc(&src, remote_bytes)); CK(cudaMalloc(&la, local_bytes)); CK(cudaMalloc(&lb, local_bytes));
    CK(cudaMemset(src, 1, remote_bytes)); CK(cudaMemset(la, 2, local_bytes));
    CK(cudaFuncSetAttribute(remote_copy_st, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE));
    CK(cudaFuncSetAttribute(remote_copy_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE));
    int lo, hi; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s_remote, s_local; CK(cudaStreamCreateWithPriority(&s_remote, cudaStreamNonBlocking, hi)); CK(cudaStreamCreateWithPriority(&s_local, cudaStreamNonBlocking, lo));
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const long long tiles = remote_bytes / TILE;
    auto remote = [&](int mode, int grid, int row){
        if (mode == 0) remote_copy_st<<<grid, 256, TILE, s_remote>>>(reinterpret_cast<const int4*>(src), reinterpret_cast<int4*>(peer), tiles);
        else remote_copy_bulk<<<grid, 256, TILE, s_remote>>>(src, peer, tiles, row);
    };
    auto local = [&]{ local_copy<<<148 * 8, 256, 0, s_local>>>(reinterpret_cast<const int4*>(la), reinterpret_cast<int4*>(lb), local_bytes / 16); };
    auto time_one = [&](auto f, cudaStream_t s){
        for(int i=0; i<2; i++) f();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s)); for(int i=0; i<5; i++) f(); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError()); return ms / 5;
    };
    float t_local = time_one(local, s_local);
    printf("local copy alone (1 GiB read + 1 GiB write): %.3f ms, %.0f GB/s\n", t_local, 2.0 * local_bytes / t_local * 1e-6);
    // correctness of the bulk path
    CK(cudaMemset(src, 7, remote_bytes)); remote(1, 296, 128); CK(cudaDeviceSynchronize());
    { unsigned char probe[4]; CK(cudaMemcpy(probe, peer + remote_bytes - 4, 4, cudaMemcpyDefault)); printf("bulk store check: %d %d (expect 7 7)\n", probe[0], probe[3]); }
    for(int mode = 0; mode < 2; mode++){
        for(int row : {128, 1024, 65536}){
            if (mode == 0 and row != 128) continue;
            for(int grid : {74, 148, 296, 444}){
                float t = time_one([&]{ remote(mode, grid, row); }, s_remote);
                // concurrent: remote first (high priority), local on the other stream; wall time of both
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, s_remote));
                CK(cudaStreamWaitEvent(s_local, e0, 0));
                for(int i=0; i<3; i++){ remote(mode, grid, row); local(); }
                CK(cudaEventRecord(e1, s_remote)); CK(cudaEventRecord(e2, s_local));
                CK(cudaStreamWaitEvent(s_remote, e2, 0));
                CK(cudaEventRecord(e1, s_remote));
                CK(cudaEventSynchronize(e1));
                float both; CK(cudaEventElapsedTime(&both, e0, e1)); both /= 3;
                printf("%-28s grid %3d row %6d: alone %.3f ms %6.1f GB/s | with the local copy %.3f ms (sum %.3f, max %.3f)\n",
                       mode == 0 ? "remote: 16-byte stores" : "remote: TMA bulk stores", grid, row, t, remote_bytes / t * 1e-6, both, t + t_local, t > t_local ? t : t_local);
                fflush(stdout);
            }
        }
    }
    return 0;
}
