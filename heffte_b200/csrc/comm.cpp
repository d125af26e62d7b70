#include "comm.h"

#include <dlfcn.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <map>
#include <string>
#include <memory>
#include <mutex>

namespace b200 {

namespace {

// ---- the handful of NCCL entry points we need, resolved with dlsym ---------------------------------------
struct nccl_comm_opaque;
typedef nccl_comm_opaque* nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { nccl_success = 0 };
enum { nccl_int8 = 0 };   // ncclInt8 / ncclChar

struct nccl_api {
    void *handle = nullptr;
    int (*GetUniqueId)(nccl_uid*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string problem;
};

nccl_api& nccl(){
    static nccl_api api;
    static std::once_flag once;
    std::call_once(once, []{
        const char *env = std::getenv("HEFFTE_B200_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for(const char *name : names){
            if (name == nullptr) continue;
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (not api.handle){ api.problem = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
        auto get = [&](const char *symbol){ void *p = dlsym(api.handle, symbol); if (p == nullptr) api.problem = std::string("missing NCCL symbol ") + symbol; return p; };
        api.GetUniqueId  = reinterpret_cast<decltype(api.GetUniqueId)>(get("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(get("ncclCommInitRank"));
        api.CommDestroy  = reinterpret_cast<decltype(api.CommDestroy)>(get("ncclCommDestroy"));
        api.Send         = reinterpret_cast<decltype(api.Send)>(get("ncclSend"));
        api.Recv         = reinterpret_cast<decltype(api.Recv)>(get("ncclRecv"));
        api.GroupStart   = reinterpret_cast<decltype(api.GroupStart)>(get("ncclGroupStart"));
        api.GroupEnd     = reinterpret_cast<decltype(api.GroupEnd)>(get("ncclGroupEnd"));
        api.AllGather    = reinterpret_cast<decltype(api.AllGather)>(get("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(get("ncclGetErrorString"));
    });
    return api;
}

class self_communicator : public communicator {
public:
    int allgather(const void *mine, void *all, size_t bytes) override { std::memcpy(all, mine, bytes); return 0; }
    int exchange(std::vector<transfer> const &sends, std::vector<transfer> const &recvs, cudaStream_t stream) override {
        // only self-to-self transfers can exist
        for(size_t i=0; i<sends.size() and i<recvs.size(); i++)
            if (sends[i].data != recvs[i].data and sends[i].bytes > 0)
                if (cudaMemcpyAsync(recvs[i].data, sends[i].data, sends[i].bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) return 1;
        return 0;
    }
    int barrier(cudaStream_t) override { return 0; }
    const char* kind() const override { return "self"; }
};

class nccl_communicator : public communicator {
public:
    nccl_communicator(int rank, int size, nccl_comm_t c) : comm(c){ my_rank = rank; nranks = size; }
    ~nccl_communicator() override {
        if (side) cudaStreamDestroy(side);
        if (scratch) cudaFree(scratch);
        if (comm) nccl().CommDestroy(comm);
    }
    int allgather(const void *mine, void *all, size_t bytes) override {
        size_t const need = bytes * (nranks + 1);
        if (need > scratch_bytes){
            if (scratch) cudaFree(scratch);
            if (cudaMalloc(&scratch, need) != cudaSuccess){ scratch = nullptr; scratch_bytes = 0; return 1; }
            scratch_bytes = need;
        }
        // on a stream of its own: the plan-time collectives never touch the legacy default stream (which would serialise with
        // whatever the caller has in flight there)
        if (side == nullptr and cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess){ side = nullptr; return 1; }
        char *send = static_cast<char*>(scratch), *recv = send + bytes;
        if (cudaMemcpyAsync(send, mine, bytes, cudaMemcpyHostToDevice, side) != cudaSuccess) return 1;
        if (nccl().AllGather(send, recv, bytes, nccl_int8, comm, side) != nccl_success) return 2;
        if (cudaMemcpyAsync(all, recv, bytes * nranks, cudaMemcpyDeviceToHost, side) != cudaSuccess) return 1;
        if (cudaStreamSynchronize(side) != cudaSuccess) return 1;
        return 0;
    }
    int exchange(std::vector<transfer> const &sends, std::vector<transfer> const &recvs, cudaStream_t stream) override {
        nccl_api &api = nccl();
        if (api.GroupStart() != nccl_success) return 2;
        int status = 0;
        for(auto const &r : recvs)
            if (r.bytes > 0 and api.Recv(r.data, r.bytes, nccl_int8, r.peer, comm, stream) != nccl_success) status = 2;
        for(auto const &s : sends)
            if (s.bytes > 0 and api.Send(s.data, s.bytes, nccl_int8, s.peer, comm, stream) != nccl_success) status = 2;
        if (api.GroupEnd() != nccl_success) status = 2;
        return status;
    }
    int barrier(cudaStream_t stream) override {
        if (scratch_bytes < static_cast<size_t>(nranks + 1)){
            if (scratch) cudaFree(scratch);
            scratch_bytes = 64 * (nranks + 1);
            if (cudaMalloc(&scratch, scratch_bytes) != cudaSuccess){ scratch = nullptr; scratch_bytes = 0; return 1; }
        }
        char *send = static_cast<char*>(scratch);
        if (nccl().AllGather(send, send + 1, 1, nccl_int8, comm, stream) != nccl_success) return 2;
        return 0;
    }
    const char* kind() const override { return "nccl"; }
    // one process per GPU on one node: CUDA IPC handles travel through the allgather, peers are opened with peer access
    bool map_peers(void *local, size_t bytes, std::vector<void*> &peers) override {
        (void) bytes;
        peers.assign(nranks, nullptr);
        const char *off = std::getenv("HEFFTE_B200_DISABLE_P2P");
        struct record { cudaIpcMemHandle_t handle; int ok; int device; char host[64]; };
        record mine{};
        mine.ok = (off == nullptr or off[0] == '0') ? 1 : 0;
        if (mine.ok and cudaIpcGetMemHandle(&mine.handle, local) != cudaSuccess){ mine.ok = 0; cudaGetLastError(); }
        cudaGetDevice(&mine.device);
        gethostname_safe(mine.host, sizeof(mine.host));
        std::vector<record> all(nranks);
        if (allgather(&mine, all.data(), sizeof(record)) != 0) return false;
        bool usable = true;
        for(int r=0; r<nranks; r++) usable = usable and all[r].ok and std::strncmp(all[r].host, mine.host, sizeof(mine.host)) == 0;
        int opened = usable ? 1 : 0;
        if (usable){
            for(int r=0; r<nranks; r++){
                if (r == my_rank){ peers[r] = local; continue; }
                if (cudaIpcOpenMemHandle(&peers[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){
                    cudaGetLastError(); peers[r] = nullptr; opened = 0; break;
                }
            }
        }
        // everybody must agree
        std::vector<int> votes(nranks);
        if (allgather(&opened, votes.data(), sizeof(int)) != 0) opened = 0;
        for(int v : votes) if (not v) opened = 0;
        if (not opened){ unmap_peers(peers); peers.clear(); return false; }
        return true;
    }
    void unmap_peers(std::vector<void*> const &peers) override {
        for(int r=0; r<static_cast<int>(peers.size()); r++)
            if (r != my_rank and peers[r] != nullptr) cudaIpcCloseMemHandle(peers[r]);
    }
    bool map_user_buffer(void *ptr, std::vector<void*> &peers) override {
        peers.assign(nranks, nullptr);
        struct record { cudaIpcMemHandle_t handle; unsigned long long offset; int ok; int device; char host[64]; };
        record mine{};
        const char *off = std::getenv("HEFFTE_B200_DISABLE_P2P");
        void *base = nullptr;
        if ((off == nullptr or off[0] == '0') and ptr != nullptr and allocation_base(ptr, &base)){
            if (cudaIpcGetMemHandle(&mine.handle, base) == cudaSuccess){
                mine.ok = 1;
                mine.offset = static_cast<unsigned long long>(static_cast<char*>(ptr) - static_cast<char*>(base));
            }else cudaGetLastError();
        }
        cudaGetDevice(&mine.device);
        gethostname_safe(mine.host, sizeof(mine.host));
        std::vector<record> all(nranks);
        if (allgather(&mine, all.data(), sizeof(record)) != 0) return false;
        int opened = 1;
        for(int r=0; r<nranks; r++) if (not all[r].ok or std::strncmp(all[r].host, mine.host, sizeof(mine.host)) != 0) opened = 0;
        for(int r=0; r<nranks and opened; r++){
            if (r == my_rank){ peers[r] = ptr; continue; }
            void *remote = open_once(all[r].handle);
            if (remote == nullptr){ opened = 0; break; }
            peers[r] = static_cast<char*>(remote) + all[r].offset;
        }
        std::vector<int> votes(nranks);
        if (allgather(&opened, votes.data(), sizeof(int)) != 0) opened = 0;
        for(int v : votes) if (not v) opened = 0;
        if (not opened) peers.clear();
        return opened != 0;
    }
private:
    // the allocation a device pointer lives in (a caching allocator hands out pieces of large allocations; CUDA IPC exports whole ones)
    static bool allocation_base(void *ptr, void **base){
#ifdef B200_HOST_EMULATION
        *base = ptr;
        return true;
#else
        typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
        static range_fn const fn = []() -> range_fn {
            void *f = nullptr;
            cudaDriverEntryPointQueryResult status;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &status) != cudaSuccess or status != cudaDriverEntryPointSuccess){
                cudaGetLastError();
                return nullptr;
            }
            return reinterpret_cast<range_fn>(f);
        }();
        if (fn == nullptr) return false;
        unsigned long long start = 0;
        size_t size = 0;
        if (fn(&start, &size, static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(ptr))) != 0) return false;
        *base = reinterpret_cast<void*>(static_cast<uintptr_t>(start));
        return true;
#endif
    }
    // an IPC handle can be opened once per process: the mappings are kept for the life of the process
    static void* open_once(cudaIpcMemHandle_t const &handle){
        static std::mutex guard;
        static std::map<std::string, void*> opened;
        std::string const key(reinterpret_cast<const char*>(&handle), sizeof(handle));
        std::lock_guard<std::mutex> lock(guard);
        auto it = opened.find(key);
        if (it != opened.end()) return it->second;
        void *remote = nullptr;
        if (cudaIpcOpenMemHandle(&remote, handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){ cudaGetLastError(); return nullptr; }
        opened[key] = remote;
        return remote;
    }
    static void gethostname_safe(char *out, size_t n){
        std::memset(out, 0, n);
        FILE *f = std::fopen("/proc/sys/kernel/random/boot_id", "r");   // same kernel instance == same node (containers share it)
        if (f){ if (std::fgets(out, static_cast<int>(n), f) == nullptr) out[0] = 0; std::fclose(f); }
    }
    nccl_comm_t comm = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaStream_t side = nullptr;     // plan-time collectives
};

class callback_communicator : public communicator {
public:
    callback_communicator(int rank, int size, allgather_callback g, exchange_callback e, void *c) : gather(g), swap(e), context(c){ my_rank = rank; nranks = size; }
    int allgather(const void *mine, void *all, size_t bytes) override { return gather ? gather(context, mine, all, bytes) : 1; }
    int exchange(std::vector<transfer> const &sends, std::vector<transfer> const &recvs, cudaStream_t stream) override {
        if (swap == nullptr) return 1;
        std::vector<int> sp, rp; std::vector<void*> sd, rd; std::vector<size_t> sb, rb;
        for(auto const &s : sends){ sp.push_back(s.peer); sd.push_back(s.data); sb.push_back(s.bytes); }
        for(auto const &r : recvs){ rp.push_back(r.peer); rd.push_back(r.data); rb.push_back(r.bytes); }
        return swap(context, (int) sp.size(), sp.data(), sd.data(), sb.data(), (int) rp.size(), rp.data(), rd.data(), rb.data(), stream);
    }
    int barrier(cudaStream_t) override { return 0; }
    const char* kind() const override { return "callback"; }
private:
    allgather_callback gather;
    exchange_callback swap;
    void *context;
};

// ---- ranks as host threads of one process ------------------------------------------------------------------------------
struct thread_group {
    int size = 0;
    std::mutex guard;
    std::condition_variable cv;
    int waiting = 0;
    long long generation = 0;
    std::vector<const void*> slot;                 // per rank: pointer published for the current collective
    struct posted { std::vector<transfer> sends; cudaEvent_t ready = nullptr, done = nullptr; int device = 0; };
    std::vector<posted> box;

    void sync(){
        std::unique_lock<std::mutex> lock(guard);
        long long const mine = generation;
        if (++waiting == size){ waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lock, [&]{ return generation != mine; });
    }
};

class thread_communicator : public communicator {
public:
    thread_communicator(int rank, std::shared_ptr<thread_group> g, int dev) : group(g), device(dev){ my_rank = rank; nranks = g->size; }
    ~thread_communicator() override {
        thread_group::posted &mine = group->box[my_rank];
        if (mine.ready){ cudaEventDestroy(mine.ready); mine.ready = nullptr; }
        if (mine.done){ cudaEventDestroy(mine.done); mine.done = nullptr; }
    }
    int allgather(const void *mine, void *all, size_t bytes) override {
        group->slot[my_rank] = mine;
        group->sync();
        for(int r=0; r<nranks; r++) std::memcpy(static_cast<char*>(all) + r * bytes, group->slot[r], bytes);
        group->sync();
        return 0;
    }
    int exchange(std::vector<transfer> const &sends, std::vector<transfer> const &recvs, cudaStream_t stream) override {
        cudaSetDevice(device);
        thread_group::posted &mine = group->box[my_rank];
        if (mine.ready == nullptr){
            if (cudaEventCreateWithFlags(&mine.ready, cudaEventDisableTiming) != cudaSuccess) return 1;
            if (cudaEventCreateWithFlags(&mine.done, cudaEventDisableTiming) != cudaSuccess) return 1;
        }
        mine.sends = sends;
        mine.device = device;
        if (cudaEventRecord(mine.ready, stream) != cudaSuccess) return 1;      // my messages are packed
        group->sync();
        int status = 0;
        for(auto const &r : recvs){
            thread_group::posted &from = group->box[r.peer];
            const transfer *match = nullptr;
            for(auto const &s : from.sends) if (s.peer == my_rank){ match = &s; break; }
            if (match == nullptr or match->bytes != r.bytes){ status = 1; continue; }
            if (cudaStreamWaitEvent(stream, from.ready, 0) != cudaSuccess) status = 1;
            if (r.bytes > 0 and cudaMemcpyPeerAsync(r.data, device, match->data, from.device, r.bytes, stream) != cudaSuccess) status = 1;
        }
        if (cudaEventRecord(mine.done, stream) != cudaSuccess) status = 1;     // I have pulled everything addressed to me
        group->sync();
        for(auto const &s : sends)                                              // my send buffer is free once the receivers are done
            if (cudaStreamWaitEvent(stream, group->box[s.peer].done, 0) != cudaSuccess) status = 1;
        group->sync();
        return status;
    }
    int barrier(cudaStream_t) override { group->sync(); return 0; }
    const char* kind() const override { return "threads"; }
    void after_peer_barrier() override { group->sync(); }
    void before_peer_barrier() override { group->sync(); }
    bool map_peers(void *local, size_t bytes, std::vector<void*> &peers) override {
        (void) bytes;
        const char *off = std::getenv("HEFFTE_B200_DISABLE_P2P");
        struct record { void *ptr; int device; int ok; };
        record mine{local, device, (off == nullptr or off[0] == '0') ? 1 : 0};
        std::vector<record> all(nranks);
        allgather(&mine, all.data(), sizeof(record));
        int usable = 1;
        cudaSetDevice(device);
        for(int r=0; r<nranks and usable; r++){
            if (not all[r].ok) usable = 0;
            if (all[r].device != device){
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, device, all[r].device) != cudaSuccess or not can) usable = 0;
                else{
                    cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
                    if (e != cudaSuccess and e != cudaErrorPeerAccessAlreadyEnabled) usable = 0;
                    cudaGetLastError();
                }
            }
        }
        std::vector<int> votes(nranks);
        allgather(&usable, votes.data(), sizeof(int));
        for(int v : votes) if (not v) usable = 0;
        peers.clear();
        if (not usable) return false;
        for(int r=0; r<nranks; r++) peers.push_back(all[r].ptr);
        return true;
    }
    bool map_user_buffer(void *ptr, std::vector<void*> &peers) override {
        struct record { void *ptr; int ok; };
        const char *off = std::getenv("HEFFTE_B200_DISABLE_P2P");
        record mine{ptr, (ptr != nullptr and (off == nullptr or off[0] == '0')) ? 1 : 0};
        std::vector<record> all(nranks);
        allgather(&mine, all.data(), sizeof(record));
        peers.clear();
        for(int r=0; r<nranks; r++) if (not all[r].ok) return false;
        for(int r=0; r<nranks; r++) peers.push_back(all[r].ptr);
        return true;
    }
private:
    std::shared_ptr<thread_group> group;
    int device;
};

} // namespace

std::vector<communicator*> make_thread_communicators(int size, const int *devices, std::string &error){
    std::vector<communicator*> out;
    if (size < 1){ error = "bad group size"; return out; }
    auto group = std::make_shared<thread_group>();
    group->size = size;
    group->slot.assign(size, nullptr);
    group->box.resize(size);
    for(int r=0; r<size; r++) out.push_back(new thread_communicator(r, group, devices ? devices[r] : 0));
    return out;
}

communicator* make_self_communicator(){ return new self_communicator(); }

int nccl_unique_id(void *out128, std::string &error){
    nccl_api &api = nccl();
    if (not api.problem.empty()){ error = api.problem; return 1; }
    nccl_uid id;
    int rc = api.GetUniqueId(&id);
    if (rc != nccl_success){ error = std::string("ncclGetUniqueId: ") + api.GetErrorString(rc); return 1; }
    std::memcpy(out128, id.internal, 128);
    return 0;
}

communicator* make_nccl_communicator(int rank, int size, const void *unique_id, std::string &error){
    nccl_api &api = nccl();
    if (not api.problem.empty()){ error = api.problem; return nullptr; }
    nccl_uid id;
    std::memcpy(id.internal, unique_id, 128);
    nccl_comm_t comm = nullptr;
    int rc = api.CommInitRank(&comm, size, id, rank);
    if (rc != nccl_success){ error = std::string("ncclCommInitRank: ") + api.GetErrorString(rc); return nullptr; }
    return new nccl_communicator(rank, size, comm);
}

communicator* make_callback_communicator(int rank, int size, allgather_callback gather, exchange_callback exchange, void *context){
    return new callback_communicator(rank, size, gather, exchange, context);
}

} // namespace b200
