#include "comm.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
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
        char *send = static_cast<char*>(scratch), *recv = send + bytes;
        if (cudaMemcpy(send, mine, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        if (nccl().AllGather(send, recv, bytes, nccl_int8, comm, nullptr) != nccl_success) return 2;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) return 1;
        if (cudaMemcpy(all, recv, bytes * nranks, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
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
private:
    nccl_comm_t comm = nullptr;
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
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

} // namespace

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
