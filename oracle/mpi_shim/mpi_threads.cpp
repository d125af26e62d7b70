/*
 * TEST INFRASTRUCTURE ONLY -- implementation of the threads-as-ranks MPI stand-in declared in mpi.h.
 * See mpi.h for why it exists.  Ranks are std::threads; collectives rendezvous through a
 * per-communicator slot table guarded by a generation barrier; point-to-point messages are
 * buffered eagerly in per-destination mailboxes.
 */
#include "mpi.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include <dlfcn.h>

namespace {

// "GPU-aware MPI": when the program is linked with the b200 library the buffers may be device pointers.  The copy goes through
// the library's own b200_copy_any (cudaMemcpy with cudaMemcpyDefault, found at run time); pure host programs use memcpy.
void copy_bytes(void *dst, const void *src, size_t bytes){
    using copy_fn = int (*)(void*, const void*, size_t);
    static copy_fn const device_copy = reinterpret_cast<copy_fn>(dlsym(RTLD_DEFAULT, "b200_copy_any"));
    static std::atomic<bool> usable{true};      // the library can be present in a process without a GPU (the CPU test-suite)
    if (bytes == 0) return;
    if (device_copy != nullptr and usable.load()){
        if (device_copy(dst, src, bytes) == 0) return;
        usable.store(false);
    }
    std::memcpy(dst, src, bytes);
}

inline size_t type_bytes(MPI_Datatype t){ return static_cast<size_t>(t & 0xff); }

struct message {
    int src, tag;
    std::vector<char> payload;
};

struct mailbox {
    std::mutex lock;
    std::condition_variable cv;
    std::deque<message> inbox;
};

} // namespace

struct shim_comm_s {
    int size = 1;
    // generation barrier
    std::mutex lock;
    std::condition_variable cv;
    int waiting = 0;
    long generation = 0;
    // rendezvous slots, one per rank
    std::vector<const void*> slot;
    std::vector<std::unique_ptr<mailbox>> boxes;
    int refs = 0; // number of ranks that still hold the communicator
    explicit shim_comm_s(int n) : size(n), slot(n, nullptr), refs(n) {
        for(int i=0; i<n; i++) boxes.emplace_back(new mailbox());
    }
    void barrier(){
        if (size == 1) return;
        std::unique_lock<std::mutex> guard(lock);
        long gen = generation;
        if (++waiting == size){
            waiting = 0;
            generation++;
            cv.notify_all();
        }else{
            cv.wait(guard, [&]{ return gen != generation; });
        }
    }
};

struct shim_group_s { std::vector<int> ranks; };

struct shim_request_s {
    bool is_recv = false;
    bool done = false;
    void *buf = nullptr;
    size_t bytes = 0;
    int src = 0, tag = 0;
    MPI_Comm comm = nullptr;
    int my_rank = 0;
};

namespace {

struct rank_binding { MPI_Comm comm; int rank; };
thread_local std::vector<rank_binding> tl_bindings; // (communicator -> my rank in it)
thread_local int tl_world_rank = 0;

std::mutex world_lock;
MPI_Comm world = nullptr;

int rank_in(MPI_Comm comm){
    if (comm == world) return tl_world_rank;
    for(auto const &b : tl_bindings) if (b.comm == comm) return b.rank;
    return 0;
}

bool try_match(shim_request_s *r){
    mailbox &mb = *r->comm->boxes[r->my_rank];
    // caller holds mb.lock
    for(auto it = mb.inbox.begin(); it != mb.inbox.end(); ++it){
        if (it->src == r->src and (r->tag == MPI_ANY_TAG or it->tag == r->tag)){
            copy_bytes(r->buf, it->payload.data(), std::min(r->bytes, it->payload.size()));
            mb.inbox.erase(it);
            r->done = true;
            return true;
        }
    }
    return false;
}

} // namespace

extern "C" {

MPI_Comm shim_comm_world(void){
    std::lock_guard<std::mutex> guard(world_lock);
    if (world == nullptr) world = new shim_comm_s(1);
    return world;
}

int shim_world_size_from_env(void){
    const char *e = std::getenv("SHIM_NP");
    int n = (e == nullptr) ? 1 : std::atoi(e);
    return (n < 1) ? 1 : n;
}

int shim_run(int nranks, int (*fn)(int, char**), int argc, char **argv){
    {
        std::lock_guard<std::mutex> guard(world_lock);
        delete world;
        world = new shim_comm_s(nranks);
    }
    std::vector<int> rc(nranks, 0);
    std::vector<std::thread> pool;
    for(int r=0; r<nranks; r++){
        pool.emplace_back([&, r]{
            tl_world_rank = r;
            tl_bindings.clear();
            rc[r] = fn(argc, argv);
        });
    }
    for(auto &t : pool) t.join();
    int worst = 0;
    for(int r : rc) if (r != 0) worst = r;
    return worst;
}

int MPI_Init(int*, char***){ (void) shim_comm_world(); return MPI_SUCCESS; }
int MPI_Finalize(void){ return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code){ std::fprintf(stderr, "MPI_Abort(%d) in the oracle shim\n", code); std::exit(code); }
double MPI_Wtime(void){
    using clock = std::chrono::steady_clock;
    static const clock::time_point origin = clock::now();
    return std::chrono::duration<double>(clock::now() - origin).count();
}

int MPI_Comm_rank(MPI_Comm comm, int *rank){ *rank = rank_in(comm); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size){ *size = comm->size; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm comm){ comm->barrier(); return MPI_SUCCESS; }

int MPI_Comm_free(MPI_Comm *comm){
    if (comm == nullptr or *comm == nullptr or *comm == world) return MPI_SUCCESS;
    MPI_Comm c = *comm;
    for(auto it = tl_bindings.begin(); it != tl_bindings.end(); ++it){
        if (it->comm == c){ tl_bindings.erase(it); break; }
    }
    bool last;
    { std::lock_guard<std::mutex> guard(c->lock); last = (--c->refs == 0); }
    if (last) delete c;
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group){
    auto *g = new shim_group_s();
    for(int i=0; i<comm->size; i++) g->ranks.push_back(i);
    *group = g;
    return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup){
    auto *g = new shim_group_s();
    for(int i=0; i<n; i++) g->ranks.push_back(group->ranks[ranks[i]]);
    *newgroup = g;
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *group){ delete *group; *group = nullptr; return MPI_SUCCESS; }

int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm){
    // collective over comm; the groups passed by different ranks are identical or disjoint
    int const me = rank_in(comm);
    int my_pos = -1;
    for(size_t i=0; i<group->ranks.size(); i++) if (group->ranks[i] == me) my_pos = static_cast<int>(i);
    MPI_Comm created = nullptr;
    if (my_pos == 0) created = new shim_comm_s(static_cast<int>(group->ranks.size()));
    comm->slot[me] = created;
    comm->barrier();
    MPI_Comm result = MPI_COMM_NULL;
    if (my_pos >= 0){
        result = static_cast<MPI_Comm>(const_cast<void*>(comm->slot[group->ranks[0]]));
        tl_bindings.push_back({result, my_pos});
    }
    comm->barrier();
    *newcomm = result;
    return MPI_SUCCESS;
}

int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                  void *recvbuf, int, MPI_Datatype, MPI_Comm comm){
    int const me = rank_in(comm);
    size_t const bytes = sendcount * type_bytes(sendtype);
    comm->slot[me] = sendbuf;
    comm->barrier();
    for(int r=0; r<comm->size; r++)
        std::memcpy(static_cast<char*>(recvbuf) + r * bytes, comm->slot[r], bytes);
    comm->barrier();
    return MPI_SUCCESS;
}

int MPI_Alltoall(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                 void *recvbuf, int, MPI_Datatype, MPI_Comm comm){
    int const me = rank_in(comm);
    size_t const bytes = sendcount * type_bytes(sendtype);
    comm->slot[me] = sendbuf;
    comm->barrier();
    for(int r=0; r<comm->size; r++)
        copy_bytes(static_cast<char*>(recvbuf) + r * bytes,
                   static_cast<const char*>(comm->slot[r]) + me * bytes, bytes);
    comm->barrier();
    return MPI_SUCCESS;
}

namespace { struct a2av_args { const void *buf; const int *counts; const int *displs; size_t tsize; }; }

int MPI_Alltoallv(const void *sendbuf, const int sendcounts[], const int sdispls[], MPI_Datatype sendtype,
                  void *recvbuf, const int recvcounts[], const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm){
    int const me = rank_in(comm);
    a2av_args mine = {sendbuf, sendcounts, sdispls, type_bytes(sendtype)};
    comm->slot[me] = &mine;
    comm->barrier();
    size_t const rsize = type_bytes(recvtype);
    for(int r=0; r<comm->size; r++){
        auto const *peer = static_cast<const a2av_args*>(comm->slot[r]);
        size_t const bytes = std::min<size_t>(peer->counts[me] * peer->tsize, recvcounts[r] * rsize);
        if (bytes > 0)
            copy_bytes(static_cast<char*>(recvbuf) + rdispls[r] * rsize,
                       static_cast<const char*>(peer->buf) + peer->displs[me] * peer->tsize, bytes);
    }
    comm->barrier();
    return MPI_SUCCESS;
}

static void reduce_into(void *acc, const void *x, int count, MPI_Datatype type, MPI_Op op){
    auto apply = [&](auto *a, auto const *b){
        for(int i=0; i<count; i++){
            if (op == MPI_MAX) a[i] = std::max(a[i], b[i]);
            else if (op == MPI_MIN) a[i] = std::min(a[i], b[i]);
            else a[i] = a[i] + b[i];
        }
    };
    if (type == MPI_DOUBLE) apply(static_cast<double*>(acc), static_cast<const double*>(x));
    else if (type == MPI_FLOAT) apply(static_cast<float*>(acc), static_cast<const float*>(x));
    else if (type == MPI_INT) apply(static_cast<int*>(acc), static_cast<const int*>(x));
    else if (type == MPI_LONG_LONG) apply(static_cast<long long*>(acc), static_cast<const long long*>(x));
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm){
    int const me = rank_in(comm);
    size_t const bytes = count * type_bytes(type);
    std::vector<char> mine(static_cast<const char*>(sendbuf), static_cast<const char*>(sendbuf) + bytes);
    comm->slot[me] = mine.data();
    comm->barrier();
    std::vector<char> acc(static_cast<const char*>(comm->slot[0]), static_cast<const char*>(comm->slot[0]) + bytes);
    for(int r=1; r<comm->size; r++) reduce_into(acc.data(), comm->slot[r], count, type, op);
    comm->barrier();
    std::memcpy(recvbuf, acc.data(), bytes);
    return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm){
    std::vector<char> tmp(count * type_bytes(type));
    MPI_Allreduce(sendbuf, tmp.data(), count, type, op, comm);
    if (rank_in(comm) == root) std::memcpy(recvbuf, tmp.data(), tmp.size());
    return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm){
    message m;
    m.src = rank_in(comm);
    m.tag = tag;
    size_t const bytes = count * type_bytes(type);
    m.payload.resize(bytes);
    copy_bytes(m.payload.data(), buf, bytes);
    mailbox &mb = *comm->boxes[dest];
    {
        std::lock_guard<std::mutex> guard(mb.lock);
        mb.inbox.push_back(std::move(m));
    }
    mb.cv.notify_all();
    return MPI_SUCCESS;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *request){
    MPI_Send(buf, count, type, dest, tag, comm); // eager, buffered
    auto *r = new shim_request_s();
    r->done = true;
    *request = r;
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *request){
    auto *r = new shim_request_s();
    r->is_recv = true;
    r->buf = buf;
    r->bytes = count * type_bytes(type);
    r->src = source;
    r->tag = tag;
    r->comm = comm;
    r->my_rank = rank_in(comm);
    *request = r;
    return MPI_SUCCESS;
}

int MPI_Waitany(int count, MPI_Request requests[], int *index, MPI_Status *status){
    // completed sends first
    for(int i=0; i<count; i++){
        if (requests[i] != MPI_REQUEST_NULL and requests[i]->done){
            if (status != MPI_STATUS_IGNORE){ status->MPI_SOURCE = requests[i]->src; status->MPI_TAG = requests[i]->tag; status->MPI_ERROR = 0; }
            delete requests[i];
            requests[i] = MPI_REQUEST_NULL;
            *index = i;
            return MPI_SUCCESS;
        }
    }
    // find the mailbox shared by the pending receives (all receives of one rank use its own box)
    shim_request_s *any = nullptr;
    for(int i=0; i<count; i++) if (requests[i] != MPI_REQUEST_NULL){ any = requests[i]; break; }
    if (any == nullptr){ *index = MPI_UNDEFINED; return MPI_SUCCESS; }
    mailbox &mb = *any->comm->boxes[any->my_rank];
    std::unique_lock<std::mutex> guard(mb.lock);
    while(true){
        for(int i=0; i<count; i++){
            if (requests[i] == MPI_REQUEST_NULL) continue;
            if (try_match(requests[i])){
                if (status != MPI_STATUS_IGNORE){ status->MPI_SOURCE = requests[i]->src; status->MPI_TAG = requests[i]->tag; status->MPI_ERROR = 0; }
                delete requests[i];
                requests[i] = MPI_REQUEST_NULL;
                *index = i;
                return MPI_SUCCESS;
            }
        }
        mb.cv.wait(guard);
    }
}

int MPI_Waitall(int count, MPI_Request requests[], MPI_Status[]){
    int remaining = 0;
    for(int i=0; i<count; i++) if (requests[i] != MPI_REQUEST_NULL) remaining++;
    while(remaining-- > 0){
        int idx;
        MPI_Waitany(count, requests, &idx, MPI_STATUS_IGNORE);
    }
    return MPI_SUCCESS;
}

} // extern "C"
