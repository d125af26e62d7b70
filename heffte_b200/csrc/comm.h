// Communicator abstraction of the b200 backend.
// The reference is hard-wired to MPI (MPI_Alltoallv / MPI_Send / MPI_Irecv in src/heffte_reshape3d.cpp:272, 402, 551-713;
// MPI_Allgather of the boxes in include/heffte_geometry.h:707-718).  Here one rank drives one GPU and the data path is
// NCCL grouped send/recv over NVLink, stream-ordered (no host synchronisation); the only host-side collective is the
// plan-time allgather of the boxes.  The NCCL library is resolved at run time (dlopen of libnccl.so.2) so that the
// copy already loaded by PyTorch is shared when the backend is used from Python.
#pragma once

#include "cuda_compat.h"

#include <cstddef>
#include <string>
#include <vector>

namespace b200 {

struct transfer {
    int peer;
    void *data;          // device pointer
    size_t bytes;
};

class communicator {
public:
    virtual ~communicator() = default;
    int rank() const { return my_rank; }
    int size() const { return nranks; }
    // host-to-host allgather of `bytes` per rank (plan time only); returns 0 on success
    virtual int allgather(const void *mine, void *all, size_t bytes) = 0;
    // one grouped exchange of device buffers on `stream` (stream ordered, asynchronous to the host)
    virtual int exchange(std::vector<transfer> const &sends, std::vector<transfer> const &recvs, cudaStream_t stream) = 0;
    virtual int barrier(cudaStream_t stream) = 0;
    virtual const char* kind() const = 0;
    // Peer memory (NVLink): COLLECTIVE.  Every rank passes one device allocation of `bytes`; on success peers[r] is an
    // address valid on THIS rank's device for rank r's allocation (peers[rank()] == local) and every rank returns true.
    // Returns false on every rank when peer mapping is not available (the plan then uses exchange()).
    virtual bool map_peers(void *local, size_t bytes, std::vector<void*> &peers){ (void) local; (void) bytes; peers.clear(); return false; }
    virtual void unmap_peers(std::vector<void*> const &peers){ (void) peers; }
    // Collective: the address of every rank's `ptr` -- device memory the CALLER allocated -- as seen from this device (peers[me] =
    // ptr).  One process per GPU: the allocation that holds ptr is opened through CUDA IPC (once per process, kept open).
    // False on every rank when any rank cannot share its memory (a pool the driver does not export, another node).
    virtual bool map_user_buffer(void *ptr, std::vector<void*> &peers){ (void) ptr; peers.clear(); return false; }
    // Called right after a peer barrier kernel has been enqueued.  Ranks that are host threads sharing one GPU rendezvous here:
    // kernels of different streams can share a hardware queue, and a spinning barrier kernel followed by dependent work of the
    // same stream would otherwise block the barrier kernel of another rank queued behind it.  One process per GPU: nothing to do.
    virtual void after_peer_barrier(){}
    // Called right before a peer barrier kernel is enqueued.  Ranks that are host threads sharing one GPU rendezvous here as
    // well: the first launch of a kernel loads its module lazily, which waits for the kernels running on the device -- a rank
    // without work in a stage (sub-communicator plans) would already spin in its barrier kernel, waiting for the very rank whose
    // launch cannot start.  After this rendezvous every rank has ENQUEUED its stage before any barrier kernel of the fence exists.
    virtual void before_peer_barrier(){}
protected:
    int my_rank = 0, nranks = 1;
};

// single-rank communicator (no communication library involved)
communicator* make_self_communicator();
// NCCL communicator from a 128-byte unique id shared by all ranks; binds to the current CUDA device
communicator* make_nccl_communicator(int rank, int size, const void *unique_id, std::string &error);
// fills 128 bytes; returns 0 on success
int nccl_unique_id(void *out128, std::string &error);

// N ranks inside ONE process, one host thread per rank (the single-process multi-GPU mode, and the way the test-suite runs
// multi-rank plans on a single GPU): rank r drives device devices[r] (repeats allowed).  Peer memory is direct (unified
// addressing + cudaDeviceEnablePeerAccess), exchange() is a stream-ordered device copy.  Returns `size` communicators that
// share one group; each must be used from its own host thread, collectively.
std::vector<communicator*> make_thread_communicators(int size, const int *devices, std::string &error);

// host-only communicator driven by callbacks (used by the CPU-side multi-process tests of the planning logic
// and by callers that own a different transport); exchange() is the caller's callback
typedef int (*allgather_callback)(void *context, const void *mine, void *all, size_t bytes);
typedef int (*exchange_callback)(void *context, int nsend, const int *send_peer, void *const *send_ptr, const size_t *send_bytes,
                                 int nrecv, const int *recv_peer, void *const *recv_ptr, const size_t *recv_bytes, void *stream);
communicator* make_callback_communicator(int rank, int size, allgather_callback gather, exchange_callback exchange, void *context);

} // namespace b200
