"""Probe: does CUDA IPC peer mapping work between the per-GPU processes of one box, and how fast is a peer write?
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ipc_probe.py"""
import os
import torch
import torch.distributed as dist
from cuda.bindings import runtime as rt

rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 256 << 20
err, ptr = rt.cudaMalloc(nbytes)
assert err == rt.cudaError_t.cudaSuccess, err
err, handle = rt.cudaIpcGetMemHandle(ptr)
print(rank, "ipc get:", err, flush=True)
handles = [None] * size
dist.all_gather_object(handles, bytes(handle.reserved))
peer = (rank + 1) % size
h = rt.cudaIpcMemHandle_t()
h.reserved = handles[peer]
err, pptr = rt.cudaIpcOpenMemHandle(h, rt.cudaIpcMemLazyEnablePeerAccess)
print(rank, "ipc open:", err, "can_access_peer:", torch.cuda.can_device_access_peer(local, peer), flush=True)
if err == rt.cudaError_t.cudaSuccess:
    err, s = rt.cudaStreamCreate()
    e0, e1 = rt.cudaEventCreate()[1], rt.cudaEventCreate()[1]
    for _ in range(3):
        rt.cudaMemcpyAsync(pptr, ptr, nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, s)
    rt.cudaStreamSynchronize(s)
    dist.barrier()
    rt.cudaEventRecord(e0, s)
    for _ in range(10):
        rt.cudaMemcpyAsync(pptr, ptr, nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, s)
    rt.cudaEventRecord(e1, s)
    rt.cudaStreamSynchronize(s)
    ms = rt.cudaEventElapsedTime(e0, e1)[1] / 10
    print(rank, "peer write (all ranks concurrently): %.1f GB/s" % (nbytes / ms * 1e-6), flush=True)
dist.barrier()
dist.destroy_process_group()
