// Peer-memory plumbing for the fused data-parallel PPO kernel: one process per GPU, receive buffers allocated with
// cudaMalloc and shared with the other ranks of the node through CUDA IPC handles (exchanged by the host side over
// torch.distributed).  Peers store straight into these buffers over NVLink from inside the persistent kernel.
#include <string.h>

#include "common.cuh"

extern "C" {

int icrl_comm_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
    ICRL_CHECK_ARG(bytes > 0 && dev_ptr && handle64, "icrl_comm_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    ICRL_CUDA(cudaMalloc(&p, (size_t)bytes));
    ICRL_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    ICRL_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return 0;
}

int icrl_comm_open(const unsigned char* handle64, void** dev_ptr) {
    ICRL_CHECK_ARG(handle64 && dev_ptr, "icrl_comm_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    ICRL_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int icrl_comm_close(void* dev_ptr) {
    if (dev_ptr) ICRL_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int icrl_comm_free(void* dev_ptr) {
    if (dev_ptr) ICRL_CUDA(cudaFree(dev_ptr));
    return 0;
}

}  // extern "C"
