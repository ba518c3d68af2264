// peer.cu — peer-visible device memory for the fused covariance all-gather (gprb_kff_multi / gprb_kfe_multi).
//
// One process per GPU on one NVLink / NVSwitch node.  Every process allocates its copy of K here, exports a
// CUDA IPC handle, and maps the other processes' copies; the covariance kernels then store finished values
// into all copies (replaces the pickle gather + bcast of RBF_mb.py:471-521).  The handles travel through
// whatever the host side uses for rendezvous (dist.py: torch.distributed all_gather).
#include "common.cuh"
#include <cstring>

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");

extern "C" int gprb_peer_alloc(void **ptr, unsigned long long bytes) {
    GPRB_REQUIRE(ptr && bytes > 0, "gprb_peer_alloc: bad argument");
    *ptr = nullptr;
    GPRB_CUDA(cudaMalloc(ptr, (size_t)bytes));
    return GPRB_OK;
}

extern "C" int gprb_peer_free(void *ptr) {
    if (ptr) GPRB_CUDA(cudaFree(ptr));
    return GPRB_OK;
}

extern "C" int gprb_peer_export(void *ptr, unsigned char *handle64) {
    GPRB_REQUIRE(ptr && handle64, "gprb_peer_export: NULL argument");
    cudaIpcMemHandle_t h;
    GPRB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    std::memcpy(handle64, &h, sizeof h);
    return GPRB_OK;
}

extern "C" int gprb_peer_open(const unsigned char *handle64, void **ptr) {
    GPRB_REQUIRE(ptr && handle64, "gprb_peer_open: NULL argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof h);
    *ptr = nullptr;
    GPRB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GPRB_OK;
}

extern "C" int gprb_peer_close(void *ptr) {
    if (ptr) GPRB_CUDA(cudaIpcCloseMemHandle(ptr));
    return GPRB_OK;
}
