// Peer-memory all-gather of per-design rows (losses, gradients) for the multi-GPU path: one process per GPU, every
// rank owns a buffer that its peers have mapped through CUDA IPC, and ONE kernel stores the rank's rows into all of
// them over NVLink / NVSwitch (plain st.global to peer addresses, 16 bytes per thread and destination).  The reference
// has no multi-process path at all (src/dfdm/optimization.py:89-97 evaluates one design per call); this is the one
// exchange SURVEY.md section 8(e) allows per batched evaluation, and it replaces ncclAllGather in bench.py because the
// stores of all ranks proceed concurrently at full NVSwitch bandwidth without a staging protocol.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace dfdm {

constexpr int kPushMaxDst = 16;
struct PushDst { void* base[kPushMaxDst]; };

__global__ void __launch_bounds__(256) k_peer_push(const uint4* __restrict__ src, int64_t n16, int64_t dst_off16, PushDst dst, int n_dst,
                                                   const double* __restrict__ tail_src, int64_t tail8, int64_t tail_off8) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = src[i];
        for (int d = 0; d < n_dst; ++d) reinterpret_cast<uint4*>(dst.base[d])[dst_off16 + i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < tail8) {                     // bytes % 16 == 8: one last double
        const double v = tail_src[threadIdx.x];
        for (int d = 0; d < n_dst; ++d) reinterpret_cast<double*>(dst.base[d])[tail_off8 + threadIdx.x] = v;
    }
}

}  // namespace dfdm

extern "C" {

int32_t dfdm_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle) {
    using namespace dfdm;
    if (bytes <= 0 || !ptr || !handle) {
        set_error("dfdm_peer_alloc: non-positive size or NULL outputs");
        return DFDM_EINVAL;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    DFDM_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t err = cudaIpcGetMemHandle(&h, p);
    if (err != cudaSuccess) {
        cudaFree(p);
        set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(err));
        return DFDM_ECUDA;
    }
    std::memcpy(handle, &h, 64);
    *ptr = p;
    return DFDM_OK;
}

int32_t dfdm_peer_open(const uint8_t* handle, void** ptr) {
    using namespace dfdm;
    if (!handle || !ptr) {
        set_error("dfdm_peer_open: NULL handle or output");
        return DFDM_EINVAL;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    DFDM_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return DFDM_OK;
}

int32_t dfdm_peer_close(void* ptr) {
    using namespace dfdm;
    if (ptr) DFDM_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return DFDM_OK;
}

int32_t dfdm_peer_free(void* ptr) {
    using namespace dfdm;
    if (ptr) DFDM_CUDA_OK(cudaFree(ptr));
    return DFDM_OK;
}

int32_t dfdm_peer_push(const void* src, int64_t bytes, int64_t dst_offset_bytes, void* const* dst_bases, int32_t n_dst, void* stream) {
    using namespace dfdm;
    if (!src || bytes < 0 || dst_offset_bytes < 0 || !dst_bases || n_dst <= 0 || n_dst > kPushMaxDst) {
        set_error("dfdm_peer_push: NULL buffers, negative sizes or more than 16 destinations");
        return DFDM_EINVAL;
    }
    if ((bytes % 8) != 0 || (dst_offset_bytes % 16) != 0 || (reinterpret_cast<uintptr_t>(src) % 16) != 0) {
        set_error("dfdm_peer_push: size must be a multiple of 8 bytes, source and destination offset 16-byte aligned");
        return DFDM_EINVAL;
    }
    if (bytes == 0) return DFDM_OK;
    PushDst dst{};
    for (int d = 0; d < n_dst; ++d) {
        if (!dst_bases[d]) {
            set_error("dfdm_peer_push: NULL destination");
            return DFDM_EINVAL;
        }
        dst.base[d] = dst_bases[d];
    }
    const int64_t n16 = bytes / 16, tail8 = (bytes % 16) / 8;
    int dev = 0, sms = 0;
    DFDM_CUDA_OK(cudaGetDevice(&dev));
    DFDM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t want = (n16 + 255) / 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    k_peer_push<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), n16, dst_offset_bytes / 16, dst, n_dst,
                                                        reinterpret_cast<const double*>(src) + n16 * 2, tail8,
                                                        dst_offset_bytes / 8 + n16 * 2);
    DFDM_LAUNCH_OK("k_peer_push");
    return DFDM_OK;
}

}  // extern "C"
