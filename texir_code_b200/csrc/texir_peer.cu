// texir_peer.cu -- peer-memory plumbing of the multi-GPU path: one process per GPU, every rank maps the
// other ranks' result buffers over NVLink / NVSwitch (CUDA IPC) so that the compute kernels' epilogues store straight
// into every replica (texir_irradiance_fwd_peers: the IrT all-gather is fused into the kernel that produces the
// values), plus the flag barrier that orders those stores between ranks.  No NCCL on this path.
//
//   texir_peer_alloc / _free      cudaMalloc'd buffer + its 64-byte IPC handle (rank-local owner)
//   texir_peer_open / _close      map a buffer another rank exported
//   texir_peer_barrier            every rank stores `epoch` into its slot of every rank's flag array
//                                 (st.release.sys) and spins on its own array (ld.acquire.sys); bounded wait
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/texir_b200.h"
#include "texir_host.h"

using texir_host::fail;
using texir_host::finish_launch;

#define PEER_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_), 2);   \
    } while (0)

struct PeerFlags {
    uint32_t* p[TEXIR_MAX_PEERS];
};

// flags[r] = rank r's array of world + 1 words: slot k < world is written by rank k, slot `world` is the timeout mark
__global__ void k_peer_barrier(PeerFlags f, int rank, int world, uint32_t epoch, long long limit_cycles) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[t] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = f.p[rank] + t;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > limit_cycles) {  // a peer never arrived: mark it instead of hanging the GPU
            f.p[rank][world] = epoch;
            break;
        }
        __nanosleep(64);
    }
}

extern "C" {

int texir_peer_alloc(int device, int64_t bytes, void** dptr, unsigned char* handle_out) {
    if (!dptr || !handle_out || bytes <= 0) return fail("texir_peer_alloc: bad argument");
    texir_host::ScopedDevice guard(device);
    void* p = nullptr;
    PEER_CUDA(cudaMalloc(&p, (size_t)bytes));
    PEER_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(std::string("texir_peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(e), 2);
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == TEXIR_PEER_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof(h));
    *dptr = p;
    return 0;
}

int texir_peer_free(int device, void* dptr) {
    if (!dptr) return 0;
    texir_host::ScopedDevice guard(device);
    PEER_CUDA(cudaFree(dptr));
    return 0;
}

int texir_peer_open(int device, const unsigned char* handle, void** dptr) {
    if (!handle || !dptr) return fail("texir_peer_open: bad argument");
    texir_host::ScopedDevice guard(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    PEER_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return 0;
}

int texir_peer_close(int device, void* dptr) {
    if (!dptr) return 0;
    texir_host::ScopedDevice guard(device);
    PEER_CUDA(cudaIpcCloseMemHandle(dptr));
    return 0;
}

int texir_peer_barrier(uint32_t* const* flags_host, int rank, int world, uint32_t epoch, int device, void* stream) {
    if (!flags_host || world <= 0 || world > TEXIR_MAX_PEERS || rank < 0 || rank >= world)
        return fail("texir_peer_barrier: bad argument");
    texir_host::ScopedDevice guard(device);
    PeerFlags f{};
    for (int k = 0; k < world; ++k) {
        if (!flags_host[k]) return fail("texir_peer_barrier: NULL flag array");
        f.p[k] = flags_host[k];
    }
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    const long long limit = (long long)(khz > 0 ? khz : 2000000) * 1000ll * 4ll;  // about 4 s of SM clocks
    k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, epoch, limit);
    return finish_launch("k_peer_barrier");
}

}  // extern "C"
