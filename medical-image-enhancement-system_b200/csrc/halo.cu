// halo.cu — mie_halo_exchange_z: the one exchange step of the path (BASELINE.json config 3).
//
// A z-slab of a sharded volume needs ONE neighbouring plane per interior face before its 3x3x3 median can
// finish.  The exchange is four NCCL point-to-point operations in one group on the CALLER's stream:
//     send first owned plane -> rank-1,  recv halo_lo <- rank-1,  send last owned plane -> rank+1,  recv halo_hi <- rank+1
// (two of them at the ends of the chain).  Nothing is allocated, nothing synchronises, and the call may be
// captured into a CUDA graph (NCCL point-to-point operations are capturable), which is how SlabPlan uses it.
//
// NCCL is NOT linked: the communicator comes from the host application (torch.distributed's ProcessGroupNCCL via
// _comm_ptr(), or an ncclCommInitRank of the caller's own), so the calls must go to the SAME libnccl that
// created it.  The entry points are therefore resolved at first use from the libnccl.so.2 already loaded into the
// process (dlopen RTLD_NOLOAD); without one the call returns MIE_E_UNSUPPORTED.
#include <dlfcn.h>
#include <cstring>

#include <mutex>

#include "mie_common.cuh"

namespace mie {

// Minimal restatement of the four prototypes (nccl.h: ncclResult_t is an int-sized enum, ncclComm_t an opaque
// pointer, ncclDataType_t an int-sized enum with ncclUint8 == 1 in every NCCL 2.x release).
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_p2p_send_fn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_p2p_recv_fn)(void*, size_t, int, int, void*, cudaStream_t);
constexpr int kNcclUint8 = 1;

struct NcclApi {
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    nccl_p2p_send_fn send = nullptr;
    nccl_p2p_recv_fn recv = nullptr;
    bool ok = false;
};

static const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the host application already uses
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = RTLD_DEFAULT;                                   // statically linked into the host, perhaps
        api.group_start = (nccl_group_fn)dlsym(h, "ncclGroupStart");
        api.group_end = (nccl_group_fn)dlsym(h, "ncclGroupEnd");
        api.send = (nccl_p2p_send_fn)dlsym(h, "ncclSend");
        api.recv = (nccl_p2p_recv_fn)dlsym(h, "ncclRecv");
        api.ok = api.group_start && api.group_end && api.send && api.recv;
    });
    return api;
}

}  // namespace mie

using namespace mie;

extern "C" {

int mie_halo_exchange_available(void) { return nccl_api().ok ? 1 : 0; }

// cudaIpcGetMemHandle wants the BASE of the allocation; a pointer into a caching allocator's segment is resolved with
// the driver's cuMemGetAddressRange, taken from the libcuda.so.1 that is already loaded into the process.
int mie_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset_bytes) {
    if (!dev_ptr || !handle64 || !offset_bytes) return MIE_E_NULL;
    typedef int (*get_range_fn)(unsigned long long*, size_t*, unsigned long long);
    static get_range_fn get_range = [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libcuda.so.1", RTLD_NOW);
        return h ? (get_range_fn)dlsym(h, "cuMemGetAddressRange_v2") : (get_range_fn) nullptr;
    }();
    if (!get_range) return MIE_E_UNSUPPORTED;
    unsigned long long base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0 || !base) return MIE_E_UNSUPPORTED;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    memcpy(handle64, &h, 64);
    *offset_bytes = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return MIE_OK;
}

int mie_ipc_open(const void* handle64, void** base_out) {
    if (!handle64 || !base_out) return MIE_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);   // current device = the reader
    if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
    *base_out = p;
    return MIE_OK;
}

int mie_ipc_close(void* base) {
    if (!base) return MIE_OK;
    cudaError_t e = cudaIpcCloseMemHandle(base);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
    return MIE_OK;
}

int mie_enable_peer_access(int peer_device) {
    int cur = -1, count = 0;
    cudaError_t e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return (int)e;
    e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return (int)e;
    if (peer_device < 0 || peer_device >= count) return MIE_E_SHAPE;
    if (peer_device == cur) return MIE_OK;
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
    if (e != cudaSuccess) return (int)e;
    if (!can) return MIE_E_UNSUPPORTED;
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return MIE_OK; }
    return e == cudaSuccess ? MIE_OK : (int)e;
}

int mie_halo_exchange_z(void* nccl_comm, int rank, int world, const void* first_plane, const void* last_plane,
                        void* halo_lo, void* halo_hi, size_t plane_bytes, void* stream) {
    if (world <= 0 || rank < 0 || rank >= world) return MIE_E_SHAPE;
    if (world == 1 || plane_bytes == 0) return MIE_OK;              // nothing to exchange
    if (!nccl_comm) return MIE_E_NULL;
    const bool has_lo = rank > 0, has_hi = rank < world - 1;
    if ((has_lo && (!first_plane || !halo_lo)) || (has_hi && (!last_plane || !halo_hi))) return MIE_E_NULL;
    const NcclApi& api = nccl_api();
    if (!api.ok) return MIE_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = api.group_start();
    if (rc) return MIE_E_NCCL_BASE - rc;
    // sends first, then receives, lower neighbour before upper: every rank issues the same order, and inside one
    // group NCCL matches the pairs regardless of it
    if (has_lo && !rc) rc = api.send(first_plane, plane_bytes, kNcclUint8, rank - 1, nccl_comm, st);
    if (has_hi && !rc) rc = api.send(last_plane, plane_bytes, kNcclUint8, rank + 1, nccl_comm, st);
    if (has_lo && !rc) rc = api.recv(halo_lo, plane_bytes, kNcclUint8, rank - 1, nccl_comm, st);
    if (has_hi && !rc) rc = api.recv(halo_hi, plane_bytes, kNcclUint8, rank + 1, nccl_comm, st);
    const int rc_end = api.group_end();
    if (rc) return MIE_E_NCCL_BASE - rc;
    if (rc_end) return MIE_E_NCCL_BASE - rc_end;
    return MIE_OK;
}

}  // extern "C"
