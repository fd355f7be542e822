// fdtd_slab.cu -- x-slab neighbours.  A slab exports one CUDA IPC handle (its u allocation, which also
// holds the flag words) plus the numbers a neighbour needs to address its ghost planes; attaching maps
// the neighbour's memory so the stencil kernel can store boundary planes into it directly (NVLink).
// One process per GPU exchanges the blobs over torch.distributed; a single process driving several
// devices (or several slabs on one device, for tests) attaches plans directly.
#include "fdtd_plan.h"

#include <string.h>

namespace {
struct SlabBlob {                  // FDTD_B200_IPC_BYTES = 160
    cudaIpcMemHandle_t handle;     // 64 bytes
    long long lvl;
    unsigned long long flags_offset;
    unsigned long long tile_flags_offset;
    int nxp, nyp, nzp, X0, X1, dev;
    int magic;
};
static_assert(sizeof(SlabBlob) <= FDTD_B200_IPC_BYTES, "blob too large");
constexpr int kMagic = 0x46445444;  // "FDTD"

// side 0: the neighbour holds the planes below ours and we fill ITS upper ghost planes (its X1, X1+1, ..) and
// raise ITS ready-from-upper flag [1]; side 1: we fill its lower ghost planes (.., its X0-2, X0-1), flag [0].
int link_to(fdtd_b200_plan *p, int side, float *peer_u, const SlabBlob &b)
{
    if (b.nyp != p->g.nyp || b.nzp != p->g.nzp) return (int)cudaErrorInvalidValue;
    int *peer_flags = reinterpret_cast<int *>(reinterpret_cast<char *>(peer_u) + b.flags_offset);
    p->link.peer_u[side] = peer_u;
    p->link.peer_lvl[side] = b.lvl;
    p->link.peer_edge[side] = side == 0 ? b.X1 : b.X0;
    p->link.peer_nxp[side] = b.nxp;
    p->tma.valid = p->tb2.valid = false;  // the plans hold tensor maps of the neighbours' arrays
    p->link.peer_flag[side] = peer_flags + (side == 0 ? 1 : 0);
    // the neighbour sees this slab on ITS side 1 - side: that is the array this slab raises
    p->link.peer_tile[side] = reinterpret_cast<int *>(reinterpret_cast<char *>(peer_u) + b.tile_flags_offset) + (side == 0 ? fdtd::kMaxFlagTiles : 0);
    return 0;
}

void fill_blob(fdtd_b200_plan *p, SlabBlob &b)
{
    memset(&b, 0, sizeof(b));
    b.lvl = p->g.lvl;
    b.flags_offset = p->flags_offset;
    b.tile_flags_offset = p->tile_flags_offset;
    b.nxp = p->g.nxp;
    b.nyp = p->g.nyp;
    b.nzp = p->g.nzp;
    b.X0 = p->g.X0;
    b.X1 = p->g.X1;
    b.dev = p->dev;
    b.magic = kMagic;
}
}  // namespace

extern "C" int fdtd_b200_plan_ipc_export(fdtd_b200_plan *p, void *blob)
{
    if (!p || !blob) return (int)cudaErrorInvalidValue;
    FDTD_CHECK(cudaSetDevice(p->dev));
    SlabBlob b;
    fill_blob(p, b);
    FDTD_CHECK(cudaIpcGetMemHandle(&b.handle, p->d_u));
    memset(blob, 0, FDTD_B200_IPC_BYTES);
    memcpy(blob, &b, sizeof(b));
    return 0;
}

extern "C" int fdtd_b200_plan_ipc_attach(fdtd_b200_plan *p, int side, const void *blob)
{
    if (!p || !blob || side < 0 || side > 1) return (int)cudaErrorInvalidValue;
    SlabBlob b;
    memcpy(&b, blob, sizeof(b));
    if (b.magic != kMagic) return (int)cudaErrorInvalidValue;
    FDTD_CHECK(cudaSetDevice(p->dev));
    void *base = nullptr;
    FDTD_CHECK(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
    p->ipc_base[side] = base;
    return link_to(p, side, static_cast<float *>(base), b);
}

extern "C" int fdtd_b200_plan_attach_local(fdtd_b200_plan *p, int side, fdtd_b200_plan *nb)
{
    if (!p || !nb || side < 0 || side > 1) return (int)cudaErrorInvalidValue;
    FDTD_CHECK(cudaSetDevice(p->dev));
    if (nb->dev != p->dev) {
        int can = 0;
        FDTD_CHECK(cudaDeviceCanAccessPeer(&can, p->dev, nb->dev));
        if (!can) return (int)cudaErrorPeerAccessUnsupported;
        cudaError_t e = cudaDeviceEnablePeerAccess(nb->dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)e;
        (void)cudaGetLastError();
    } else {
        // slabs on one device (tests): all of them use one process-wide stream of that device so their
        // steps serialise -- a kernel may never spin on a flag that a not-yet-scheduled kernel of the same
        // GPU has to raise.  The shared stream lives until the process exits.
        static cudaStream_t shared[64] = {nullptr};
        const int d = p->dev & 63;
        if (!shared[d]) FDTD_CHECK(cudaStreamCreateWithFlags(&shared[d], cudaStreamNonBlocking));
        for (fdtd_b200_plan *q : {p, nb}) {
            if (q->stream == shared[d]) continue;
            FDTD_CHECK(cudaStreamSynchronize(q->stream));
            if (q->owns_stream) cudaStreamDestroy(q->stream);
            q->stream = shared[d];
            q->owns_stream = false;
        }
    }
    SlabBlob b;
    fill_blob(nb, b);
    return link_to(p, side, nb->d_u, b);
}
