// Halo exchange of psi between the boxes of a 2-d (y x z) process grid, enqueued on the context's compute stream:
//   [pack y rows] -> ncclGroup{send/recv to the y neighbours} -> [unpack] -> ncclGroup{send/recv z planes incl. the fresh y halos}
// so that corners are correct and the host never blocks.  NCCL is loaded lazily with dlopen (no hard link-time dependency; inside a
// torch process the already-loaded libnccl.so.2 is reused).  One process per GPU; the unique id travels through the host's own
// plumbing (torch.distributed in bench.py, a socket / MPI elsewhere).
#include "engine.h"

#ifndef MVD_HOST_EMU
#include <dlfcn.h>
#endif

namespace mvd {

#ifndef MVD_HOST_EMU
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) throw Error("NCCL is not available (dlopen libnccl.so.2 failed): multi-GPU halo exchange needs it");
#define MVD_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); if (!api.field) throw Error("NCCL symbol missing: " name);
    MVD_SYM(GetUniqueId, "ncclGetUniqueId") MVD_SYM(CommInitRank, "ncclCommInitRank") MVD_SYM(CommDestroy, "ncclCommDestroy")
    MVD_SYM(GroupStart, "ncclGroupStart") MVD_SYM(GroupEnd, "ncclGroupEnd") MVD_SYM(Send, "ncclSend") MVD_SYM(Recv, "ncclRecv")
    MVD_SYM(GetErrorString, "ncclGetErrorString")
#undef MVD_SYM
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != 0) throw Error(std::string("NCCL ") + what + ": " + nccl().GetErrorString(r));
}
constexpr int kNcclFloat = 7;
}  // namespace

#endif

// copy rows [ya, ya+rows) of planes [za, za+planes) between the local volume and a dense staging buffer
struct PackRows {
    float* vol; float* stage; int nx, ny, ya, za, rows; int to_stage;
    MVD_HD void operator()(long long i) const {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % rows), z = (int)(r / rows);
        const long long vi = ((long long)(za + z) * ny + (ya + y)) * nx + x;
        if (to_stage) stage[i] = vol[vi]; else vol[vi] = stage[i];
    }
};

void NcclComm::unique_id(char out[128]) {
#ifndef MVD_HOST_EMU
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, id.internal, 128);
#else
    (void)out;
    throw Error("the host emulator has no NCCL");
#endif
}

NcclComm::NcclComm(const char id[128], int world, int rank, int device) : world_(world), rank_(rank), device_(device) {
    if (world < 1 || rank < 0 || rank >= world) throw Error("bad communicator rank / size");
#ifndef MVD_HOST_EMU
    dev::set_device(device);
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    ncclComm_t c = nullptr;
    nccl_check(nccl().CommInitRank(&c, world, uid, rank), "ncclCommInitRank");
    comm_ = c;
#else
    (void)id;
    throw Error("the host emulator has no NCCL");
#endif
}
NcclComm::~NcclComm() {
#ifndef MVD_HOST_EMU
    if (comm_) nccl().CommDestroy((ncclComm_t)comm_);
#endif
}

HaloComm::HaloComm(std::shared_ptr<NcclComm> comm, int py, int pz, const Geometry& g, int halo_y, int halo_z, stream_t s)
    : comm_(std::move(comm)), py_(py), pz_(pz), g_(g), hy_(halo_y), hz_(halo_z), stream_(s) {
    if (!comm_ || py * pz != comm_->world()) throw Error("bad process grid");
    ry_ = comm_->rank() / pz; rz_ = comm_->rank() % pz;
    if (py > 1) {
        const size_t n = (size_t)hy_ * g.vol[0] * (size_t)(g.own_hi[2] - g.own_lo[2]);
        for (int i = 0; i < 4; ++i) stage_[i] = (float*)dev::alloc(sizeof(float) * n);
    }
}

HaloComm::~HaloComm() {
    for (float* p : stage_) dev::free_(p);
}

void HaloComm::exchange(float* psi) {
#ifndef MVD_HOST_EMU
    NcclApi& n = nccl();
    const int nx = g_.vol[0], ny = g_.vol[1];
    const int ylo = g_.own_lo[1] - g_.goff[1], yhi = g_.own_hi[1] - g_.goff[1];       // local indices
    const int zlo = g_.own_lo[2] - g_.goff[2], zhi = g_.own_hi[2] - g_.goff[2];
    if (py_ > 1) {
        const int planes = zhi - zlo;
        const long long cnt = (long long)hy_ * nx * planes;
        const bool lower = ry_ > 0, upper = ry_ < py_ - 1;
        if (lower) pfor(cnt, PackRows{psi, stage_[0], nx, ny, ylo, zlo, hy_, 1}, stream_);
        if (upper) pfor(cnt, PackRows{psi, stage_[1], nx, ny, yhi - hy_, zlo, hy_, 1}, stream_);
        nccl_check(n.GroupStart(), "group");
        if (lower) {
            nccl_check(n.Send(stage_[0], (size_t)cnt, kNcclFloat, (ry_ - 1) * pz_ + rz_, (ncclComm_t)comm_->raw(), stream_), "send");
            nccl_check(n.Recv(stage_[2], (size_t)cnt, kNcclFloat, (ry_ - 1) * pz_ + rz_, (ncclComm_t)comm_->raw(), stream_), "recv");
        }
        if (upper) {
            nccl_check(n.Send(stage_[1], (size_t)cnt, kNcclFloat, (ry_ + 1) * pz_ + rz_, (ncclComm_t)comm_->raw(), stream_), "send");
            nccl_check(n.Recv(stage_[3], (size_t)cnt, kNcclFloat, (ry_ + 1) * pz_ + rz_, (ncclComm_t)comm_->raw(), stream_), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
        if (lower) pfor(cnt, PackRows{psi, stage_[2], nx, ny, ylo - hy_, zlo, hy_, 0}, stream_);
        if (upper) pfor(cnt, PackRows{psi, stage_[3], nx, ny, yhi, zlo, hy_, 0}, stream_);
    }
    if (pz_ > 1) {
        const size_t plane = (size_t)nx * ny, cnt = plane * hz_;
        nccl_check(n.GroupStart(), "group");
        if (rz_ > 0) {
            nccl_check(n.Send(psi + plane * zlo, cnt, kNcclFloat, ry_ * pz_ + rz_ - 1, (ncclComm_t)comm_->raw(), stream_), "send");
            nccl_check(n.Recv(psi + plane * (zlo - hz_), cnt, kNcclFloat, ry_ * pz_ + rz_ - 1, (ncclComm_t)comm_->raw(), stream_), "recv");
        }
        if (rz_ < pz_ - 1) {
            nccl_check(n.Send(psi + plane * (zhi - hz_), cnt, kNcclFloat, ry_ * pz_ + rz_ + 1, (ncclComm_t)comm_->raw(), stream_), "send");
            nccl_check(n.Recv(psi + plane * zhi, cnt, kNcclFloat, ry_ * pz_ + rz_ + 1, (ncclComm_t)comm_->raw(), stream_), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
    }
#else
    (void)psi;
#endif
}

}  // namespace mvd
