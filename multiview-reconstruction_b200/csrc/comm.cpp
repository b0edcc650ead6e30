// Halo exchange of psi between the boxes of a 2-d (y x z) process grid, enqueued on the context's compute stream:
//   [pack y rows] -> ncclGroup{send/recv to the y neighbours} -> [unpack] -> ncclGroup{send/recv z planes incl. the fresh y halos}
// so that corners are correct and the host never blocks.  NCCL is loaded lazily with dlopen (no hard link-time dependency; inside a
// torch process the already-loaded libnccl.so.2 is reused).  One process per GPU; the unique id travels through the host's own
// plumbing (torch.distributed in bench.py, a socket / MPI elsewhere).
#include "engine.h"

#include <cstdint>

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

// copy rows [ya, ya+rows) of planes [za, za+planes) between the array and a dense staging buffer (float4 granules when aligned)
template <typename T>
struct PackRows {
    T* vol; T* stage; long long nx; int ny, ya, za, rows; int to_stage;
    MVD_HD void operator()(long long i) const {
        const long long x = i % nx;
        const long long r = i / nx;
        const int y = (int)(r % rows), z = (int)(r / rows);
        const long long vi = ((long long)(za + z) * ny + (ya + y)) * nx + x;
        if (to_stage) stage[i] = vol[vi]; else vol[vi] = stage[i];
    }
};
struct Quad { float a, b, c, d; };
static void pack_rows(const HaloBox& b, float* stage, int ya, int rows, int to_stage, stream_t s) {
    const int planes = b.z1 - b.z0;
    if (rows <= 0 || planes <= 0) return;
    if (b.row_floats % 4 == 0 && ((uintptr_t)b.base % 16) == 0 && ((uintptr_t)stage % 16) == 0) {
        PackRows<Quad> p{(Quad*)b.base, (Quad*)stage, b.row_floats / 4, b.nrows, ya, b.z0, rows, to_stage};
        pfor((long long)rows * p.nx * planes, p, s);
    } else {
        PackRows<float> p{b.base, stage, b.row_floats, b.nrows, ya, b.z0, rows, to_stage};
        pfor((long long)rows * p.nx * planes, p, s);
    }
}

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

HaloComm::HaloComm(std::shared_ptr<NcclComm> comm, int py, int pz, stream_t s)
    : comm_(std::move(comm)), py_(py), pz_(pz), stream_(s) {
    if (!comm_ || py < 1 || pz < 1 || py * pz != comm_->world()) throw Error("bad process grid");
    ry_ = comm_->rank() / pz; rz_ = comm_->rank() % pz;
}

HaloComm::~HaloComm() {
    for (float* p : stage_) dev::free_(p);
}

void HaloComm::reserve(size_t floats) {
    if (floats <= stage_floats_) return;
    dev::sync(stream_);
    for (float*& p : stage_) { dev::free_(p); p = (float*)dev::alloc(sizeof(float) * floats); }
    stage_floats_ = floats;
}

// Neighbour below (ry-1 / rz-1): it needs my first h*_hi own rows / planes (its upper halo) and sends its last h*_lo ones (my lower
// halo); the neighbour above mirrors that.  The widths are the same on every rank (they come from the kernel extents).
void HaloComm::exchange(const HaloBox& b) {
#ifndef MVD_HOST_EMU
    NcclApi& n = nccl();
    ncclComm_t c = (ncclComm_t)comm_->raw();
    if (py_ > 1 && (b.hy_lo > 0 || b.hy_hi > 0)) {
        const bool lower = ry_ > 0, upper = ry_ < py_ - 1;
        if ((lower && (b.y0 - b.hy_lo < 0 || b.y0 + b.hy_hi > b.y1)) || (upper && (b.y1 + b.hy_hi > b.nrows || b.y1 - b.hy_lo < b.y0)))
            throw Error("halo exchange: the array does not contain the halo rows");
        const size_t per_row = (size_t)b.row_floats * (size_t)(b.z1 - b.z0);
        reserve((size_t)std::max(b.hy_lo, b.hy_hi) * per_row);
        const size_t cnt_lo = (size_t)b.hy_lo * per_row, cnt_hi = (size_t)b.hy_hi * per_row;
        if (lower) pack_rows(b, stage_[0], b.y0, b.hy_hi, 1, stream_);
        if (upper) pack_rows(b, stage_[1], b.y1 - b.hy_lo, b.hy_lo, 1, stream_);
        nccl_check(n.GroupStart(), "group");
        if (lower) {
            const int peer = (ry_ - 1) * pz_ + rz_;
            if (cnt_hi) nccl_check(n.Send(stage_[0], cnt_hi, kNcclFloat, peer, c, stream_), "send");
            if (cnt_lo) nccl_check(n.Recv(stage_[2], cnt_lo, kNcclFloat, peer, c, stream_), "recv");
        }
        if (upper) {
            const int peer = (ry_ + 1) * pz_ + rz_;
            if (cnt_lo) nccl_check(n.Send(stage_[1], cnt_lo, kNcclFloat, peer, c, stream_), "send");
            if (cnt_hi) nccl_check(n.Recv(stage_[3], cnt_hi, kNcclFloat, peer, c, stream_), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
        if (lower) pack_rows(b, stage_[2], b.y0 - b.hy_lo, b.hy_lo, 0, stream_);
        if (upper) pack_rows(b, stage_[3], b.y1, b.hy_hi, 0, stream_);
    }
    if (pz_ > 1 && (b.hz_lo > 0 || b.hz_hi > 0)) {
        const bool lower = rz_ > 0, upper = rz_ < pz_ - 1;
        if ((lower && (b.z0 - b.hz_lo < 0 || b.z0 + b.hz_hi > b.z1)) || (upper && (b.z1 + b.hz_hi > b.nplanes || b.z1 - b.hz_lo < b.z0)))
            throw Error("halo exchange: the array does not contain the halo planes");
        const size_t plane = (size_t)b.row_floats * (size_t)b.nrows;
        const size_t cnt_lo = plane * b.hz_lo, cnt_hi = plane * b.hz_hi;
        nccl_check(n.GroupStart(), "group");
        if (lower) {
            const int peer = ry_ * pz_ + rz_ - 1;
            if (cnt_hi) nccl_check(n.Send(b.base + plane * b.z0, cnt_hi, kNcclFloat, peer, c, stream_), "send");
            if (cnt_lo) nccl_check(n.Recv(b.base + plane * (b.z0 - b.hz_lo), cnt_lo, kNcclFloat, peer, c, stream_), "recv");
        }
        if (upper) {
            const int peer = ry_ * pz_ + rz_ + 1;
            if (cnt_lo) nccl_check(n.Send(b.base + plane * (b.z1 - b.hz_lo), cnt_lo, kNcclFloat, peer, c, stream_), "send");
            if (cnt_hi) nccl_check(n.Recv(b.base + plane * b.z1, cnt_hi, kNcclFloat, peer, c, stream_), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
    }
#else
    (void)b;
#endif
}

}  // namespace mvd
