// Halo exchange of psi between the boxes of a 2-d (y x z) process grid, enqueued on the context's compute stream:
//   [pack y rows] -> ncclGroup{send/recv to the y neighbours} -> [unpack] -> ncclGroup{send/recv z planes incl. the fresh y halos}
// so that corners are correct and the host never blocks.  NCCL is loaded lazily with dlopen (no hard link-time dependency; inside a
// torch process the already-loaded libnccl.so.2 is reused).  One process per GPU; the unique id travels through the host's own
// plumbing (torch.distributed in bench.py, a socket / MPI elsewhere).
#include "engine.h"

#include <cstdint>

#ifndef MVD_HOST_EMU
#include <dlfcn.h>
#include <unistd.h>
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
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (api.handle) return api;
    // an NCCL the process already holds (e.g. the one bundled with torch) wins; MVD_NCCL_LIB names a specific file
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.handle) break; }
    if (!api.handle) if (const char* e = std::getenv("MVD_NCCL_LIB")) api.handle = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    for (const char* n : names) { if (api.handle) break; api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL); }
    if (!api.handle) throw Error("NCCL is not available (dlopen libnccl.so.2 failed): multi-GPU halo exchange needs it");
#define MVD_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name)); if (!api.field) throw Error("NCCL symbol missing: " name);
    MVD_SYM(GetUniqueId, "ncclGetUniqueId") MVD_SYM(CommInitRank, "ncclCommInitRank") MVD_SYM(CommDestroy, "ncclCommDestroy")
    MVD_SYM(GroupStart, "ncclGroupStart") MVD_SYM(GroupEnd, "ncclGroupEnd") MVD_SYM(Send, "ncclSend") MVD_SYM(Recv, "ncclRecv")
    MVD_SYM(GetErrorString, "ncclGetErrorString") MVD_SYM(AllGather, "ncclAllGather") MVD_SYM(AllReduce, "ncclAllReduce")
#undef MVD_SYM
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != 0) throw Error(std::string("NCCL ") + what + ": " + nccl().GetErrorString(r));
}
constexpr int kNcclFloat = 7, kNcclChar = 0, kNcclDouble = 8, kNcclSum = 0, kNcclMax = 2;
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
struct alignas(16) Quad { float a, b, c, d; };
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

// whole planes [za, za + planes) <-> dense staging, scalar granules (the fallback for boxes whose planes are not 16-byte multiples)
static void pack_planes(const HaloBox& b, float* stage, int za, int planes, int to_stage, stream_t s) {
    if (planes <= 0) return;
    PackRows<float> p{b.base, stage, b.row_floats, b.nrows, 0, za, b.nrows, to_stage};
    pfor((long long)b.nrows * b.row_floats * planes, p, s);
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

HaloComm::HaloComm(std::shared_ptr<NcclComm> comm, int py, int pz, stream_t s, size_t need_y, size_t need_z)
    : comm_(std::move(comm)), py_(py), pz_(pz), stream_(s) {
    if (!comm_ || py < 1 || pz < 1 || py * pz != comm_->world()) throw Error("bad process grid");
    ry_ = comm_->rank() / pz; rz_ = comm_->rank() % pz;
    setup_peer(need_y, need_z);
}

HaloComm::~HaloComm() {
    for (float* p : stage_) dev::free_(p);
    dev::free_(red_dev_);
    close_peer();
}

void HaloComm::reserve(size_t floats, stream_t also) {
    if (floats <= stage_floats_) return;
    dev::sync(stream_);
    if (also && also != stream_) dev::sync(also);          // an earlier exchange on the exchange stream may still read the old buffers
    for (float*& p : stage_) { dev::free_(p); p = (float*)dev::alloc(sizeof(float) * floats); }
    stage_floats_ = floats;
}

static void check_box(const HaloBox& b, bool ylower, bool yupper, bool zlower, bool zupper) {
    if ((ylower && (b.y0 - b.hy_lo < 0 || b.y0 + b.hy_hi > b.y1)) || (yupper && (b.y1 + b.hy_hi > b.nrows || b.y1 - b.hy_lo < b.y0)))
        throw Error("halo exchange: the array does not contain the halo rows");
    if ((zlower && (b.z0 - b.hz_lo < 0 || b.z0 + b.hz_hi > b.z1)) || (zupper && (b.z1 + b.hz_hi > b.nplanes || b.z1 - b.hz_lo < b.z0)))
        throw Error("halo exchange: the array does not contain the halo planes");
}

void HaloComm::exchange(const HaloBox& b, bool force_nccl, stream_t s) {
    if (!s) s = stream_;
    const bool ydo = py_ > 1 && (b.hy_lo > 0 || b.hy_hi > 0), zdo = pz_ > 1 && (b.hz_lo > 0 || b.hz_hi > 0);
    check_box(b, ydo && ry_ > 0, ydo && ry_ < py_ - 1, zdo && rz_ > 0, zdo && rz_ < pz_ - 1);
    if (peer_ && !force_nccl) exchange_peer(b, s); else exchange_nccl(b, s);
}

void HaloComm::all_reduce(double* host_values, int count, int op) {
#ifndef MVD_HOST_EMU
    if (count <= 0) return;
    if ((size_t)count > red_cap_) {
        dev::sync(stream_);
        dev::free_(red_dev_);
        red_cap_ = std::max((size_t)count, (size_t)256);
        red_dev_ = (double*)dev::alloc(sizeof(double) * red_cap_);
    }
    dev::h2d(red_dev_, host_values, sizeof(double) * (size_t)count, stream_);
    nccl_check(nccl().AllReduce(red_dev_, red_dev_, (size_t)count, kNcclDouble, op == 1 ? kNcclMax : kNcclSum, (ncclComm_t)comm_->raw(), stream_),
               "allreduce");
    dev::d2h(host_values, red_dev_, sizeof(double) * (size_t)count, stream_);
    dev::sync(stream_);
#else
    (void)host_values; (void)count; (void)op;
    throw Error("the host emulator has no NCCL");
#endif
}

// Neighbour below (ry-1 / rz-1): it needs my first h*_hi own rows / planes (its upper halo) and sends its last h*_lo ones (my lower
// halo); the neighbour above mirrors that.  The widths are the same on every rank (they come from the kernel extents).
void HaloComm::exchange_nccl(const HaloBox& b, stream_t xs) {
#ifndef MVD_HOST_EMU
    NcclApi& n = nccl();
    ncclComm_t c = (ncclComm_t)comm_->raw();
    if (py_ > 1 && (b.hy_lo > 0 || b.hy_hi > 0)) {
        const bool lower = ry_ > 0, upper = ry_ < py_ - 1;
        const size_t per_row = (size_t)b.row_floats * (size_t)(b.z1 - b.z0);
        reserve((size_t)std::max(b.hy_lo, b.hy_hi) * per_row, xs);
        const size_t cnt_lo = (size_t)b.hy_lo * per_row, cnt_hi = (size_t)b.hy_hi * per_row;
        if (lower) pack_rows(b, stage_[0], b.y0, b.hy_hi, 1, xs);
        if (upper) pack_rows(b, stage_[1], b.y1 - b.hy_lo, b.hy_lo, 1, xs);
        nccl_check(n.GroupStart(), "group");
        if (lower) {
            const int peer = (ry_ - 1) * pz_ + rz_;
            if (cnt_hi) nccl_check(n.Send(stage_[0], cnt_hi, kNcclFloat, peer, c, xs), "send");
            if (cnt_lo) nccl_check(n.Recv(stage_[2], cnt_lo, kNcclFloat, peer, c, xs), "recv");
        }
        if (upper) {
            const int peer = (ry_ + 1) * pz_ + rz_;
            if (cnt_lo) nccl_check(n.Send(stage_[1], cnt_lo, kNcclFloat, peer, c, xs), "send");
            if (cnt_hi) nccl_check(n.Recv(stage_[3], cnt_hi, kNcclFloat, peer, c, xs), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
        if (lower) pack_rows(b, stage_[2], b.y0 - b.hy_lo, b.hy_lo, 0, xs);
        if (upper) pack_rows(b, stage_[3], b.y1, b.hy_hi, 0, xs);
    }
    if (pz_ > 1 && (b.hz_lo > 0 || b.hz_hi > 0)) {
        const bool lower = rz_ > 0, upper = rz_ < pz_ - 1;
        const size_t plane = (size_t)b.row_floats * (size_t)b.nrows;
        const size_t cnt_lo = plane * b.hz_lo, cnt_hi = plane * b.hz_hi;
        nccl_check(n.GroupStart(), "group");
        if (lower) {
            const int peer = ry_ * pz_ + rz_ - 1;
            if (cnt_hi) nccl_check(n.Send(b.base + plane * b.z0, cnt_hi, kNcclFloat, peer, c, xs), "send");
            if (cnt_lo) nccl_check(n.Recv(b.base + plane * (b.z0 - b.hz_lo), cnt_lo, kNcclFloat, peer, c, xs), "recv");
        }
        if (upper) {
            const int peer = ry_ * pz_ + rz_ + 1;
            if (cnt_lo) nccl_check(n.Send(b.base + plane * (b.z1 - b.hz_lo), cnt_lo, kNcclFloat, peer, c, xs), "send");
            if (cnt_hi) nccl_check(n.Recv(b.base + plane * b.z1, cnt_hi, kNcclFloat, peer, c, xs), "recv");
        }
        nccl_check(n.GroupEnd(), "group");
    }
#else
    (void)b; (void)xs;
#endif
}

// ------------------------------------------------------------------------------------------------------------------------------
// peer transport: the halo rows / planes are stored straight into a landing buffer in the neighbour's HBM over NVLink (no NCCL
// FIFO, no proxy, no staging on the sending side), followed by a release of an arrival counter in the neighbour's memory; the
// receiver spins on its own counter and unpacks the landing buffer locally.  Landing buffers alternate with the exchange parity:
// exchange s+2 can only be pushed after the neighbour's push of s+1 arrived, which it enqueued after unpacking s.
// ------------------------------------------------------------------------------------------------------------------------------
float* HaloComm::landing(float* region, int src, int parity) const {
    return src < 2 ? region + (size_t)(src * 2 + parity) * cap_y_ : region + 4 * cap_y_ + (size_t)((src - 2) * 2 + parity) * cap_z_;
}
unsigned* HaloComm::flags(float* region) const { return reinterpret_cast<unsigned*>(region + 4 * cap_y_ + 4 * cap_z_); }

struct SignalArrival {     // runs after the push kernels of this stream completed, i.e. after their stores were performed
    volatile unsigned* flag; unsigned seq;
    MVD_HD void operator()(long long) const {
#if defined(__CUDA_ARCH__)
        __threadfence_system();
#endif
        *flag = seq;
    }
};
struct AwaitArrival {
    volatile unsigned* f0; volatile unsigned* f1; unsigned seq;
    MVD_HD void operator()(long long) const {
#if defined(__CUDA_ARCH__)
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int k = 0; k < 2; ++k) {
            volatile unsigned* f = k == 0 ? f0 : f1;
            if (!f) continue;
            while ((int)(*f - seq) < 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > 120000000000ULL) __trap();           // a neighbour died: fail loudly instead of hanging the device
                __nanosleep(64);
            }
        }
        __threadfence_system();
#endif
    }
};
struct CopyQuads {
    const Quad* src; Quad* dst;
    MVD_HD void operator()(long long i) const { dst[i] = src[i]; }
};

#ifndef MVD_HOST_EMU
namespace {
// One halo slab as rows [ya, ya + rows) x planes [za, za + planes) of a [.][ny][nx4] volume of 16-byte granules and its dense image
// in a landing buffer (z slabs: all rows of whole planes).
struct HaloSeg {
    Quad* vol; Quad* stage; long long nx4, count; int ny, ya, za, rows;
    __device__ __forceinline__ long long vol_index(long long i) const {
        const long long x = i % nx4, r = i / nx4;
        const int y = (int)(r % rows), z = (int)(r / rows);
        return ((long long)(za + z) * ny + (ya + y)) * nx4 + x;
    }
};
inline HaloSeg make_seg(const HaloBox& b, float* stage, int ya, int rows, int za, int planes) {
    HaloSeg s;
    s.vol = (Quad*)b.base; s.stage = (Quad*)stage; s.nx4 = b.row_floats / 4; s.ny = b.nrows; s.ya = ya; s.za = za; s.rows = rows;
    s.count = (rows > 0 && planes > 0) ? s.nx4 * rows * planes : 0;
    return s;
}
inline unsigned long long exchange_timeout_ns() {      // MVD_EXCHANGE_TIMEOUT_S: watchdog of the arrival wait (default 120 s, 0 = wait forever)
    static const unsigned long long v = [] {
        const char* e = std::getenv("MVD_EXCHANGE_TIMEOUT_S");
        return (unsigned long long)(e ? std::atof(e) : 120.0) * 1000000000ULL;
    }();
    return v;
}
// push: both slabs of one exchange phase go straight into the neighbours' landing buffers (remote stores over NVLink); the CTA that
// finishes last releases the neighbours' arrival flags -- pack, pack, signal, signal in ONE launch.
__global__ void __launch_bounds__(256) halo_push_kernel(HaloSeg s0, HaloSeg s1, unsigned* flag0, unsigned* flag1, unsigned seq, unsigned* done) {
    const long long n = s0.count + s1.count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i < s0.count) s0.stage[i] = s0.vol[s0.vol_index(i)];
        else { const long long k = i - s0.count; s1.stage[k] = s1.vol[s1.vol_index(k)]; }
    }
    __threadfence_system();                 // this thread's remote stores are performed before its CTA reports in
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {        // every CTA of the launch has stored (and fenced) its part
            *done = 0;
            __threadfence_system();
            if (flag0) *(volatile unsigned*)flag0 = seq;
            if (flag1) *(volatile unsigned*)flag1 = seq;
        }
    }
}
// pull: wait for the neighbours' arrival flags, then unpack both landing buffers -- await, unpack, unpack in ONE launch
__global__ void __launch_bounds__(256) halo_pull_kernel(HaloSeg s0, HaloSeg s1, const unsigned* f0, const unsigned* f1, unsigned seq,
                                                        unsigned long long timeout_ns) {
    if (threadIdx.x == 0) {
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int k = 0; k < 2; ++k) {
            const volatile unsigned* f = k == 0 ? f0 : f1;
            if (!f) continue;
            while ((int)(*f - seq) < 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (timeout_ns && t - t0 > timeout_ns) __trap();      // a neighbour died: fail loudly instead of hanging the device
                __nanosleep(32);
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const long long n = s0.count + s1.count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (i < s0.count) s0.vol[s0.vol_index(i)] = s0.stage[i];
        else { const long long k = i - s0.count; s1.vol[s1.vol_index(k)] = s1.stage[k]; }
    }
}
inline int halo_grid(long long n, int per_sm = 8) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * per_sm) blocks = 148LL * per_sm;
    return (int)(blocks < 1 ? 1 : blocks);
}
struct PeerInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long pid, ptr, need_y, need_z;
    int device, ok;
};
// byte-wise all-gather of one small record per rank through the communicator
template <class T>
std::vector<T> gather_records(const T& mine, NcclComm& comm, stream_t s) {
    const int W = comm.world();
    char* d = (char*)dev::alloc(sizeof(T) * (size_t)(W + 1));
    dev::h2d(d, &mine, sizeof(T), s);
    nccl_check(nccl().AllGather(d, d + sizeof(T), sizeof(T), kNcclChar, (ncclComm_t)comm.raw(), s), "allgather");
    std::vector<T> all((size_t)W);
    dev::d2h(all.data(), d + sizeof(T), sizeof(T) * (size_t)W, s);
    dev::sync(s);
    dev::free_(d);
    return all;
}
}  // namespace
#endif

void HaloComm::setup_peer(size_t need_y, size_t need_z) {
#ifndef MVD_HOST_EMU
    if (const char* e = std::getenv("MVD_EXCHANGE")) if (std::string(e) == "nccl") return;      // A/B switch
    if (comm_->world() == 1) return;
    // round 1: sizes.  Every landing buffer gets the largest push any rank will ever make (multiple of 64 floats = 256 B).
    PeerInfo me;
    std::memset(&me, 0, sizeof(me));
    me.pid = (unsigned long long)getpid();
    me.device = comm_->device();
    me.need_y = need_y; me.need_z = need_z;
    std::vector<PeerInfo> all = gather_records(me, *comm_, stream_);
    for (const PeerInfo& p : all) { cap_y_ = std::max(cap_y_, (size_t)p.need_y); cap_z_ = std::max(cap_z_, (size_t)p.need_z); }
    cap_y_ = (cap_y_ + 63) / 64 * 64; cap_z_ = (cap_z_ + 63) / 64 * 64;
    const size_t bytes = sizeof(float) * (4 * cap_y_ + 4 * cap_z_) + 256;
    int ok = 1;
    if (cudaMalloc((void**)&region_, bytes) != cudaSuccess) { cudaGetLastError(); region_ = nullptr; ok = 0; }
    if (ok) {
        MVD_CUDA_CHECK(cudaMemsetAsync(flags(region_), 0, 256, stream_));
        dev::sync(stream_);
        if (cudaIpcGetMemHandle(&me.handle, region_) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    me.ptr = (unsigned long long)(uintptr_t)region_;
    me.ok = ok;
    // round 2: handles; map the (up to four) neighbours
    all = gather_records(me, *comm_, stream_);
    const int nb_rank[4] = {ry_ > 0 ? (ry_ - 1) * pz_ + rz_ : -1, ry_ < py_ - 1 ? (ry_ + 1) * pz_ + rz_ : -1,
                            rz_ > 0 ? ry_ * pz_ + rz_ - 1 : -1, rz_ < pz_ - 1 ? ry_ * pz_ + rz_ + 1 : -1};
    for (int i = 0; i < 4 && ok; ++i) {
        if (nb_rank[i] < 0) continue;
        const PeerInfo& p = all[(size_t)nb_rank[i]];
        if (!p.ok) { ok = 0; break; }
        if (p.pid == me.pid) {                               // neighbour context in this process: plain peer access
            if (p.device != me.device) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, me.device, p.device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(p.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
                cudaGetLastError();
            }
            nb_region_[i] = reinterpret_cast<float*>((uintptr_t)p.ptr);
        } else {
            void* q = nullptr;
            if (cudaIpcOpenMemHandle(&q, p.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            nb_region_[i] = (float*)q;
            nb_ipc_[i] = true;
        }
    }
    // round 3: every rank must have mapped its neighbours, otherwise all stay on NCCL
    me.ok = ok;
    all = gather_records(me, *comm_, stream_);
    for (const PeerInfo& p : all) ok = ok && p.ok;
    if (!ok) { close_peer(); return; }
    peer_ = true;
#else
    (void)need_y; (void)need_z;
#endif
}

void HaloComm::close_peer() {
#ifndef MVD_HOST_EMU
    for (int i = 0; i < 4; ++i) {
        if (nb_region_[i] && nb_ipc_[i]) cudaIpcCloseMemHandle(nb_region_[i]);
        nb_region_[i] = nullptr; nb_ipc_[i] = false;
    }
    if (region_) cudaFree(region_);
    region_ = nullptr;
    peer_ = false;
#endif
}

void HaloComm::exchange_peer(const HaloBox& b, stream_t xs) {
#ifndef MVD_HOST_EMU
    ++seq_;
    const int par = (int)(seq_ & 1u);
    unsigned* mine = flags(region_);
    unsigned* done = mine + 16;                           // launch-local CTA counter of the push kernels (returns to 0 after every launch)
    const bool quads = b.row_floats % 4 == 0 && ((uintptr_t)b.base % 16) == 0;
    const unsigned long long tmo = exchange_timeout_ns();
    if (py_ > 1 && (b.hy_lo > 0 || b.hy_hi > 0)) {
        const bool lower = ry_ > 0, upper = ry_ < py_ - 1;
        const int planes = b.z1 - b.z0;
        const size_t per_row = (size_t)b.row_floats * (size_t)planes;
        if ((size_t)std::max(b.hy_lo, b.hy_hi) * per_row > cap_y_) throw Error("halo exchange: landing buffer too small (y)");
        // I am the lower neighbour's UPPER neighbour: my first hy_hi own rows land in its "from upper y" buffer (src 1), and vice versa
        if (quads && (lower || upper)) {
            HaloSeg p0 = make_seg(b, lower ? landing(nb_region_[0], 1, par) : nullptr, b.y0, lower ? b.hy_hi : 0, b.z0, planes);
            HaloSeg p1 = make_seg(b, upper ? landing(nb_region_[1], 0, par) : nullptr, b.y1 - b.hy_lo, upper ? b.hy_lo : 0, b.z0, planes);
            halo_push_kernel<<<halo_grid(p0.count + p1.count, 4), 256, 0, xs>>>(p0, p1, lower ? flags(nb_region_[0]) + 1 : nullptr,
                                                                                   upper ? flags(nb_region_[1]) + 0 : nullptr, seq_, done);
            HaloSeg u0 = make_seg(b, lower ? landing(region_, 0, par) : nullptr, b.y0 - b.hy_lo, lower ? b.hy_lo : 0, b.z0, planes);
            HaloSeg u1 = make_seg(b, upper ? landing(region_, 1, par) : nullptr, b.y1, upper ? b.hy_hi : 0, b.z0, planes);
            halo_pull_kernel<<<halo_grid(u0.count + u1.count, 2), 256, 0, xs>>>(u0, u1, lower ? mine + 0 : nullptr, upper ? mine + 1 : nullptr, seq_, tmo);
            MVD_CUDA_CHECK(cudaGetLastError());
        } else {
            if (lower) { pack_rows(b, landing(nb_region_[0], 1, par), b.y0, b.hy_hi, 1, xs); pfor(1, SignalArrival{flags(nb_region_[0]) + 1, seq_}, xs); }
            if (upper) { pack_rows(b, landing(nb_region_[1], 0, par), b.y1 - b.hy_lo, b.hy_lo, 1, xs); pfor(1, SignalArrival{flags(nb_region_[1]) + 0, seq_}, xs); }
            if (lower || upper) pfor(1, AwaitArrival{lower ? mine + 0 : nullptr, upper ? mine + 1 : nullptr, seq_}, xs);
            if (lower) pack_rows(b, landing(region_, 0, par), b.y0 - b.hy_lo, b.hy_lo, 0, xs);
            if (upper) pack_rows(b, landing(region_, 1, par), b.y1, b.hy_hi, 0, xs);
        }
    }
    if (pz_ > 1 && (b.hz_lo > 0 || b.hz_hi > 0)) {
        const bool lower = rz_ > 0, upper = rz_ < pz_ - 1;
        const size_t plane = (size_t)b.row_floats * (size_t)b.nrows;
        if (plane * (size_t)std::max(b.hz_lo, b.hz_hi) > cap_z_) throw Error("halo exchange: landing buffer too small (z)");
        if (quads && (lower || upper)) {                   // whole planes including the fresh y halos: all rows
            HaloSeg p0 = make_seg(b, lower ? landing(nb_region_[2], 3, par) : nullptr, 0, b.nrows, b.z0, lower ? b.hz_hi : 0);
            HaloSeg p1 = make_seg(b, upper ? landing(nb_region_[3], 2, par) : nullptr, 0, b.nrows, b.z1 - b.hz_lo, upper ? b.hz_lo : 0);
            halo_push_kernel<<<halo_grid(p0.count + p1.count, 4), 256, 0, xs>>>(p0, p1, lower ? flags(nb_region_[2]) + 3 : nullptr,
                                                                                   upper ? flags(nb_region_[3]) + 2 : nullptr, seq_, done);
            HaloSeg u0 = make_seg(b, lower ? landing(region_, 2, par) : nullptr, 0, b.nrows, b.z0 - b.hz_lo, lower ? b.hz_lo : 0);
            HaloSeg u1 = make_seg(b, upper ? landing(region_, 3, par) : nullptr, 0, b.nrows, b.z1, upper ? b.hz_hi : 0);
            halo_pull_kernel<<<halo_grid(u0.count + u1.count, 2), 256, 0, xs>>>(u0, u1, lower ? mine + 2 : nullptr, upper ? mine + 3 : nullptr, seq_, tmo);
            MVD_CUDA_CHECK(cudaGetLastError());
        } else {
            // scalar path for boxes whose rows are not 16-byte granules (arbitrary bounding boxes): whole planes as dense rows
            const HaloBox pb = b;
            if (lower) { pack_planes(pb, landing(nb_region_[2], 3, par), b.z0, b.hz_hi, 1, xs); pfor(1, SignalArrival{flags(nb_region_[2]) + 3, seq_}, xs); }
            if (upper) { pack_planes(pb, landing(nb_region_[3], 2, par), b.z1 - b.hz_lo, b.hz_lo, 1, xs); pfor(1, SignalArrival{flags(nb_region_[3]) + 2, seq_}, xs); }
            if (lower || upper) pfor(1, AwaitArrival{lower ? mine + 2 : nullptr, upper ? mine + 3 : nullptr, seq_}, xs);
            if (lower) pack_planes(pb, landing(region_, 2, par), b.z0 - b.hz_lo, b.hz_lo, 0, xs);
            if (upper) pack_planes(pb, landing(region_, 3, par), b.z1, b.hz_hi, 0, xs);
        }
    }
#else
    (void)b; (void)xs;
#endif
}

}  // namespace mvd
