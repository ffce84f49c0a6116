// comm.cu -- see comm.h
#include <dlfcn.h>

#include <string>
#include <vector>

#include <cstdlib>

#include "comm.h"

namespace ib {
namespace comm {

namespace {

// the few NCCL entry points used, resolved with dlsym (ABI of nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMin = 3 };

struct Nccl {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    if (n.lib) return n;
    // prefer a libnccl already in the process (torch's bundled one), then the system one
    n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!n.lib) return n;
#define SYM(field, name) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name))
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.Send && n.Recv &&
           n.GroupStart && n.GroupEnd;
    return n;
}

int comm_fail(ibis_b200_ctx* c, const std::string& msg) {
    if (c) c->last_error = msg;
    return IBIS_B200_ECOMM;
}

#define NCCL_TRY(c, expr)                                                                     \
    do {                                                                                      \
        int r__ = (expr);                                                                     \
        if (r__ != ncclSuccess)                                                               \
            return comm_fail(c, std::string(#expr) + ": " +                                   \
                                    (nccl().GetErrorString ? nccl().GetErrorString(r__) : "NCCL error")); \
    } while (0)

}  // namespace

int unique_id(void* out) {
    Nccl& n = nccl();
    if (!n.ok) return IBIS_B200_ECOMM;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess) return IBIS_B200_ECOMM;
    memcpy(out, &id, sizeof id);
    return 0;
}

int init(ibis_b200_ctx* c, const void* id_bytes, int rank, int n_ranks) {
    if (n_ranks <= 1) { c->rank = 0; c->n_ranks = 1; return 0; }
    Nccl& n = nccl();
    if (!n.ok) return comm_fail(c, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    cudaSetDevice(c->device);
    NCCL_TRY(c, n.CommInitRank(&comm, n_ranks, id, rank));
    c->nccl_comm = comm;
    c->rank = rank;
    c->n_ranks = n_ranks;
    return 0;
}

int destroy(ibis_b200_ctx* c) {
    if (c->nccl_comm) {
        nccl().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr;
    }
    Halo& h = c->halo;
    if (h.d_send_cells) cudaFree(h.d_send_cells);
    if (h.d_recv_cells) cudaFree(h.d_recv_cells);
    if (h.d_send_buf) cudaFree(h.d_send_buf);
    if (h.d_recv_buf) cudaFree(h.d_recv_buf);
    for (unsigned char* p : h.peer_base)
        if (p) cudaIpcCloseMemHandle(p);
    if (h.d_ipc) cudaFree(h.d_ipc);
    if (h.d_counters) cudaFree(h.d_counters);
    h = Halo();
    return 0;
}

int allreduce(ibis_b200_ctx* c, double* d_buf, int count, int op) {
    if (c->n_ranks <= 1) return 0;
    if (!c->nccl_comm) return comm_fail(c, "allreduce before comm_init");
    NCCL_TRY(c, nccl().AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, op == OP_MIN ? ncclMin : ncclSum,
                                 static_cast<ncclComm_t>(c->nccl_comm), c->stream));
    return 0;
}

int halo_setup(ibis_b200_ctx* c, int n_peers, const int32_t* peer_ranks, const int64_t* send_offsets,
               const int64_t* send_cells, const int64_t* recv_offsets, const int64_t* recv_cells) {
    Halo& h = c->halo;
    h.n_peers = n_peers;
    h.peer_ranks.assign(peer_ranks, peer_ranks + n_peers);
    h.send_off.assign(send_offsets, send_offsets + n_peers + 1);
    h.recv_off.assign(recv_offsets, recv_offsets + n_peers + 1);
    h.n_send = h.send_off[n_peers];
    h.n_recv = h.recv_off[n_peers];
    std::vector<int> s(h.n_send), r(h.n_recv);
    // the lists arrive in reference numbering
    for (int64_t i = 0; i < h.n_send; i++) s[i] = c->ref2int[send_cells[i]];
    for (int64_t i = 0; i < h.n_recv; i++) r[i] = c->ref2int[recv_cells[i]];
    if (cudaMalloc(&h.d_send_cells, std::max<size_t>(s.size(), 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h.d_recv_cells, std::max<size_t>(r.size(), 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h.d_send_buf, std::max<size_t>(s.size(), 1) * 7 * 16) != cudaSuccess ||
        cudaMalloc(&h.d_recv_buf, std::max<size_t>(r.size(), 1) * 7 * 16) != cudaSuccess)
        return comm_fail(c, "out of device memory (halo buffers)");
    cudaMemcpyAsync(h.d_send_cells, s.data(), s.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(h.d_recv_cells, r.data(), r.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream);
    cudaStreamSynchronize(c->stream);
    return 0;
}

constexpr size_t FLAG_BYTES = 256;     // flags region of the exported allocation (<= 32 peers)

int p2p_export(ibis_b200_ctx* c, void* handle64, int64_t* n_recv) {
    Halo& h = c->halo;
    if (h.n_peers > (int)(FLAG_BYTES / 8)) return comm_fail(c, "too many peers for the peer-memory exchange");
    if (!h.d_ipc) {
        h.area_bytes = (size_t)std::max<int64_t>(h.n_recv, 1) * 7 * 16;
        cudaSetDevice(c->device);
        if (cudaMalloc(&h.d_ipc, FLAG_BYTES + 2 * h.area_bytes) != cudaSuccess ||
            cudaMalloc(&h.d_counters, sizeof(unsigned) * std::max(h.n_peers, 1)) != cudaSuccess)
            return comm_fail(c, "out of device memory (peer-memory halo area)");
        cudaMemsetAsync(h.d_ipc, 0, FLAG_BYTES, c->stream);
        cudaMemsetAsync(h.d_counters, 0, sizeof(unsigned) * std::max(h.n_peers, 1), c->stream);
        cudaStreamSynchronize(c->stream);
        h.peer_base.assign(h.n_peers, nullptr);
        h.peer_n_recv.assign(h.n_peers, 0);
        h.peer_my_off.assign(h.n_peers, 0);
        h.peer_my_slot.assign(h.n_peers, 0);
    }
    cudaIpcMemHandle_t mh;
    cudaError_t e = cudaIpcGetMemHandle(&mh, h.d_ipc);
    if (e != cudaSuccess) return comm_fail(c, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &mh, 64);
    if (n_recv) *n_recv = h.n_recv;
    return 0;
}

int p2p_import(ibis_b200_ctx* c, int p, const void* handle64, int64_t peer_n_recv, int64_t my_off, int my_slot) {
    Halo& h = c->halo;
    if (!h.d_ipc) return comm_fail(c, "halo_p2p_import before halo_p2p_export");
    if (p < 0 || p >= h.n_peers || my_slot < 0 || my_slot >= (int)(FLAG_BYTES / 8))
        return comm_fail(c, "halo_p2p_import: bad peer index");
    cudaIpcMemHandle_t mh;
    memcpy(&mh, handle64, 64);
    void* base = nullptr;
    cudaSetDevice(c->device);
    cudaError_t e = cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return comm_fail(c, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    h.peer_base[p] = static_cast<unsigned char*>(base);
    h.peer_n_recv[p] = peer_n_recv;
    h.peer_my_off[p] = my_off;
    h.peer_my_slot[p] = my_slot;
    return 0;
}

int p2p_enable(ibis_b200_ctx* c, int on) {
    Halo& h = c->halo;
    if (on)
        for (int p = 0; p < h.n_peers; p++)
            if (!h.peer_base.size() || !h.peer_base[p]) return comm_fail(c, "halo_p2p_enable: a peer is not imported");
    h.p2p = on != 0;
    return 0;
}

// The exchange without NCCL: every rank stores its boundary cells straight into its peers' receive
// areas (remote stores over NVLink from k_halo_push) and then waits, on the device, for its own
// flags (k_halo_pull).  Two receive areas alternate: a peer can only start exchange k + 2 after
// it has consumed my data of exchange k + 1, which I sent after my own pull of exchange k.
template <typename S>
static int halo_push_p2p(ibis_b200_ctx* c, S* const* fs) {
    Halo& h = c->halo;
    h.iter++;
    const int parity = (int)(h.iter & 1);
    for (int p = 0; p < h.n_peers; p++) {
        const int64_t so = h.send_off[p], sn = h.send_off[p + 1] - so;
        unsigned char* peer_area = h.peer_base[p] + FLAG_BYTES + (size_t)parity * ((size_t)std::max<int64_t>(h.peer_n_recv[p], 1) * 7 * 16);
        S* remote = reinterpret_cast<S*>(peer_area) + h.peer_my_off[p] * 7;
        unsigned long long* rflag = reinterpret_cast<unsigned long long*>(h.peer_base[p]) + h.peer_my_slot[p];
        launch_halo_push<S>(c, fs, h.d_send_cells + so, sn, remote, rflag, h.d_counters + p, h.iter);
    }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
template <typename S>
static int halo_pull_p2p(ibis_b200_ctx* c, S* const* fs) {
    Halo& h = c->halo;
    const int parity = (int)(h.iter & 1);
    for (int p = 0; p < h.n_peers; p++) {
        const int64_t ro = h.recv_off[p], rn = h.recv_off[p + 1] - ro;
        const S* area = reinterpret_cast<const S*>(h.d_ipc + FLAG_BYTES + (size_t)parity * h.area_bytes) + ro * 7;
        launch_halo_pull<S>(c, fs, h.d_recv_cells + ro, rn, area,
                            reinterpret_cast<const unsigned long long*>(h.d_ipc) + p, h.iter);
    }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

template <typename S>
int halo_exchange(ibis_b200_ctx* c, S* const* fs) {
    Halo& h = c->halo;
    if (h.n_peers == 0 || c->n_ranks <= 1) return 0;
    if (h.p2p) {
        int rc = halo_push_p2p<S>(c, fs);
        return rc ? rc : halo_pull_p2p<S>(c, fs);
    }
    if (!c->nccl_comm) return comm_fail(c, "halo exchange before comm_init");
    Nccl& n = nccl();
    S* sbuf = reinterpret_cast<S*>(h.d_send_buf);
    S* rbuf = reinterpret_cast<S*>(h.d_recv_buf);
    launch_halo_pack<S>(c, fs, h.d_send_cells, h.n_send, sbuf);
    const size_t per_cell = 7 * sizeof(S) / sizeof(double);     // doubles per cell
    NCCL_TRY(c, n.GroupStart());
    for (int p = 0; p < h.n_peers; p++) {
        const int64_t so = h.send_off[p], sn = h.send_off[p + 1] - so;
        const int64_t ro = h.recv_off[p], rn = h.recv_off[p + 1] - ro;
        if (sn > 0)
            NCCL_TRY(c, n.Send(reinterpret_cast<double*>(sbuf) + so * per_cell, (size_t)sn * per_cell, ncclFloat64,
                               h.peer_ranks[p], static_cast<ncclComm_t>(c->nccl_comm), c->stream));
        if (rn > 0)
            NCCL_TRY(c, n.Recv(reinterpret_cast<double*>(rbuf) + ro * per_cell, (size_t)rn * per_cell, ncclFloat64,
                               h.peer_ranks[p], static_cast<ncclComm_t>(c->nccl_comm), c->stream));
    }
    NCCL_TRY(c, n.GroupEnd());
    launch_halo_unpack<S>(c, fs, h.d_recv_cells, h.n_recv, rbuf);
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
bool halo_split(const ibis_b200_ctx* c) { return c->halo.p2p && c->halo.n_peers > 0 && c->n_ranks > 1; }
// The overlapped exchange: the pushes read the flow state and write into the peers' memory, nothing this rank's
// evaluation waits for -- they go onto a side stream (forked from the compute stream by an event) and run BESIDE the
// early tiles of pass G' (the persistent kernel leaves registers and thread slots free on every SM; the push kernels
// use no shared memory).  halo_exchange_join makes the compute stream wait for them before the flow state is
// overwritten (end of the evaluation).  IBIS_B200_PUSH_STREAM=0 keeps them on the compute stream.
template <typename S>
int halo_exchange_begin(ibis_b200_ctx* c, S* const* fs) {
    if (!halo_split(c)) return 0;
    Halo& h = c->halo;
    static const bool side = !(getenv("IBIS_B200_PUSH_STREAM") && getenv("IBIS_B200_PUSH_STREAM")[0] == '0');
    if (!side) return halo_push_p2p<S>(c, fs);
    if (!h.side_ready) {
        if (gpu::stream_create(&h.push_stream) != 0 || gpu::event_create(&h.fork) != 0 || gpu::event_create(&h.pushed) != 0)
            return comm_fail(c, "halo exchange: cannot create the push stream");
        h.side_ready = true;
    }
    if (gpu::event_record(h.fork, c->stream) != 0 || gpu::stream_wait_event(h.push_stream, h.fork) != 0)
        return comm_fail(c, "halo exchange: fork failed");
    gpu::stream_t main_stream = c->stream;
    c->stream = h.push_stream;                       // the launchers take the context's stream
    int rc = halo_push_p2p<S>(c, fs);
    c->stream = main_stream;
    if (gpu::event_record(h.pushed, h.push_stream) != 0) return comm_fail(c, "halo exchange: join event failed");
    h.push_pending = true;
    return rc;
}
int halo_exchange_join(ibis_b200_ctx* c) {
    Halo& h = c->halo;
    if (!h.push_pending) return 0;
    h.push_pending = false;
    return gpu::stream_wait_event(c->stream, h.pushed) != 0 ? comm_fail(c, "halo exchange: join failed") : 0;
}
template <typename S>
int halo_exchange_end(ibis_b200_ctx* c, S* const* fs) {
    return halo_split(c) ? halo_pull_p2p<S>(c, fs) : halo_exchange<S>(c, fs);
}
template int halo_exchange_begin<double>(ibis_b200_ctx*, double* const*);
template int halo_exchange_begin<dual>(ibis_b200_ctx*, dual* const*);
template int halo_exchange_end<double>(ibis_b200_ctx*, double* const*);
template int halo_exchange_end<dual>(ibis_b200_ctx*, dual* const*);
template int halo_exchange<double>(ibis_b200_ctx*, double* const*);
template int halo_exchange<dual>(ibis_b200_ctx*, dual* const*);

}  // namespace comm
}  // namespace ib
