// comm.h -- domain-decomposition plumbing of one rank: ghost-cell halo exchange of the
// FlowStates rows and the scalar reductions (MIN for dt, SUM for norms / dots / bad-cell
// counts) over NCCL on the context's stream -- or, for the halo rows, direct stores into the peer's
// memory over NVLink (p2p_*: CUDA IPC mapping, device-side release/acquire flags, no NCCL call).  NCCL is bound at run time (dlopen of the
// libnccl.so.2 already loaded by torch.distributed, or the system one), so single-GPU
// runs have no NCCL dependency at all.  No reference counterpart (the reference is a
// single block on a single device).
#pragma once

#include "context.h"

namespace ib {
namespace comm {

enum { OP_SUM = 0, OP_MIN = 1 };

int init(ibis_b200_ctx* c, const void* unique_id, int rank, int n_ranks);
int destroy(ibis_b200_ctx* c);
int unique_id(void* out_128_bytes);
int allreduce(ibis_b200_ctx* c, double* d_buf, int count, int op);   // no-op on one rank
int halo_setup(ibis_b200_ctx* c, int n_peers, const int32_t* peer_ranks, const int64_t* send_offsets,
               const int64_t* send_cells, const int64_t* recv_offsets, const int64_t* recv_cells);
template <typename S>
int halo_exchange(ibis_b200_ctx* c, S* const* fs);                   // no-op without peers
// split form for overlap: begin = this rank's stores into its peers (returns at once), end = wait for
// the peers' data and scatter it; halo_split() says whether the split form does anything useful (peer
// memory enabled).  Between the two, kernels that touch no halo cell may run.
template <typename S>
int halo_exchange_begin(ibis_b200_ctx* c, S* const* fs);
template <typename S>
int halo_exchange_end(ibis_b200_ctx* c, S* const* fs);
bool halo_split(const ibis_b200_ctx* c);
// the compute stream waits for the pushes of halo_exchange_begin (they run on a side stream): call before the flow
// state they read is overwritten
int halo_exchange_join(ibis_b200_ctx* c);
// peer-memory (NVLink P2P) data path: export this rank's receive allocation, map a peer's, switch on
int p2p_export(ibis_b200_ctx* c, void* handle64, int64_t* n_recv);
int p2p_import(ibis_b200_ctx* c, int peer_index, const void* handle64, int64_t peer_n_recv,
               int64_t my_offset_in_peer, int my_slot_in_peer);
int p2p_enable(ibis_b200_ctx* c, int on);

}  // namespace comm
}  // namespace ib
