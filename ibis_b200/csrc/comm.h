// comm.h -- domain-decomposition plumbing of one rank: ghost-cell halo exchange of the
// FlowStates rows and the scalar reductions (MIN for dt, SUM for norms / dots / bad-cell
// counts) over NCCL on the context's stream.  NCCL is bound at run time (dlopen of the
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

}  // namespace comm
}  // namespace ib
