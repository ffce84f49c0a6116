// context.h -- host-side state behind the opaque ibis_b200_ctx and the launcher
// interface between the C ABI (abi.cu) and the kernel translation units.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

#include "backend.h"
#include "fv_bodies.h"
#include "solver_bodies.h"
#include "ibis_b200.h"

namespace ib {
// records msg as the thread's ibis_b200_last_error(NULL) and returns code (abi.cu)
int fail_global(int code, const char* msg);

enum BufKind { BUF_FREE = 0, BUF_FS = 1, BUF_CQ = 2, BUF_VEC = 3, BUF_BASIS = 4 };

struct Buffer {
    BufKind kind = BUF_FREE;
    bool is_dual = false;
    int ncomp = 0;         // FS 7, CQ 5 (internal: mz slot always present), VEC 1, BASIS n_cols
    int64_t n = 0;         // entries per component
    int64_t stride = 0;    // component stride (entries)
    void* d = nullptr;     // double* or dual*
};

struct Halo {
    int n_peers = 0;
    std::vector<int> peer_ranks;
    std::vector<int64_t> send_off, recv_off;   // per peer offsets into the index lists
    int* d_send_cells = nullptr;               // internal ids
    int* d_recv_cells = nullptr;
    double* d_send_buf = nullptr;              // packed [peer][var][cell], real or dual
    double* d_recv_buf = nullptr;
    int64_t n_send = 0, n_recv = 0;
    // peer-memory exchange over NVLink (comm.cu): this rank exports ONE allocation
    //   [flags: one uint64 per peer, padded to 256 B][receive area, parity 0][receive area, parity 1]
    // and maps the peers' allocations; a receive area holds n_recv cells x 7 rows of dual size,
    // peer segment p at recv_off[p], stored [row][cell of the segment]
    bool p2p = false;
    unsigned char* d_ipc = nullptr;
    size_t area_bytes = 0;                     // one parity
    std::vector<unsigned char*> peer_base;     // the peers' allocations as mapped here
    std::vector<int64_t> peer_n_recv;          // the peer's n_recv (locates its parity-1 area)
    std::vector<int64_t> peer_my_off;          // offset (cells) of my segment in the peer's receive order
    std::vector<int> peer_my_slot;             // my index in the peer's peer list (its flag for me)
    unsigned* d_counters = nullptr;            // last-block counters, one per peer
    unsigned long long iter = 0;               // exchanges done (same on every rank)
    // the pushes of the overlapped exchange run on their own stream, beside the early tiles of pass G'
    ib::gpu::stream_t push_stream = nullptr;
    ib::gpu::event_t fork, pushed;
    bool side_ready = false, push_pending = false;
};

}  // namespace ib

struct ibis_b200_ctx {
    int device = 0;
    int n_sm = 148;                      // multiprocessors of the device (persistent grids)
    ib::gpu::stream_t stream = nullptr;
    std::string last_error;              // sticky: a device / memory / communication failure (abi.cu::fail)
    std::string last_message;            // the most recent failure of any kind, argument errors included
    int64_t launches = 0;

    // sizes
    int dim = 0, n_cons = 0, n_tags = 0;
    int64_t n_owned = 0, n_cells = 0, n_ghost = 0, n_total = 0, n_faces = 0, n_bfaces = 0;
    int64_t cstride = 0, fstride = 0;
    ibis_b200_config cfg;
    std::vector<ibis_b200_bc_action> bc_actions;
    bool has_previscous = false;
    bool has_flux_actions = false;
    // moving grid (ibis_b200_set_face_vertices / _set_vertex_velocities / _compute_dudt_moving)
    std::vector<int> face_ref2int;       // reference interface id -> internal id
    int* d_fverts = nullptr;             // [4][fstride] vertex ids of every interface, internal numbering, -1 padded
    int64_t n_vertices = 0;
    double* d_vertex_vel = nullptr;      // n_vertices x 3
    double* d_face_vel = nullptr;        // [3][fstride]
    bool have_vertex_vel = false;
    int* d_cverts = nullptr;             // [8][cstride] vertex ids of every valid cell (internal numbering)
    int* d_ctype = nullptr;
    double* d_vertex_pos = nullptr;      // n_vertices x 3
    bool have_vertex_pos = false;
    bool geometry_moved = false;         // the fused static paths hold tables of the ORIGINAL geometry       // any constant_flux: evaluations go through the face-flux array

    // internal renumbering (layout.h): reference cell id -> internal id, nullptr = identity
    int* d_ref2int = nullptr;
    std::vector<int> ref2int;
    ib::GridDev g;                       // device pointers + sizes
    std::vector<void*> owned_allocs;     // everything to free at destroy

    // FiniteVolume-owned temporaries (finite_volume.h:116-141), real and dual
    double* grad_r = nullptr;
    double* lim_r = nullptr;
    double* flux_r = nullptr;
    ib::dual* grad_d = nullptr;
    ib::dual* lim_d = nullptr;
    ib::dual* flux_d = nullptr;
    double* face_T = nullptr;
    double* face_rho = nullptr;
    // interface-side records of the sides path (fv_bodies.h), real and dual
    double* sidesP_r = nullptr;
    ib::dual* sidesP_d = nullptr;
    int64_t p_total = 0;                 // records in a side-state array (all owned tiles)
    int pipe_prefetch = 2;               // IBIS_B200_PREFETCH: L2 prefetch distance (tiles) of the pipelined kernels, 0 = off
    int pipe_f_stages = 0;               // input stages of pass F' (0: default; IBIS_B200_F_STAGES)
    bool use_sides = true;               // IBIS_B200_SIDES=0: second order through the gather path (grad + limiter arrays)
    bool dual_sides = true;               // IBIS_B200_DUAL_SIDES=0: dual-number evaluations through the gather path
    int last_fs = 0;                     // FlowStates handle of the last real evaluation (debug_gradients)

    // reductions
    double* d_partials = nullptr;        // [n_blocks * 8]
    double* d_scalars = nullptr;         // results of final reductions (64 doubles)
    double* h_scalars = nullptr;         // pinned mirror
    int n_red_blocks = 0;
    unsigned int* d_pipe_dbg = nullptr;  // abort flag + stall record of the pipelined kernels (FvArgs::pipe_dbg), 8 words
    unsigned int* h_pipe_dbg = nullptr;  // pinned mirror

    // I/O staging (reference AoS <-> internal SoA)
    double* d_stage = nullptr;
    int64_t stage_elems = 0;
    // snapshot ring (abi.cu): two upload and two download slots of 5 x n_cells doubles, a stream for each direction
    struct SnapshotRing {
        bool ready = false;
        ib::gpu::stream_t up = nullptr, down = nullptr;
        double* d_up[2] = {nullptr, nullptr};
        double* d_down[2] = {nullptr, nullptr};
        ib::gpu::event_t uploaded[2], unpacked[2], packed[2], downloaded[2];
        bool up_used[2] = {false, false}, down_used[2] = {false, false};
    } ring;

    // explicit solver scratch: k_[stages], k_tmp_, flow_tmp_ (runge_kutta.cpp:45-52)
    std::vector<int> rk_k;            // handle of stage i (all equal when the stages alias)
    std::vector<int> rk_k_own;        // the buffers actually allocated
    bool rk_alias_stages = true;      // the reference's k_ stages share one array
    bool rk_unfused = false;          // IBIS_B200_RK_UNFUSED=1: one kernel per apply_time_derivative / conversion
    int rk_ktmp = 0, rk_fstmp = 0;

    // optional per-kernel event timing (ibis_b200_profile_*)
    bool prof_on = false;
    // parity knob (IBIS_B200_ORDERED_DOT=1): dot / norm2 summed in index order by one thread, the
    // order of the reference's parallel_reduce with one OpenMP thread (tests of ill-conditioned
    // Krylov solves); the default is the two-stage tree reduction
    bool ordered_dot = false;
    std::vector<void*> prof_ev[4];       // event pairs (start, stop, start, stop, ...) per kind
    size_t prof_used[4] = {0, 0, 0, 0};

    std::vector<ib::Buffer> bufs;        // handle = index (0 unused)
    ib::Halo halo;
    void* nccl_comm = nullptr;
    int rank = 0, n_ranks = 1;
};

namespace ib {

// ---- launchers implemented in fv_kernels.cu (CUDA) / tests/emu (host loops) ----
template <typename S>
void launch_bc(ibis_b200_ctx* c, const FvArgs<S>& A, int phase);
void launch_wls_r(ibis_b200_ctx* c, double* r_out);
void launch_cell_static(ibis_b200_ctx* c, double* w_out, double* d_out);   // cf_w, cf_d (after wls_r)
void launch_pack_pipe_tables(ibis_b200_ctx* c, double* t_wd);   // tile-major cf_w / cf_d
template <typename S>
void launch_grad_limit(ibis_b200_ctx* c, const FvArgs<S>& A, bool viscous_pass);
template <typename S>
void launch_residual(ibis_b200_ctx* c, const FvArgs<S>& A);        // fused inviscid path (order 1)
template <typename S>
void launch_grad_sides(ibis_b200_ctx* c, const FvArgs<S>& A, int which);   // sides path, pass G' (0 all | 1 early | 2 late tiles)
template <typename S>
void launch_flux_tile(ibis_b200_ctx* c, const FvArgs<S>& A);       // sides path, pass F'
// fused dt reduction in pass F': before the launch / after it (device: integer keys; emulation: doubles)
void launch_dt_reduce_init(ibis_b200_ctx* c, double* p);
void launch_dt_reduce_finish(ibis_b200_ctx* c, double* p);
template <typename S>
void launch_face_flux(ibis_b200_ctx* c, const FvArgs<S>& A);       // general path
template <typename S>
void launch_face_flux_moving(ibis_b200_ctx* c, const FvArgs<S>& A);   // moving grid: relative to A.face_vel
void launch_face_vel(ibis_b200_ctx* c, const int* fverts, const double* vvel, double* face_vel);   // compute_face_vel
void launch_geom_update(ibis_b200_ctx* c, const GeomUpdate& U);       // compute_geometric_data after a vertex move
void launch_move_vertices(ibis_b200_ctx* c, double* pos, const double* vel, double scale, int64_t n3);
template <typename S>
void launch_gcl(ibis_b200_ctx* c, const FvArgs<S>& A, S* const* cq);   // geometric conservation law
template <typename S>
void launch_face_viscous(ibis_b200_ctx* c, const FvArgs<S>& A);
template <typename S>
void launch_integral(ibis_b200_ctx* c, const FvArgs<S>& A);
template <typename S>
void launch_local_dt_min(ibis_b200_ctx* c, const FvArgs<S>& A, double* d_out);

// ---- launchers implemented in solver_kernels.cu ----
template <typename S>
void launch_cons_to_prim(ibis_b200_ctx* c, S* const* cq, S* const* fs, int64_t n);
template <typename S>
void launch_prim_to_cons(ibis_b200_ctx* c, S* const* cq, S* const* fs, int64_t n);
template <typename S>
void launch_apply_dt(ibis_b200_ctx* c, S* const* u0, S* const* u1, S* const* k, double dt,
                     const double* d_dt, double scale, int64_t n);
template <typename S>
void launch_rk_update(ibis_b200_ctx* c, S* const* u_in, const RkTerms<S>& T, const double* d_dt, S* const* u_out,
                      S* const* fs, int64_t n);     // fused stage / final update + cons_to_prim (u_out may be null)
template <typename S>
void launch_l2_norms(ibis_b200_ctx* c, S* const* cq, int64_t n, double* d_out6);
template <typename S>
void launch_count_bad(ibis_b200_ctx* c, S* const* fs, int64_t n, double* d_out);
void launch_rk_dt(ibis_b200_ctx* c, const double* d_dt_stable, double cfl, double dt_prev,
                  double t_remaining, double* d_dt);
// AoS (reference numbering) <-> SoA (internal), ncomp_ref components per row
template <typename S>
void launch_pack_in(ibis_b200_ctx* c, const double* d_re, const double* d_du, S* const* comps,
                    const int* comp_map, int ncomp_ref, int64_t n);
template <typename S>
void launch_pack_out(ibis_b200_ctx* c, S* const* comps, const int* comp_map, int ncomp_ref,
                     int64_t n, double* d_re, double* d_du);
template <typename S>
void launch_flow_read(ibis_b200_ctx* c, const double* stage, int64_t n, int has_vz, S* const* fs);
template <typename S>
void launch_flow_write(ibis_b200_ctx* c, S* const* fs, double* stage, int64_t n, int has_vz);
void launch_gather_debug(ibis_b200_ctx* c, const double* src, int ncomp, int64_t n, double* out);
void launch_dual_kat(ibis_b200_ctx* c, const double* in, double* out);   // solver_bodies.h::dual_kat_body, one thread
void launch_division_check(ibis_b200_ctx* c, int64_t n, uint64_t seed, double* d_out);   // mismatches of div_pair / div_shared / div_z
// vectors
void launch_dot(ibis_b200_ctx* c, const double* a, const double* b, int64_t n, double* d_out,
                int post);   // post: 0 none, 1 sqrt, 2 sqrt and also store 1/sqrt at d_out[1]
void launch_scalar_post(ibis_b200_ctx* c, double* d_out, int post);   // the post step alone, after a cross-rank sum
void launch_axpy(ibis_b200_ctx* c, double* y, const double* x, double s, const double* d_s,
                 double sign, int64_t n);
void launch_scale(ibis_b200_ctx* c, const double* x, double* r, double s, const double* d_s,
                  int64_t n);
void launch_sub(ibis_b200_ctx* c, const double* a, const double* b, double* r, int64_t n);
void launch_gemv(ibis_b200_ctx* c, const double* basis, int64_t n, int64_t stride, int m,
                 const double* d_y, double* out);
void launch_fill(ibis_b200_ctx* c, double* p, double v, int64_t n);
// JFNK
void launch_jfnk_set_dual(ibis_b200_ctx* c, dual* const* cq, const double* vec, dual* const* cq_tmp);
void launch_jfnk_set_vector(ibis_b200_ctx* c, dual* const* res, const double* vec, double dt_star,
                            double* result);
void launch_jfnk_eval_rhs(ibis_b200_ctx* c, dual* const* res, double* rhs);
void launch_jfnk_apply_update(ibis_b200_ctx* c, dual* const* cq, const double* dU);
// halo
template <typename S>
void launch_halo_pack(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, S* buf);
template <typename S>
void launch_halo_unpack(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, const S* buf);
// peer-memory exchange: gather n cells and store them, [row][cell], straight into the peer's receive
// area (remote), then release-store `iter` to the peer's flag from the last block to finish
template <typename S>
void launch_halo_push(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, S* remote,
                      unsigned long long* remote_flag, unsigned* counter, unsigned long long iter);
// wait (acquire) until the local flag reaches `iter`, then scatter the received segment
template <typename S>
void launch_halo_pull(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, const S* area,
                      const unsigned long long* flag, unsigned long long iter);

int reduction_blocks(int64_t n);
// event bracket around a launch when profiling is on (kind 0 = pass G, 1 = pass F, 2 = halo
// exchange, 3 = everything else of an rk_step: stage updates, conversions, dt)
void prof_begin(ibis_b200_ctx* c, int kind);
void prof_end(ibis_b200_ctx* c, int kind);
int prof_read(ibis_b200_ctx* c, double* ms, int64_t* count);

}  // namespace ib
