// fv_pipe.cu -- the sides path of the inviscid second-order residual as two PERSISTENT, warp-specialised
// sm_100a kernels (one block per SM):
//
//   k_grad_pipe  pass G': WLS gradients (grid/gradient.h:66-116), Barth-Jespersen limiters (limiter.cpp:24-63)
//                and the MUSCL extrapolation + thermodynamic closure of linear_reconstruct
//                (convective_flux.cpp:243-420, gas_model.h:91-117) -> one 48-byte side record per interface side
//   k_flux_pipe  pass F': rotate, flux calculator, rotate back (convective_flux.cpp:148-164) and
//                flux_surface_integral (finite_volume.cpp:221-256), owner computes, no atomics on the residual
//
// Both have the same anatomy.  ONE loader warp walks the block's tiles and keeps a ring of shared-memory
// stages full: every static table of a tile is a packed per-tile record (GridDev::t_wd, t_grec, t_frec) or a
// contiguous interface range (tile-owned interface numbering, layout.h), so a stage is filled by a handful of
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier; the only gathered data -- the state of
// the neighbour cells outside the tile -- is fetched by the loader's lanes with cp.async into the same stage and
// reported to the same mbarrier.  The CONSUMER warps never touch global memory on the input side: they wait on the
// stage's mbarrier, take work items (32 interfaces, or 32 (cell, variable) pairs) off a per-stage counter, and the
// warp that finishes a tile's last item does the tile's epilogue (cell sums / bulk store of the side records).
// No block-wide barrier after start-up; warps of one block are at different tiles at the same time, which is what
// hides the latency of the fp64 dependency chains and of the loads.
//
// Arithmetic: the same per-value expression trees as the reference (and as the host emulation in tests/emu runs
// them through fv_bodies.h::pipe_grad_cell / flux_from_sides6 / pipe_cell_sum), hence the same bits.
#include <cuda_runtime.h>

#include "async_copy.h"
#include "context.h"

#ifndef IBIS_PIPE_F_WARPS
#define IBIS_PIPE_F_WARPS 19      // consumer warps of pass F'
#endif
#ifndef IBIS_PIPE_G_WARPS
#define IBIS_PIPE_G_WARPS 19      // consumer warps of pass G'
#endif
#ifndef IBIS_PIPE_LOOKAHEAD
#define IBIS_PIPE_LOOKAHEAD 2     // pass G': the static record is requested this many tiles ahead of the gather
#endif

namespace ib {

#define CUDA_CHECK_LAUNCH(c)                                                             \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess && (c)->last_error.empty())                               \
            (c)->last_error = std::string("kernel launch: ") + cudaGetErrorString(e__);  \
        (c)->launches++;                                                                 \
    } while (0)

constexpr int PIPE_MAX_STAGES = 8;
constexpr size_t PIPE_SMEM_BUDGET = 227 * 1024 - 1024;   // dynamic shared memory of one block (static barriers aside)

// Bounded waits.  Both return false -- for the whole warp -- when the pipeline has been aborted: by this warp,
// after PIPE_TIMEOUT_CLOCKS in this one wait (it records which wait in dbg[1..5]), or by another.  Every caller
// leaves the kernel on false, so a protocol error surfaces as an error at the next synchronise
// (abi.cu::check_pipe_stall) instead of a hung stream.
enum { PW_FULL = 1, PW_EMPTY = 2, PW_REC = 3, PW_OUT = 4, PW_STAGE = 5, PW_DONE = 6 };
constexpr long long PIPE_TIMEOUT_CLOCKS = 6000000000ll;          // about 3 s
__device__ __forceinline__ bool pipe_timed_out(long long t0, unsigned int* dbg, int kind, int k, unsigned aux) {
    if (*reinterpret_cast<volatile unsigned int*>(dbg) != 0u) return true;
    if (clock64() - t0 > PIPE_TIMEOUT_CLOCKS) {
        if (atomicCAS(dbg, 0u, 1u) == 0u) {
            dbg[1] = (unsigned)kind; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x >> 5; dbg[4] = (unsigned)k; dbg[5] = aux;
            __threadfence();
        }
        return true;
    }
    return false;
}
// the phase of an mbarrier (the hardware suspends the warp inside try_wait)
__device__ __forceinline__ bool pipe_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, int kind, int k) {
    bool ok = ac::mbar_try_wait(bar, parity);
    if (!ok) {
        const long long t0 = clock64();
        for (int it = 1;; it++) {
            if (ac::mbar_try_wait(bar, parity)) { ok = true; break; }
            if ((it & 15) == 0 && pipe_timed_out(t0, dbg, kind, k, parity)) break;
        }
    }
    return __all_sync(0xffffffffu, ok);
}
// a monotone counter in shared memory reaching `target` (no parity, hence no aliasing when a warp is late)
__device__ __forceinline__ bool pipe_spin_ge(const int* p, int target, unsigned int* dbg, int kind, int k) {
    const volatile int* vp = p;
    bool ok = *vp >= target;
    if (!ok) {
        const long long t0 = clock64();
        for (int it = 1;; it++) {
            __nanosleep(64);
            if (*vp >= target) { ok = true; break; }
            if ((it & 63) == 0 && pipe_timed_out(t0, dbg, kind, k, (unsigned)target)) break;
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_block();
    return ok;
}
// next item of the block's ordered work queue
__device__ __forceinline__ int pipe_next_item(int* head, int lane) {
    int i = 0;
    if (lane == 0) i = atomicAdd(head, 1);
    return __shfl_sync(0xffffffffu, i, 0);
}

// a 64-int tile descriptor held by a warp: lane l has ints l and 32 + l
struct TileDesc {
    int d0, d1;
};
__device__ __forceinline__ TileDesc td_load(const int* ti, int lane) {
    TileDesc d;
    d.d0 = __ldg(ti + lane);
    d.d1 = __ldg(ti + 32 + lane);
    return d;
}
// int idx of the descriptor (converged call; idx may differ between the lanes)
__device__ __forceinline__ int td_get(const TileDesc& d, int idx) {
    const int a = __shfl_sync(0xffffffffu, d.d0, idx & 31), b = __shfl_sync(0xffffffffu, d.d1, idx & 31);
    return idx < 32 ? a : b;
}

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w < v) ? w : v;
    }
    return v;
}
__device__ __forceinline__ double shfl_xor_s(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ dual shfl_xor_s(const dual& v, int m) {
    return dual(__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.du, m));
}

// ---------------------------------------------------------------------------------------------------------
// setup: tile-major copies of the static tables
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_wd(const GridDev g, double* t_wd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.n_grad) return;
    const int T = g.tile, NT = g.pipe_nf * 3;
    const bool owned = c < g.n_owned;
    const int tile = owned ? c / T : g.n_tiles + (c - g.n_owned) / T;
    const int i = owned ? c % T : (c - g.n_owned) % T;
    double* dst = t_wd + (size_t)tile * 2 * NT * T + i;
    for (int r = 0; r < NT; r++) {
        dst[(size_t)r * T] = r < g.max_faces * 3 ? g.cf_w[(size_t)r * g.cstride + c] : 0.0;
        dst[(size_t)(NT + r) * T] = g.cf_d[(size_t)r * g.cstride + c];
    }
}
__global__ void __launch_bounds__(256) k_pack_frame6(const GridDev g, double* fr6) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= g.n_faces) return;
    for (int k = 0; k < 3; k++) {
        fr6[(size_t)f * 6 + k] = g.f_t1[(size_t)k * g.fstride + f];
        fr6[(size_t)f * 6 + 3 + k] = g.f_t2[(size_t)k * g.fstride + f];
    }
}
void launch_pack_pipe_tables(ibis_b200_ctx* c, double* t_wd, double* f_frame6) {
    const GridDev& g = c->g;
    if (g.n_grad > 0) k_pack_wd<<<(g.n_grad + 255) / 256, 256, 0, c->stream>>>(g, t_wd);
    if (g.n_faces > 0) k_pack_frame6<<<(g.n_faces + 255) / 256, 256, 0, c->stream>>>(g, f_frame6);
    CUDA_CHECK_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------------------
// pass F'
// ---------------------------------------------------------------------------------------------------------
// stage = [left sides cap x 6 S | right sides cap x 6 S | frames cap x 6 double | cell record frec_stride |
//          (dt fused into this evaluation) vx vy vz T of the tile's cells, 4 x T S]
template <typename S>
__host__ __device__ inline size_t flux_stage_bytes(int cap, int frec_stride, int T, bool with_dt) {
    size_t b = (size_t)cap * frec<S>::bytes + (size_t)frec_stride + (with_dt ? (size_t)4 * T * sizeof(S) : 0);
    return (b + 127) / 128 * 128;
}

// Work items of a block, handed out in this order by one counter: for its m-th tile the NCH = ceil(cap / 32) flux
// chunks (32 interfaces each), then the T / 32 cell parts of tile m - 1 (one lane per cell sums its faces; queued
// behind the NEXT tile's chunks so that the fluxes they need are normally complete when a warp gets there).  A warp
// holding an item of tile k waits for stage_tile[s] >= k (the loader has re-designated the stage) before it tests
// the stage barrier's parity: the stage cannot be recycled while an item of its tile is outstanding, so the barrier is
// then exactly at the awaited phase or one behind, never two.
template <int DIM, int FLUX, typename S>
__global__ void __launch_bounds__(32 * (IBIS_PIPE_F_WARPS + 1), 1) k_flux_pipe(const FvArgs<S> A, const int nst) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[PIPE_MAX_STAGES], empty[PIPE_MAX_STAGES];
    __shared__ int stage_tile[PIPE_MAX_STAGES], done_items[PIPE_MAX_STAGES], meta[PIPE_MAX_STAGES][4], q_head;
    const GridDev& g = A.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;
    const int cap = g.tile_cap, T = g.tile;
    const bool with_dt = A.dt_out != nullptr;
    const size_t stage_bytes = flux_stage_bytes<S>(cap, g.frec_stride, T, with_dt);
    const size_t offR = (size_t)cap * 6 * sizeof(S), offF = 2 * offR, offRec = offF + (size_t)cap * 48,
                 offDt = offRec + (size_t)g.frec_stride;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; s++) {
            ac::mbar_init(&full[s], 1);
            ac::mbar_init(&empty[s], 1);
            stage_tile[s] = -1;
            done_items[s] = 0;
        }
        q_head = 0;
        ac::fence_barrier_init();
    }
    __syncthreads();
    const int n_my = (int)blockIdx.x < g.n_tiles ? (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == ncw) {
        // ---- loader: the descriptor of the next tile is in registers before this tile's stage is waited for ----
        TileDesc cur = td_load(g.tile_info + (size_t)TILE_INFO * blockIdx.x, lane);
        for (int k = 0; k < n_my; k++) {
            const int t = (int)blockIdx.x + k * (int)gridDim.x, s = k % nst;
            const bool more = k + 1 < n_my;
            const TileDesc nxt = more ? td_load(g.tile_info + (size_t)TILE_INFO * (t + gridDim.x), lane) : cur;
            const int h0 = td_get(cur, 2), nh = td_get(cur, 3), ent_off = td_get(cur, 6), n_ent = td_get(cur, 7),
                      n_slots = td_get(cur, 9);
            if (k >= nst && !pipe_wait(&empty[s], ((k / nst) - 1) & 1, A.pipe_dbg, PW_EMPTY, k)) return;
            int c0, nc;
            tile_cells(g, t, c0, nc);
            unsigned char* st = smem + (size_t)s * stage_bytes;
            S* stL = reinterpret_cast<S*>(st);
            S* stR = reinterpret_cast<S*>(st + offR);
            double* stF = reinterpret_cast<double*>(st + offF);
            if (lane == 0) {
                meta[s][0] = n_slots;
                meta[s][1] = nc;
                meta[s][2] = c0;
                __threadfence_block();
                *reinterpret_cast<volatile int*>(&stage_tile[s]) = k;
                const uint32_t bytes = (uint32_t)n_slots * (uint32_t)frec<S>::bytes + (uint32_t)g.frec_stride +
                                       (with_dt ? 4u * (uint32_t)T * (uint32_t)sizeof(S) : 0u);
                ac::mbar_arrive_expect_tx(&full[s], bytes);
            }
            __syncwarp();
            const int n_q = 3 * (1 + n_ent);
            const bool inl = n_ent <= TILE_INLINE_ENT;
            for (int q0 = 0; q0 < n_q; q0 += 32) {
                const int q = q0 + lane, e = q / 3 - 1, which = q % 3;
                int idx = 12 + 5 * (e < 0 ? 0 : e);
                idx = idx > 59 ? 59 : idx;
                int f0 = td_get(cur, idx), count = td_get(cur, idx + 1), slot = td_get(cur, idx + 3);
                if (!inl && e >= 0 && q < n_q) {
                    const int* te = g.tile_ent + 5 * (size_t)(ent_off + e);
                    f0 = __ldg(te); count = __ldg(te + 1); slot = __ldg(te + 3);
                }
                if (e < 0) { f0 = h0; count = nh; slot = 0; }
                if (q >= n_q || count <= 0) continue;
                const uint32_t bytes = (uint32_t)count * (which == 2 ? 48u : 6u * (uint32_t)sizeof(S));
                if (which == 0) ac::bulk_g2s(stL + (long)slot * 6, A.sideL + (long)f0 * 6, bytes, &full[s]);
                else if (which == 1) ac::bulk_g2s(stR + (long)slot * 6, A.sideR + (long)f0 * 6, bytes, &full[s]);
                else ac::bulk_g2s(stF + (long)slot * 6, g.f_frame6 + (long)f0 * 6, bytes, &full[s]);
            }
            if (lane == 31) ac::bulk_g2s(st + offRec, g.t_frec + (size_t)t * g.frec_stride, (uint32_t)g.frec_stride, &full[s]);
            if (with_dt && lane >= 27 && lane < 31) {
                const int v = lane - 27;
                const S* src = A.fs[v < 3 ? F_VX + v : F_T] + c0;
                ac::bulk_g2s(st + offDt + (size_t)v * T * sizeof(S), src, (uint32_t)T * (uint32_t)sizeof(S), &full[s]);
            }
            cur = nxt;
        }
        return;
    }
    // ---- consumers ----
    const int NCH = (cap + 31) >> 5, NCL = (T + 31) >> 5, P = NCH + NCL;
    const int total = (n_my + 1) * P;
    for (;;) {
        const int item = pipe_next_item(&q_head, lane);
        if (item >= total) break;
        const int m = item / P, r = item % P;
        const bool is_flux = r < NCH;
        const int k = is_flux ? m : m - 1;
        if (k < 0 || k >= n_my) continue;
        const int s = k % nst;
        if (!pipe_spin_ge(&stage_tile[s], k, A.pipe_dbg, PW_STAGE, k)) return;
        if (!pipe_wait(&full[s], (k / nst) & 1, A.pipe_dbg, PW_FULL, k)) return;
        unsigned char* st = smem + (size_t)s * stage_bytes;
        S* stL = reinterpret_cast<S*>(st);
        const double* stF = reinterpret_cast<const double*>(st + offF);
        if (is_flux) {
            const S* stR = reinterpret_cast<const S*>(st + offR);
            const int slot = r * 32 + lane;
            if (slot < meta[s][0])
                flux_from_sides6<DIM, FLUX>(A, stL + (long)slot * 6, stR + (long)slot * 6, stF + (long)slot * 6, stL + (long)slot * 6);
            ac::fence_proxy_async();        // the flux written over the left side precedes the next bulk copy into the stage
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&done_items[s], 1);
            }
        } else {
            // every flux of the tile must be in shared memory
            if (!pipe_spin_ge(&done_items[s], NCH, A.pipe_dbg, PW_DONE, k)) return;
            const int nc = meta[s][1], c0 = meta[s][2];
            const FRec Rc = frec_view(st + offRec, T);
            const int i = (r - NCH) * 32 + lane;
            if (i < nc) pipe_cell_sum<DIM>(A, c0 + i, i, T, Rc, stL);
            if (with_dt) {                      // estimate_dt fused into this evaluation (uniform branch)
                const S* dtx = reinterpret_cast<const S*>(st + offDt);
                double dl = INFINITY;
                if (i < nc)
                    dl = pipe_cell_dt(re(dtx[i]), re(dtx[T + i]), re(dtx[2 * T + i]), re(dtx[3 * T + i]), re(A.gas.R),
                                      re(A.gas.gamma), A.signal_factor, i, T, Rc, stF);
                dl = warp_min_d(dl);
                // order-preserving key of a double (negative local time steps of inverted cells included: the
                // reference's min picks them, finite_volume.cpp:180): flip all bits of negatives, the sign of
                // the others; NaN never wins a `<` in the reference and is skipped here
                if (lane == 0 && dl == dl) {
                    const unsigned long long b = (unsigned long long)__double_as_longlong(dl);
                    atomicMin(A.dt_out, b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull));
                }
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(&done_items[s], 1) == P - 1) {
                    // the tile's last item: the stage goes back to the loader
                    done_items[s] = 0;
                    __threadfence_block();
                    ac::mbar_arrive(&empty[s]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass G'
// ---------------------------------------------------------------------------------------------------------
// the tile's slice of t_wd in shared memory: W(j, k) / D(j, k) of tile-local cell idx
template <int NF, int T>
struct TileTables {
    const double* w;
    int idx;
    __device__ __forceinline__ double W(int j, int k) const { return w[(j * 3 + k) * T + idx]; }
    __device__ __forceinline__ double D(int j, int k) const { return w[(NF * 3 + j * 3 + k) * T + idx]; }
};

// input stage = [t_wd slice 2 x NF x 3 x T double | t_grec record grec_stride | state values 5 x (T + halo_cap) S]
template <typename S>
__host__ __device__ inline size_t grad_stage_bytes(int nf, int T, int grec_stride, int halo_cap) {
    size_t b = (size_t)2 * nf * 3 * T * 8 + (size_t)grec_stride + (size_t)5 * (T + halo_cap) * sizeof(S);
    return (b + 127) / 128 * 128;
}
template <typename S>
__host__ __device__ inline size_t grad_out_bytes(int stage_cap) {
    return ((size_t)stage_cap * 6 * sizeof(S) + 127) / 128 * 128;
}

// One work item of pass G': 32 (cell, variable) pairs of a tile.
//   items [0, T / 16):  16 cells x (rho, second thermodynamic variable) -- lanes 0-15 extrapolate rho, lanes 16-31
//                       the other variable, of the same 16 cells; the two lanes of a cell then close the
//                       thermodynamic state (gas_model.h:91-117) of half of the cell's faces each;
//   the others:         32 cells x one velocity component (vx, vy[, vz]).
// One code path computes the gradient, the limiter and the extrapolation of the lane's variable (var_to_faces);
// only the epilogue differs.
template <int DIM, int NF, int T, typename S>
__device__ __forceinline__ void grad_task(const FvArgs<S>& A, const double* wd, const GRec& R, const S* sv, const int svs,
                                          S* stg, const int task, const int nc, const int lane) {
    constexpr int N_THERMO = T / 16;
    const bool thermo = task < N_THERMO;
    const int q = task - N_THERMO;
    const int var = thermo ? (lane >> 4) : 2 + q / (T / 32);
    const int i = thermo ? task * 16 + (lane & 15) : (q % (T / 32)) * 32 + lane;
    const bool active = i < nc;
    const bool limit = A.limiter_on != 0;
    TileTables<NF, T> Tb;
    Tb.w = wd;
    Tb.idx = i;
    int fl[NF], st[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) {
        fl[j] = R.nbr[j * T + i];
        st[j] = R.stage[j * T + i];
    }
    const S ui = sv[var * svs + i];
    S un[NF], r[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) un[j] = sv[var * svs + (fl[j] & NBR_LOCAL_MASK)];
    var_to_faces<DIM, NF>(ui, un, Tb, A.eps, limit, r);
    const int nR = R.hdr[1], n_own_l = R.hdr[2];
    if (thermo) {
        S o[NF];
#pragma unroll
        for (int j = 0; j < NF; j++) o[j] = shfl_xor_s(r[j], 16);
        constexpr int H = (NF + 1) / 2;
#pragma unroll
        for (int jj = 0; jj < H; jj++) {
            const int j0 = jj, j1 = H + jj;
            S rho_j, th_j;
            int st_j;
            if (j1 < NF) {
                rho_j = var ? o[j1] : r[j0];
                th_j = var ? r[j1] : o[j0];
                st_j = var ? st[j1] : st[j0];
            } else {
                rho_j = r[j0];
                th_j = o[j0];
                st_j = var ? 0xFFFF : st[j0];
            }
            if (active && st_j != 0xFFFF) side_close_store(A.thermo, A.gas, rho_j, th_j, stg + (long)st_j * 6);
        }
        // the far side of a boundary interface is the ghost cell itself: zero gradient, unit limiter
        // (convective_flux.cpp:281-300), i.e. its values, closed thermodynamically like any other side
#pragma unroll
        for (int j = 0; j < NF; j++) {
            const bool gh = active && st[j] != 0xFFFF && (fl[j] & NBR_GHOST_BIT);
            if (__any_sync(0xffffffffu, gh)) {
                const S uo = shfl_xor_s(un[j], 16);
                if (gh && var == 0) {
                    const int far_idx = far_stage_index((fl[j] & NBR_LEFT_BIT) != 0, st[j], nR, n_own_l);
                    side_close_store(A.thermo, A.gas, un[j], uo, stg + (long)far_idx * 6);
                }
            }
        }
    } else if (active) {
        const int kq = 3 + (var - 2);
#pragma unroll
        for (int j = 0; j < NF; j++) {
            if (st[j] == 0xFFFF) continue;
            S* qd = stg + (long)st[j] * 6;
            qd[kq] = r[j];
            // 2-D: vz has no gradient (convective_flux.cpp:206-208) and a unit limiter: the vy lane copies the value
            if (DIM == 2 && var == 3) qd[5] = sv[4 * svs + i];
            if (fl[j] & NBR_GHOST_BIT) {
                S* f = stg + (long)far_stage_index((fl[j] & NBR_LEFT_BIT) != 0, st[j], nR, n_own_l) * 6;
                f[kq] = un[j];
                if (DIM == 2 && var == 3) f[5] = sv[4 * svs + (fl[j] & NBR_LOCAL_MASK)];
            }
        }
    }
}

// store transfer q of a tile from its descriptor ti (a copy in shared memory): q = 0 home left range, 1 home right
// range, 2 + e entry e (fv_bodies.h::tile_store_xfer reads the descriptor from global memory)
__device__ __forceinline__ TileXfer tile_store_xfer_ti(const GridDev& g, const int* ti, int q, int n_own_l) {
    TileXfer x;
    if (q == 0) { x.f0 = ti[2] + ti[4]; x.count = ti[3] - ti[4]; x.stage = 0; x.slot = ti[4]; x.left = true; }
    else if (q == 1) { x.f0 = ti[2]; x.count = ti[4] + ti[5]; x.stage = n_own_l; x.slot = 0; x.left = false; }
    else {
        const int* te = tile_entry(g, ti, q - 2);
        x.f0 = te[0]; x.count = te[1]; x.stage = te[4]; x.slot = te[3]; x.left = te[2] == 0;
    }
    return x;
}

// Work items of a block, in queue order: the N_TASKS (cell, variable) items of its m-th tile, then the STORE item of
// tile m - 1 (the tile's side records leave the staging buffer as a handful of bulk stores; queued behind the next
// tile's items so that the records are normally complete when a warp gets there).  Input stages recycle through
// mbarriers guarded by stage_tile (see k_flux_pipe); the staging buffers through the monotone counter out_done.
template <int DIM, int NF, int T, typename S>
__global__ void __launch_bounds__(32 * (IBIS_PIPE_G_WARPS + 1), 1)
    k_grad_pipe(const FvArgs<S> A, const int* __restrict__ tile_list, const int n_list, const int nsi, const int nso) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t rec_full[PIPE_MAX_STAGES], full[PIPE_MAX_STAGES], empty[PIPE_MAX_STAGES];
    __shared__ int stage_tile[PIPE_MAX_STAGES], done_tasks[PIPE_MAX_STAGES], out_done[PIPE_MAX_STAGES],
        meta[PIPE_MAX_STAGES][4], q_head;
    const GridDev& g = A.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - 1;
    const int svs = T + g.halo_cap;
    constexpr int WD_BYTES = 2 * NF * 3 * T * 8;
    const size_t in_bytes = grad_stage_bytes<S>(NF, T, g.grec_stride, g.halo_cap);
    const size_t out_bytes = grad_out_bytes<S>(g.stage_cap);
    unsigned char* out_base = smem + (size_t)nsi * in_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nsi; s++) {
            ac::mbar_init(&rec_full[s], 1);
            ac::mbar_init(&full[s], 33);          // the loader's arrive.expect_tx + one cp.async arrival per lane
            ac::mbar_init(&empty[s], 1);
            stage_tile[s] = -1;
            done_tasks[s] = 0;
        }
        for (int o = 0; o < nso; o++) out_done[o] = -1;
        q_head = 0;
        ac::fence_barrier_init();
    }
    __syncthreads();
    const int n_my = (int)blockIdx.x < n_list ? (n_list - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == ncw) {
        // ---- loader: records IBIS_PIPE_LOOKAHEAD tiles ahead of the state gather ----
        const int la = IBIS_PIPE_LOOKAHEAD < nsi ? IBIS_PIPE_LOOKAHEAD : nsi - 1;
        for (int k = 0; k < n_my + la; k++) {
            if (k < n_my) {
                const int s = k % nsi;
                const int idx = (int)blockIdx.x + k * (int)gridDim.x;
                const int t = tile_list ? __ldg(tile_list + idx) : idx;
                if (k >= nsi && !pipe_wait(&empty[s], ((k / nsi) - 1) & 1, A.pipe_dbg, 16 + PW_EMPTY, k)) return;
                unsigned char* st = smem + (size_t)s * in_bytes;
                if (lane == 0) {
                    int c0, nc;
                    tile_cells(g, t, c0, nc);
                    meta[s][0] = t;
                    meta[s][1] = nc;
                    meta[s][2] = c0;
                    __threadfence_block();
                    *reinterpret_cast<volatile int*>(&stage_tile[s]) = k;
                    ac::mbar_arrive_expect_tx(&rec_full[s], (uint32_t)WD_BYTES + (uint32_t)g.grec_stride);
                    ac::bulk_g2s(st, g.t_wd + (size_t)t * (WD_BYTES / 8), (uint32_t)WD_BYTES, &rec_full[s]);
                    ac::bulk_g2s(st + WD_BYTES, g.t_grec + (size_t)t * g.grec_stride, (uint32_t)g.grec_stride, &rec_full[s]);
                }
                __syncwarp();
            }
            if (k >= la) {
                const int kk = k - la, s = kk % nsi;
                if (!pipe_wait(&rec_full[s], (kk / nsi) & 1, A.pipe_dbg, 16 + PW_REC, kk)) return;
                unsigned char* st = smem + (size_t)s * in_bytes;
                const GRec R = grec_view(st + WD_BYTES, NF, T);
                S* sv = reinterpret_cast<S*>(st + WD_BYTES + g.grec_stride);
                const int c0 = meta[s][2];
                const int n_halo = R.hdr[0];
                // own cells: contiguous rows (bulk copies when 16-byte aligned), halo list: gathered
                const bool bulk_own = ((size_t)c0 * sizeof(S)) % 16 == 0;
                if (lane == 0) ac::mbar_arrive_expect_tx(&full[s], bulk_own ? 5u * (uint32_t)T * (uint32_t)sizeof(S) : 0u);
                __syncwarp();
                if (bulk_own) {
                    if (lane < 5) ac::bulk_g2s(sv + (size_t)lane * svs, A.rc.val[lane] + c0, (uint32_t)T * (uint32_t)sizeof(S), &full[s]);
                } else {
                    for (int q = lane; q < 5 * T; q += 32) {
                        const int v = q / T, i = q % T;
                        ac::cp_async_small<(int)sizeof(S)>(sv + (size_t)v * svs + i, A.rc.val[v] + c0 + i);
                    }
                }
                for (int h = lane; h < n_halo; h += 32) {
                    const int id = R.halo[h];
#pragma unroll
                    for (int v = 0; v < 5; v++) ac::cp_async_small<(int)sizeof(S)>(sv + (size_t)v * svs + T + h, A.rc.val[v] + id);
                }
                ac::cp_async_mbar_arrive_noinc(&full[s]);
            }
        }
        return;
    }
    // ---- consumers ----
    constexpr int N_THERMO = T / 16, NV = DIM == 3 ? 3 : 2, N_TASKS = N_THERMO + NV * (T / 32), P = N_TASKS + 1;
    const int total = (n_my + 1) * P;
    for (;;) {
        const int item = pipe_next_item(&q_head, lane);
        if (item >= total) break;
        const int m = item / P, r = item % P;
        const bool is_task = r < N_TASKS;
        const int k = is_task ? m : m - 1;
        if (k < 0 || k >= n_my) continue;
        const int s = k % nsi, o = k % nso;
        if (!pipe_spin_ge(&stage_tile[s], k, A.pipe_dbg, 16 + PW_STAGE, k)) return;
        if (!pipe_wait(&rec_full[s], (k / nsi) & 1, A.pipe_dbg, 16 + PW_REC, k)) return;
        unsigned char* st = smem + (size_t)s * in_bytes;
        const GRec R = grec_view(st + WD_BYTES, NF, T);
        S* stg = reinterpret_cast<S*>(out_base + (size_t)o * out_bytes);
        if (is_task) {
            if (!pipe_wait(&full[s], (k / nsi) & 1, A.pipe_dbg, 16 + PW_FULL, k)) return;
            // the staging buffer is free once the stores of the tile nso tiles back have read it
            if (k >= nso && !pipe_spin_ge(&out_done[o], k - nso, A.pipe_dbg, 16 + PW_OUT, k)) return;
            const double* wd = reinterpret_cast<const double*>(st);
            const S* sv = reinterpret_cast<const S*>(st + WD_BYTES + g.grec_stride);
            grad_task<DIM, NF, T, S>(A, wd, R, sv, svs, stg, r, meta[s][1], lane);
            ac::fence_proxy_async();                       // the staging writes -> visible to the TMA unit
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&done_tasks[s], 1);
            }
        } else {
            // the tile's side records are complete in the staging buffer: they leave as a handful of bulk stores
            if (!pipe_spin_ge(&done_tasks[s], N_TASKS, A.pipe_dbg, 16 + PW_DONE, k)) return;
            ac::fence_proxy_async();
            const int n_ent = R.info[7], n_own_l = R.hdr[2];
            for (int q = lane; q < 2 + n_ent; q += 32) {
                const TileXfer x = tile_store_xfer_ti(g, R.info, q, n_own_l);
                if (x.count > 0) {
                    S* dst = (x.left ? A.sideL : A.sideR) + (long)x.f0 * 6;
                    ac::bulk_s2g(dst, stg + (long)x.stage * 6, (uint32_t)x.count * 6 * (uint32_t)sizeof(S));
                }
            }
            __syncwarp();
            if (lane == 0) {                               // the input stage goes back to the loader
                done_tasks[s] = 0;
                __threadfence_block();
                ac::mbar_arrive(&empty[s]);
            }
            ac::bulk_commit_wait_read();                   // the staging buffer may be written again after this
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                *reinterpret_cast<volatile int*>(&out_done[o]) = k;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
template <typename K>
static bool set_smem(ibis_b200_ctx* c, K kernel, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        if (c->last_error.empty()) c->last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        return false;
    }
    return true;
}

template <int DIM, int NF, int T, typename S>
static void launch_grad_pipe_t(ibis_b200_ctx* c, const FvArgs<S>& A, const int* list, int n_list) {
    const size_t in_b = grad_stage_bytes<S>(NF, T, A.g.grec_stride, A.g.halo_cap), out_b = grad_out_bytes<S>(A.g.stage_cap);
    // stages: as many as fit, inputs : outputs about 3 : 2
    int nsi = 2, nso = 1;
    for (;;) {
        const bool grow_in = nsi < PIPE_MAX_STAGES && 2 * nsi <= 3 * nso + 2;
        const size_t next = (size_t)(nsi + (grow_in ? 1 : 0)) * in_b + (size_t)(nso + (grow_in ? 0 : 1)) * out_b;
        if (next > PIPE_SMEM_BUDGET || (!grow_in && nso >= PIPE_MAX_STAGES)) break;
        if (grow_in) nsi++; else nso++;
    }
    const size_t smem = (size_t)nsi * in_b + (size_t)nso * out_b;
    if (smem > PIPE_SMEM_BUDGET) {
        if (c->last_error.empty()) c->last_error = "pass G': a tile does not fit shared memory";
        return;
    }
    auto kernel = k_grad_pipe<DIM, NF, T, S>;
    if (!set_smem(c, kernel, smem)) return;
    const int grid = n_list < c->n_sm ? n_list : c->n_sm;
    kernel<<<grid, 32 * (IBIS_PIPE_G_WARPS + 1), smem, c->stream>>>(A, list, n_list, nsi, nso);
}

template <int DIM, int T, typename S>
static void launch_grad_pipe_nf(ibis_b200_ctx* c, const FvArgs<S>& A, const int* list, int n_list) {
    switch (A.g.pipe_nf) {
        case 6: launch_grad_pipe_t<DIM, 6, T, S>(c, A, list, n_list); break;
        case 5: launch_grad_pipe_t<DIM, 5, T, S>(c, A, list, n_list); break;
        case 4: launch_grad_pipe_t<DIM, 4, T, S>(c, A, list, n_list); break;
        default: launch_grad_pipe_t<DIM, 3, T, S>(c, A, list, n_list); break;
    }
}

// which: 0 = every tile, 1 = the early tiles (no halo cell, no ghost among their neighbours), 2 = the late ones
template <typename S>
void launch_grad_sides(ibis_b200_ctx* c, const FvArgs<S>& A, int which) {
    const int* list = which == 1 ? A.g.tiles_early : which == 2 ? A.g.tiles_late : nullptr;
    const int n_list = which == 1 ? A.g.n_early : which == 2 ? A.g.n_late : A.g.n_tiles_all;
    if (n_list == 0) return;
    prof_begin(c, 0);
    if (A.g.dim == 3) {
        if (A.g.tile == 64) launch_grad_pipe_nf<3, 64, S>(c, A, list, n_list);
        else launch_grad_pipe_nf<3, 128, S>(c, A, list, n_list);
    } else {
        if (A.g.tile == 64) launch_grad_pipe_nf<2, 64, S>(c, A, list, n_list);
        else launch_grad_pipe_nf<2, 128, S>(c, A, list, n_list);
    }
    prof_end(c, 0);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_grad_sides<double>(ibis_b200_ctx*, const FvArgs<double>&, int);
template void launch_grad_sides<dual>(ibis_b200_ctx*, const FvArgs<dual>&, int);

template <int DIM, int FLUX, typename S>
static void launch_flux_pipe_k(ibis_b200_ctx* c, const FvArgs<S>& A) {
    const size_t stage = flux_stage_bytes<S>(A.g.tile_cap, A.g.frec_stride, A.g.tile, A.dt_out != nullptr);
    int nst = (int)(PIPE_SMEM_BUDGET / stage);
    if (nst > PIPE_MAX_STAGES) nst = PIPE_MAX_STAGES;
    if (nst < 2) {
        if (c->last_error.empty()) c->last_error = "pass F': a tile does not fit shared memory";
        return;
    }
    auto kernel = k_flux_pipe<DIM, FLUX, S>;
    if (!set_smem(c, kernel, (size_t)nst * stage)) return;
    const int grid = A.g.n_tiles < c->n_sm ? A.g.n_tiles : c->n_sm;
    kernel<<<grid, 32 * (IBIS_PIPE_F_WARPS + 1), (size_t)nst * stage, c->stream>>>(A, nst);
}
template <int DIM, typename S>
static void launch_flux_pipe_dim(ibis_b200_ctx* c, const FvArgs<S>& A) {
    switch (c->cfg.flux_calculator) {
        case FLUX_HANEL: launch_flux_pipe_k<DIM, FLUX_HANEL, S>(c, A); break;
        case FLUX_AUSMDV: launch_flux_pipe_k<DIM, FLUX_AUSMDV, S>(c, A); break;
        case FLUX_LDFSS2: launch_flux_pipe_k<DIM, FLUX_LDFSS2, S>(c, A); break;
        default: launch_flux_pipe_k<DIM, FLUX_RUSANOV, S>(c, A); break;
    }
}

template <typename S>
void launch_flux_tile(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_tiles == 0) return;
    prof_begin(c, 1);
    if (A.g.dim == 3) launch_flux_pipe_dim<3, S>(c, A);
    else launch_flux_pipe_dim<2, S>(c, A);
    prof_end(c, 1);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_flux_tile<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_flux_tile<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

// the fused dt reduction of pass F' works on order-preserving integer keys (see k_flux_pipe)
__global__ void k_dt_key_init(unsigned long long* p) { *p = 0x7ff0000000000000ull ^ 0x8000000000000000ull; }   // +inf
__global__ void k_dt_key_decode(unsigned long long* p) {
    const unsigned long long k = *p;
    *p = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
}
void launch_dt_reduce_init(ibis_b200_ctx* c, double* p) {
    k_dt_key_init<<<1, 1, 0, c->stream>>>(reinterpret_cast<unsigned long long*>(p));
    CUDA_CHECK_LAUNCH(c);
}
void launch_dt_reduce_finish(ibis_b200_ctx* c, double* p) {
    k_dt_key_decode<<<1, 1, 0, c->stream>>>(reinterpret_cast<unsigned long long*>(p));
    CUDA_CHECK_LAUNCH(c);
}

}  // namespace ib
