// fv_pipe.cu -- the sides path of the inviscid second-order residual as two PERSISTENT, warp-specialised
// sm_100a kernels (one block per SM):
//
//   k_grad_pipe  pass G': WLS gradients (grid/gradient.h:66-116), Barth-Jespersen limiters (limiter.cpp:24-63)
//                and the MUSCL extrapolation + thermodynamic closure of linear_reconstruct
//                (convective_flux.cpp:243-420, gas_model.h:91-117) -> one 48-byte side record per interface side
//   k_flux_pipe  pass F': rotate, flux calculator, rotate back (convective_flux.cpp:148-164) and
//                flux_surface_integral (finite_volume.cpp:221-256), owner computes, no atomics on the residual
//
// Both have the same anatomy.  ONE loader warp walks the block's tiles and keeps a ring of shared-memory stages
// full.  Everything a tile needs is a packed per-tile record or a per-tile array (GridDev::t_wd, t_grec, t_freca,
// t_frec, the side-state arrays), so a stage is filled by two to seven TMA bulk copies (cp.async.bulk, SASS UBLKCP)
// completing on an mbarrier -- the loader issues its bulk copies one lane at a time, so their NUMBER per tile is what
// bounds it; the only gathered data (the state of the neighbour cells outside the tile, pass G') is fetched by the
// loader's lanes with cp.async into the same stage and reported to the same mbarrier.  The loader also asks the TMA
// unit to pull the data of the tile a few iterations ahead into L2 (cp.async.bulk.prefetch.L2), so that a stage waits
// for an L2 round trip, not a DRAM one.  The CONSUMER warps never touch global memory on the input side: they take
// work items (32 interfaces, or 32 (cell, variable) pairs) off the block's ordered queue, wait for the item's stage,
// and the warp that completes a tile's last item hands the stage back (and, pass G', sends the tile's side records
// off: the staging image into the tile's own side-state array, the runs a neighbouring tile needs into that
// tile's array -- pass F' then reads ONE contiguous array per tile).  No block-wide barrier after start-up; warps of
// one block are at different tiles at the same time, which is what hides the latency of the fp64 dependency chains.
//
// Arithmetic: the same per-value expression trees as the reference (and as the host emulation in tests/emu runs
// them through fv_bodies.h::pipe_grad_cell / flux_from_sides6 / pipe_cell_sum), hence the same bits.
#include <cuda_runtime.h>

#include "async_copy.h"
#include "context.h"

#ifndef IBIS_PIPE_F_WARPS
#define IBIS_PIPE_F_WARPS 14      // consumer warps of pass F'
#endif
#ifndef IBIS_PIPE_F_LOADERS
#define IBIS_PIPE_F_LOADERS 2     // loader warps of pass F' (loader w takes the block's tiles w, w + NL, ...)
#endif
#ifndef IBIS_PIPE_G_WARPS
#define IBIS_PIPE_G_WARPS 13      // consumer warps of pass G'
#endif
#ifndef IBIS_PIPE_G_LOADERS
#define IBIS_PIPE_G_LOADERS 2     // loader warps of pass G' (one more service warp sends the side records off)
#endif
#ifndef IBIS_PIPE_F_ROUNDS
#define IBIS_PIPE_F_ROUNDS 1      // interfaces per lane of a flux item of pass F' (measured: 2 is slower)
#endif
#ifndef IBIS_PIPE_G_SPLITVEL
#define IBIS_PIPE_G_SPLITVEL 0    // 1: a velocity item of pass G' covers 32 cells (one pass) instead of the whole tile
#endif
#ifndef IBIS_PIPE_PREFETCH
#define IBIS_PIPE_PREFETCH 6      // a loader pulls a tile's data into L2 this many tiles (a multiple of its stride) ahead
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
constexpr int PIPE_HALO_REGS = 8;                        // halo ids a loader lane keeps in registers (halo list <= 256)
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
// Ordering between the warps of a block (flag published after data / data read after flag): an acquire-release fence
// at block scope.  __threadfence_block() is the sequentially consistent one (MEMBAR.SC.CTA), which every item paid
// for several times on its critical path.
__device__ __forceinline__ void pipe_fence() {
#ifdef IBIS_PIPE_SC_FENCE
    __threadfence_block();
#else
    asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
}
// Waits poll with a growing sleep instead of suspending on the barrier: a stage barrier sees some forty events per
// phase (one per bulk copy), and an event-driven sleep wakes every waiting warp at each of them -- measured, that
// cost more issue slots than the work itself.
__device__ __forceinline__ unsigned pipe_backoff(unsigned ns) {
    __nanosleep(ns);
    return ns < 512u ? ns * 2u : ns;
}
// the phase of an mbarrier (the polling loops are out of line: the instruction cache should see the work, not the waits)
__device__ __noinline__ bool pipe_wait_slow(uint64_t* bar, uint32_t parity, unsigned int* dbg, int kind, int k) {
    const long long t0 = clock64();
    unsigned ns = 64u;
    for (int it = 1;; it++) {
        ns = pipe_backoff(ns);
        if (ac::mbar_test_wait(bar, parity)) return true;
        if ((it & 63) == 0 && pipe_timed_out(t0, dbg, kind, k, parity)) return false;
    }
}
__device__ __forceinline__ bool pipe_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, int kind, int k) {
    bool ok = ac::mbar_test_wait(bar, parity);
    if (!ok) ok = pipe_wait_slow(bar, parity, dbg, kind, k);
    return __all_sync(0xffffffffu, ok);
}
// a monotone counter in shared memory reaching `target` (no parity, hence no aliasing when a warp is late)
__device__ __noinline__ bool pipe_spin_slow(const int* p, int target, unsigned int* dbg, int kind, int k) {
    const volatile int* vp = p;
    const long long t0 = clock64();
    unsigned ns = 64u;
    for (int it = 1;; it++) {
        ns = pipe_backoff(ns);
        if (*vp >= target) return true;
        if ((it & 63) == 0 && pipe_timed_out(t0, dbg, kind, k, (unsigned)target)) return false;
    }
}
__device__ __forceinline__ bool pipe_spin_ge(const int* p, int target, unsigned int* dbg, int kind, int k) {
    bool ok = *reinterpret_cast<const volatile int*>(p) >= target;
    if (!ok) ok = pipe_spin_slow(p, target, dbg, kind, k);
    ok = __all_sync(0xffffffffu, ok);
    pipe_fence();
    return ok;
}
// the same for a single thread (the loaders of pass F' run on one lane)
__device__ __forceinline__ bool pipe_spin_ge_1(const int* p, int target, unsigned int* dbg, int kind, int k) {
    bool ok = *reinterpret_cast<const volatile int*>(p) >= target;
    if (!ok) ok = pipe_spin_slow(p, target, dbg, kind, k);
    pipe_fence();
    return ok;
}
// next item of the block's ordered work queue
__device__ __forceinline__ int pipe_next_item(int* head, int lane) {
    int i = 0;
    if (lane == 0) i = atomicAdd(head, 1);
    return __shfl_sync(0xffffffffu, i, 0);
}

// n / d for the small n of the work queue, d given as m = ceil(2^32 / d) (exact while n * d < 2^32)
__host__ __device__ inline unsigned pipe_magic(int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }
__device__ __forceinline__ int pipe_div(int n, unsigned m) { return (int)__umulhi((unsigned)n, m); }

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
// setup: tile-major copy of the static WLS tables
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
void launch_pack_pipe_tables(ibis_b200_ctx* c, double* t_wd) {
    const GridDev& g = c->g;
    if (g.n_grad > 0) k_pack_wd<<<(g.n_grad + 255) / 256, 256, 0, c->stream>>>(g, t_wd);
    CUDA_CHECK_LAUNCH(c);
}

// ---------------------------------------------------------------------------------------------------------
// pass F'
// ---------------------------------------------------------------------------------------------------------
// Two rings in shared memory:
//   input stage  [the tile's side-state array p_cap x 6 S | its interface record freca_stride]: the operands of the
//                tile's interfaces.  A flux item moves its 32 interfaces' operands into registers first, and the
//                stage goes back to the loader as soon as every item of the tile has done so -- it is occupied for
//                a load and a few hundred cycles, not for the flux computation;
//   flux buffer  [fluxes cap x 5 S | cell record frec_stride | (dt fused) vx vy vz T of the tile's cells, 4 x T S]:
//                lives until the tile's cells have summed their faces.
// Shared memory holds what is IN FLIGHT from L2 plus the small per-tile results, instead of every operand for the
// whole time a tile is being worked on.
template <typename S>
__host__ __device__ inline size_t flux_in_bytes(int p_cap, int freca_stride) {
    return ((size_t)p_cap * SIDE_REC * sizeof(S) + 127) / 128 * 128 + (size_t)freca_stride;
}
template <typename S>
__host__ __device__ inline size_t flux_fb_bytes(int cap, int frec_stride, int T, bool with_dt) {
    size_t b = (size_t)cap * 5 * sizeof(S) + (size_t)frec_stride + (with_dt ? (size_t)4 * T * sizeof(S) : 0);
    return (b + 127) / 128 * 128;
}

// Work items of a block, handed out in this order by one counter: for its m-th tile the NCH flux chunks
// (32 interfaces each, one per lane), then the T / 32 cell parts of tile m - 1 (one lane per cell sums its faces; queued
// behind the NEXT tile's chunks so that the fluxes they need are normally complete when a warp gets there).  A warp
// holding an item of tile k first waits for stage_tile[s] >= k / fb_tile[o] >= k (the loader has re-designated the
// stage / buffer) and only then tests the barrier's parity: neither can be recycled while an item of its tile is
// outstanding, so the barrier is then exactly at the awaited phase or one behind, never two.
template <int DIM, int FLUX, typename S>
__global__ void __launch_bounds__(32 * (IBIS_PIPE_F_WARPS + IBIS_PIPE_F_LOADERS), 1) k_flux_pipe(const FvArgs<S> A, const int nsi, const int nfb, const int pf) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_in[PIPE_MAX_STAGES], rec_full[PIPE_MAX_STAGES];
    __shared__ int stage_tile[PIPE_MAX_STAGES], stage_free[PIPE_MAX_STAGES], reads_done[PIPE_MAX_STAGES], n_slots_of[PIPE_MAX_STAGES];
    __shared__ int fb_tile[PIPE_MAX_STAGES], fb_done[PIPE_MAX_STAGES], done_items[PIPE_MAX_STAGES], meta[PIPE_MAX_STAGES][4];
    __shared__ int q_head;
    const GridDev& g = A.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - IBIS_PIPE_F_LOADERS;
    const int cap = g.tile_cap, T = g.tile;
    const bool with_dt = A.dt_out != nullptr;
    const unsigned in_bytes = (unsigned)flux_in_bytes<S>(g.p_cap, g.freca_stride), fb_bytes = (unsigned)flux_fb_bytes<S>(cap, g.frec_stride, T, with_dt);
    const unsigned offA = in_bytes - (unsigned)g.freca_stride;             // interface record inside the input stage
    const unsigned offRec = (unsigned)cap * 5u * (unsigned)sizeof(S), offDt = offRec + (unsigned)g.frec_stride;
    unsigned char* fb_base = smem + (unsigned)nsi * in_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nsi; s++) {
            ac::mbar_init(&full_in[s], 1);
            stage_tile[s] = -1;
            stage_free[s] = -1;
            reads_done[s] = 0;
        }
        for (int o = 0; o < nfb; o++) {
            ac::mbar_init(&rec_full[o], 1);
            fb_tile[o] = -1;
            fb_done[o] = -1;
            done_items[o] = 0;
        }
        q_head = 0;
        ac::fence_barrier_init();
    }
    __syncthreads();
    const int n_my = (int)blockIdx.x < g.n_tiles ? (g.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp >= ncw) {
        // ---- loaders: loader w takes the block's tiles w, w + NL, ...  Per tile: the tile's own side-state array,
        // its interface record and its cell record (+ 4 rows with the fused dt) as bulk copies issued by lane 0, and
        // the runs of the neighbouring tiles' arrays that complete the image of the tile's interface sides, one
        // per lane from lane 1 on (GridDev::t_pull); the same ranges of a later tile of its own as L2 prefetches ----
        const int step = (int)gridDim.x, w = warp - ncw;
        constexpr int NL = IBIS_PIPE_F_LOADERS;
        const int PF = (pf + NL - 1) / NL * NL;
        // header {home, n_home, n_slots, n_pull, n_image} of the tile and this lane's run, fetched one tile ahead
        int h0 = 0, h1 = 0, h2 = 0, h3 = 0, h4 = 0, ps = 0, pc = 0, pd = 0;
        auto fetch_pull = [&](size_t t, int& a0, int& a1, int& a2, int& a3, int& a4, int& b0, int& b1, int& b2) {
            const int* r = g.t_pull + t * (size_t)g.pull_stride;
            a0 = __ldg(r); a1 = __ldg(r + 1); a2 = __ldg(r + 2); a3 = __ldg(r + 3); a4 = __ldg(r + 4);
            b0 = b1 = b2 = 0;
            if (lane >= 1 && lane <= g.pull_cap) {
                b0 = __ldg(r + 8 + 3 * (lane - 1)); b1 = __ldg(r + 9 + 3 * (lane - 1)); b2 = __ldg(r + 10 + 3 * (lane - 1));
            }
        };
        if (w < n_my) fetch_pull((size_t)blockIdx.x + (size_t)w * step, h0, h1, h2, h3, h4, ps, pc, pd);
        // the prefetch tile's header and this lane's run: read one iteration before the prefetches are issued
        int f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, fps = 0, fpc = 0, fpd = 0;
        bool have_pf = false;
        for (int k = w; k < n_my; k += NL) {
            const int t = (int)blockIdx.x + k * step, s = k % nsi, o = k % nfb;
            if (have_pf) {                                    // L2 prefetch of tile k - NL + PF (list read an iteration ago)
                const size_t tp = (size_t)t + (size_t)(PF - NL) * step;
                constexpr uint32_t RBp = (uint32_t)SIDE_REC * (uint32_t)sizeof(S);
                if (lane == 0) {
                    ac::bulk_prefetch_l2(A.sidesP + (size_t)f0 * SIDE_REC, (uint32_t)f1 * RBp);
                    ac::bulk_prefetch_l2(g.t_freca + tp * g.freca_stride, (uint32_t)g.freca_stride);
                    ac::bulk_prefetch_l2(g.t_frec + tp * g.frec_stride, (uint32_t)g.frec_stride);
                    if (with_dt) {
                        int c0p, ncp;
                        tile_cells(g, (int)tp, c0p, ncp);
                        for (int v = 0; v < 4; v++) ac::bulk_prefetch_l2(A.fs[v < 3 ? F_VX + v : F_T] + c0p, (uint32_t)T * (uint32_t)sizeof(S));
                    }
                } else if (lane <= f3) {
                    ac::bulk_prefetch_l2(A.sidesP + (size_t)fps * SIDE_REC, (uint32_t)fpc * RBp);
                }
            }
            have_pf = PF > 0 && k + PF < n_my;
            if (have_pf) fetch_pull((size_t)t + (size_t)PF * step, f0, f1, f2, f3, f4, fps, fpc, fpd);
            int n0 = 0, n1 = 0, n2 = 0, n3 = 0, n4 = 0, ns_ = 0, nc_ = 0, nd_ = 0;   // this loader's next tile, in flight during the waits
            if (k + NL < n_my) fetch_pull((size_t)t + (size_t)NL * step, n0, n1, n2, n3, n4, ns_, nc_, nd_);
            if (k >= nsi && !pipe_spin_ge(&stage_free[s], k - nsi, A.pipe_dbg, PW_EMPTY, k)) return;
            if (k >= nfb && !pipe_spin_ge(&fb_done[o], k - nfb, A.pipe_dbg, PW_OUT, k)) return;
            int c0, nc;
            tile_cells(g, t, c0, nc);
            unsigned char* st = smem + (unsigned)s * in_bytes;
            unsigned char* fb = fb_base + (unsigned)o * fb_bytes;
            constexpr uint32_t RB = (uint32_t)SIDE_REC * (uint32_t)sizeof(S);
            if (lane == 0) {
                n_slots_of[s] = h2;
                meta[o][0] = nc;
                meta[o][1] = c0;
                meta[o][2] = t;
                pipe_fence();
                *reinterpret_cast<volatile int*>(&fb_tile[o]) = k;
                *reinterpret_cast<volatile int*>(&stage_tile[s]) = k;
                ac::mbar_arrive_expect_tx(&full_in[s], (uint32_t)h4 * RB + (uint32_t)g.freca_stride);
                ac::mbar_arrive_expect_tx(&rec_full[o], (uint32_t)g.frec_stride + (with_dt ? 4u * (uint32_t)T * (uint32_t)sizeof(S) : 0u));
                ac::bulk_g2s(st, A.sidesP + (size_t)h0 * SIDE_REC, (uint32_t)h1 * RB, &full_in[s]);
                ac::bulk_g2s(st + offA, g.t_freca + (size_t)t * g.freca_stride, (uint32_t)g.freca_stride, &full_in[s]);
                ac::bulk_g2s(fb + offRec, g.t_frec + (size_t)t * g.frec_stride, (uint32_t)g.frec_stride, &rec_full[o]);
                if (with_dt)
                    for (int v = 0; v < 4; v++)
                        ac::bulk_g2s(fb + offDt + (unsigned)v * (unsigned)T * (unsigned)sizeof(S), A.fs[v < 3 ? F_VX + v : F_T] + c0,
                                     (uint32_t)T * (uint32_t)sizeof(S), &rec_full[o]);
            }
            __syncwarp();                                     // the runs follow the barrier's expect_tx
            if (lane >= 1 && lane <= h3)
                ac::bulk_g2s(st + (size_t)pd * RB, A.sidesP + (size_t)ps * SIDE_REC, (uint32_t)pc * RB, &full_in[s]);
            h0 = n0; h1 = n1; h2 = n2; h3 = n3; h4 = n4; ps = ns_; pc = nc_; pd = nd_;
        }
        return;
    }
    // ---- consumers ----
    const int NCH = (cap + 32 * IBIS_PIPE_F_ROUNDS - 1) / (32 * IBIS_PIPE_F_ROUNDS), NCL = (T + 31) >> 5, P = NCH + NCL;
    const int total = (n_my + 1) * P;
    const unsigned mP = pipe_magic(P), mI = pipe_magic(nsi), mO = pipe_magic(nfb);
    // (claiming the next item while the current one is worked on hides the latency of the claim, but an item claimed
    // by a busy warp cannot be started by an idle one: measured, that costs far more than it saves)
    for (;;) {
        const int item = pipe_next_item(&q_head, lane);
        if (item >= total) break;
        const int m = pipe_div(item, mP), r = item - m * P;
        const bool is_flux = r < NCH;
        const int k = is_flux ? m : m - 1;
        if (k < 0 || k >= n_my) continue;
        const int ui = pipe_div(k, mI), uo = pipe_div(k, mO);         // use counts of the stage and of the buffer
        const int s = k - ui * nsi, o = k - uo * nfb;
        unsigned char* fb = fb_base + (unsigned)o * fb_bytes;
        S* flux = reinterpret_cast<S*>(fb);
        if (is_flux) {
            if (!pipe_spin_ge(&stage_tile[s], k, A.pipe_dbg, PW_STAGE, k)) return;
            if (!pipe_wait(&full_in[s], ui & 1, A.pipe_dbg, PW_FULL, k)) return;
            const unsigned char* st = smem + (unsigned)s * in_bytes;
            const S* Pt = reinterpret_cast<const S*>(st);
            const FRecA Ra = freca_view(st + offA, cap);
            const int n_slots = n_slots_of[s];
            // IBIS_PIPE_F_ROUNDS rounds of 32 interfaces; the operands of the last round are in registers before the
            // stage is declared read
#pragma unroll 1
            for (int rr = 0; rr < IBIS_PIPE_F_ROUNDS; rr++) {
                const int slot = r * (32 * IBIS_PIPE_F_ROUNDS) + rr * 32 + lane;
                const bool active = slot < n_slots;
                FluxOperands<S> O;
                if (active) flux_load_operands(A, Pt + (long)Ra.lidx[slot] * SIDE_REC, Pt + (long)Ra.ridx[slot] * SIDE_REC, Ra.frame + (long)slot * 6, O);
                if (rr == IBIS_PIPE_F_ROUNDS - 1) {
                    // once every item of the tile has its operands the stage is the loader's again
                    __syncwarp();
                    if (lane == 0) {
                        pipe_fence();
                        if (atomicAdd(&reads_done[s], 1) == NCH - 1) {
                            reads_done[s] = 0;
                            pipe_fence();
                            *reinterpret_cast<volatile int*>(&stage_free[s]) = k;
                        }
                    }
                }
                if (active) flux_from_operands<DIM, FLUX>(A, O, flux + (long)slot * 5);
            }
            __syncwarp();
            if (lane == 0) {
                pipe_fence();
                atomicAdd(&done_items[o], 1);
            }
        } else {
            if (!pipe_spin_ge(&fb_tile[o], k, A.pipe_dbg, PW_STAGE, k)) return;
            if (!pipe_wait(&rec_full[o], uo & 1, A.pipe_dbg, PW_REC, k)) return;
            // every flux of the tile must be in the buffer
            if (!pipe_spin_ge(&done_items[o], NCH, A.pipe_dbg, PW_DONE, k)) return;
            const int nc = meta[o][0], c0 = meta[o][1], t = meta[o][2];
            const FRec Rc = frec_view(fb + offRec, T);
            const int i = (r - NCH) * 32 + lane;
            if (i < nc) pipe_cell_sum<DIM, 5>(A, c0 + i, i, T, Rc, flux);
            if (with_dt) {                      // estimate_dt fused into this evaluation (uniform branch)
                // the staged frames have left shared memory with the input stage: the tile's interface record in
                // global memory serves (L2: this block fetched it a few microseconds ago)
                const S* dtx = reinterpret_cast<const S*>(fb + offDt);
                const double* frames = freca_view(g.t_freca + (size_t)t * g.freca_stride, cap).frame;
                double dl = INFINITY;
                if (i < nc)
                    dl = pipe_cell_dt(re(dtx[i]), re(dtx[T + i]), re(dtx[2 * T + i]), re(dtx[3 * T + i]), re(A.gas.R),
                                      re(A.gas.gamma), A.signal_factor, i, T, Rc, frames);
                dl = warp_min_d(dl);
                // order-preserving key of a double (negative local time steps of inverted cells included: the
                // reference's min picks them, finite_volume.cpp:180): flip all bits of negatives, the sign of
                // the others; NaN never wins a `<` in the reference and is skipped here
                if (lane == 0 && dl == dl) {
                    const unsigned long long b = (unsigned long long)__double_as_longlong(dl);
                    atomicMin(A.dt_out, b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull));
                }
            }
            ac::fence_proxy_async();            // reads of the cell record precede the next bulk copy into the buffer
            __syncwarp();
            if (lane == 0) {
                pipe_fence();
                if (atomicAdd(&done_items[o], 1) == P - 1) {
                    // the tile's last item: the buffer goes back to the loader
                    done_items[o] = 0;
                    pipe_fence();
                    *reinterpret_cast<volatile int*>(&fb_done[o]) = k;
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
// staging buffer = [side records stage_cap x SIDE_REC S | int home, n_stage: where the image goes (first record of the
// tile's side-state array, record count)]
template <typename S>
__host__ __device__ inline size_t grad_out_bytes(int stage_cap) {
    return ((size_t)stage_cap * SIDE_REC * sizeof(S) + 16 + 127) / 128 * 128;
}

// One work item of pass G', one cell per lane:
//   items [0, T / 32):  32 cells x (rho, then the second thermodynamic variable) -> the first two places of the
//                       cell's side records (fv_bodies.h::side_second; pass F' finishes the closure, gas_model.h:91-117);
//   the others:         all T cells (T / 32 rounds of 32) x one velocity component -> vx | vy | vz of the records.
// Items are of about the same size (two gradient + limiter + extrapolation passes each at T = 64).  ONE loop body
// serves every pass of both kinds -- a single instance of var_to_faces in the kernel: with twenty warps at different
// places of the code, its size is what the instruction cache sees (a variant with four inlined instances spent 37 %
// of its stall samples waiting for instructions).
template <int DIM, int NF, int T, typename S>
__device__ __forceinline__ void grad_item(const FvArgs<S>& A, const double* wd, const GRec& R, const S* sv, const int svs,
                                          S* stg, const int task, const int nc, const int lane) {
    constexpr int N_THERMO = T / 32;
    const bool limit = A.limiter_on != 0;
    const bool thermo = task < N_THERMO;
    const bool ghosts = R.hdr[4] != 0;                    // warp-uniform: the tile has boundary interfaces at all
    const int nR = R.hdr[1], n_own_l = R.hdr[2];
    constexpr int VR = IBIS_PIPE_G_SPLITVEL ? T / 32 : 1;  // velocity items per component
    const int v = (task - N_THERMO) / VR;                 // velocity item: 0, 1, 2 = vx, vy, vz
    const int vround = (task - N_THERMO) % VR;
    const int n_pass = thermo ? 2 : (IBIS_PIPE_G_SPLITVEL ? 1 : T / 32);
#pragma unroll 1
    for (int pass = 0; pass < n_pass; pass++) {
        const int i = (thermo ? task : (IBIS_PIPE_G_SPLITVEL ? vround : pass)) * 32 + lane;
        if (i >= nc) continue;
        const S* row = sv + (size_t)(thermo ? pass : 2 + v) * svs;
        TileTables<NF, T> Tb;
        Tb.w = wd;
        Tb.idx = i;
        int fl[NF], st[NF];
#pragma unroll
        for (int j = 0; j < NF; j++) {
            fl[j] = R.nbr[j * T + i];
            st[j] = R.stage[j * T + i];
        }
        const S ui = row[i];
        S un[NF], r[NF];
#pragma unroll
        for (int j = 0; j < NF; j++) un[j] = row[fl[j] & NBR_LOCAL_MASK];
        var_to_faces<DIM, NF>(ui, un, Tb, A.eps, limit, r);
        // the far side of a boundary interface is the ghost cell itself: zero gradient, unit limiter
        // (convective_flux.cpp:281-300), i.e. its values
        const int slot_v = thermo ? pass : 2 + v;             // place of this variable in the side records
        const bool second = thermo && pass == 1;              // fv_bodies.h::side_second: T = u / Cv for rho_u
#pragma unroll
        for (int j = 0; j < NF; j++) {
            if (st[j] == 0xFFFF) continue;
            stg[(long)st[j] * SIDE_REC + slot_v] = second ? side_second(A.thermo, A.gas, r[j]) : r[j];
            if (ghosts && (fl[j] & NBR_GHOST_BIT))
                stg[(long)far_stage_index((fl[j] & NBR_LEFT_BIT) != 0, st[j], nR, n_own_l) * SIDE_REC + slot_v] =
                    second ? side_second(A.thermo, A.gas, un[j]) : un[j];
        }
        if (DIM == 2 && !thermo && v == 1) {
            // 2-D: vz has no gradient (convective_flux.cpp:206-208) and a unit limiter: the value is copied
            const S* rz = sv + (size_t)4 * svs;
#pragma unroll
            for (int j = 0; j < NF; j++) {
                if (st[j] == 0xFFFF) continue;
                stg[(long)st[j] * SIDE_REC + 4] = rz[i];
                if (ghosts && (fl[j] & NBR_GHOST_BIT))
                    stg[(long)far_stage_index((fl[j] & NBR_LEFT_BIT) != 0, st[j], nR, n_own_l) * SIDE_REC + 4] = rz[fl[j] & NBR_LOCAL_MASK];
            }
        }
    }
}

// Work items of a block, in queue order: the N_TASKS items of its tiles (grad_item).  The warp that completes a
// tile's last item hands the input stage back to the loaders and the staging buffer over to the store warp (one bulk
// store of the staging image into the tile's side-state array).  Input stages recycle through an mbarrier guarded by stage_tile
// (see k_flux_pipe), the staging buffers through the monotone counter out_done.
template <int DIM, int NF, int T, typename S>
__global__ void __launch_bounds__(32 * (IBIS_PIPE_G_WARPS + IBIS_PIPE_G_LOADERS + 1), 1)
    k_grad_pipe(const FvArgs<S> A, const int* __restrict__ tile_list, const int n_list, const int nsi, const int nso, const int pf) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[PIPE_MAX_STAGES];
    __shared__ int stage_tile[PIPE_MAX_STAGES], stage_free[PIPE_MAX_STAGES], done_tasks[PIPE_MAX_STAGES], out_ready[PIPE_MAX_STAGES],
        out_done[PIPE_MAX_STAGES], nc_of[PIPE_MAX_STAGES], q_head;
    const GridDev& g = A.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (int)(blockDim.x >> 5) - IBIS_PIPE_G_LOADERS - 1;
    const int svs = T + g.halo_cap;
    constexpr int WD_BYTES = 2 * NF * 3 * T * 8;
    const unsigned in_bytes = (unsigned)grad_stage_bytes<S>(NF, T, g.grec_stride, g.halo_cap);
    const unsigned out_bytes = (unsigned)grad_out_bytes<S>(g.stage_cap);
    const unsigned off_dst = (unsigned)g.stage_cap * (unsigned)SIDE_REC * (unsigned)sizeof(S);   // {home, n_stage} inside a staging buffer
    unsigned char* out_base = smem + (unsigned)nsi * in_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nsi; s++) {
            ac::mbar_init(&full[s], 33);          // the loader's arrive.expect_tx + one cp.async arrival per lane
            stage_tile[s] = -1;
            stage_free[s] = -1;
            done_tasks[s] = 0;
        }
        for (int o = 0; o < nso; o++) { out_done[o] = -1; out_ready[o] = -1; }
        q_head = 0;
        ac::fence_barrier_init();
    }
    __syncthreads();
    const int n_my = (int)blockIdx.x < n_list ? (n_list - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (warp == ncw + IBIS_PIPE_G_LOADERS) {
        // ---- the store warp: when a tile's side records are complete in a staging buffer (out_ready) the image leaves
        // as one bulk store into the tile's side-state array; the buffer is free again (out_done) once the store has
        // read it ----
        for (int k = 0; k < n_my; k++) {
            const int o = k % nso;
            if (!pipe_spin_ge(&out_ready[o], k, A.pipe_dbg, 16 + PW_DONE, k)) return;
            ac::fence_proxy_async();
            const unsigned char* ob = out_base + (unsigned)o * out_bytes;
            const S* stg = reinterpret_cast<const S*>(ob);
            const int* pl = reinterpret_cast<const int*>(ob + off_dst);
            if (lane == 0) {
                const int home = pl[0], n_stage = pl[1];
                if (n_stage > 0) ac::bulk_s2g(A.sidesP + (size_t)home * SIDE_REC, stg, (uint32_t)n_stage * (uint32_t)SIDE_REC * (uint32_t)sizeof(S));
                ac::bulk_commit_wait_read();               // the store warp has nothing else to do than to see it drained
                pipe_fence();
                *reinterpret_cast<volatile int*>(&out_done[o]) = k;
            }
            __syncwarp();
        }
        return;
    }
    if (warp >= ncw) {
        // ---- loaders: loader w takes the block's tiles w, w + NL, ...  Per tile the WLS slice, the record and five
        // state rows as bulk copies, the halo list's state as cp.async gathers; the halo ids of the loader's NEXT tile
        // are read from global memory (L2: prefetched) one tile ahead, so that the gathers of a tile are issued in the
        // same breath as its bulk copies.  One warp issues a tile's ~25 dependent instructions per lane in about a
        // microsecond: several loaders keep the stages full ----
        const int step = (int)gridDim.x, w = warp - ncw;
        constexpr int NL = IBIS_PIPE_G_LOADERS;
        const int PF = (pf + NL - 1) / NL * NL;
        const int hoff = 4 * NF * T + 32;                                     // halo list inside the record
        auto tile_at = [&](int k) { const int idx = (int)blockIdx.x + k * step; return tile_list ? __ldg(tile_list + idx) : idx; };
        int ids[PIPE_HALO_REGS], n_halo = 0, ids_n[PIPE_HALO_REGS], n_halo_n = 0;
        auto fetch_ids = [&](int t, int* out, int& nh) {
            const unsigned char* rec = g.t_grec + (size_t)t * g.grec_stride;
            nh = __ldg(reinterpret_cast<const int*>(rec + 4 * NF * T));
            const int* hl = reinterpret_cast<const int*>(rec + hoff);
#pragma unroll
            for (int q = 0; q < PIPE_HALO_REGS; q++) out[q] = (lane + 32 * q < g.halo_cap) ? __ldg(hl + lane + 32 * q) : 0;
        };
        if (w < n_my) fetch_ids(tile_at(w), ids, n_halo);
        for (int k = w; k < n_my; k += NL) {
            const int t = tile_at(k), s = k % nsi;
            if (PF > 0 && k + PF < n_my && lane < 7) {        // L2 prefetch of a later tile of this loader
                const int tp = tile_at(k + PF);
                int c0p, ncp;
                tile_cells(g, tp, c0p, ncp);
                if (lane == 0) ac::bulk_prefetch_l2(g.t_wd + (size_t)tp * (WD_BYTES / 8), (uint32_t)WD_BYTES);
                else if (lane == 1) ac::bulk_prefetch_l2(g.t_grec + (size_t)tp * g.grec_stride, (uint32_t)g.grec_stride);
                else if (((size_t)c0p * sizeof(S)) % 16 == 0) ac::bulk_prefetch_l2(A.rc.val[lane - 2] + c0p, (uint32_t)T * (uint32_t)sizeof(S));
            }
            if (k + NL < n_my) fetch_ids(tile_at(k + NL), ids_n, n_halo_n);   // in flight during the wait below
            if (k >= nsi && !pipe_spin_ge(&stage_free[s], k - nsi, A.pipe_dbg, 16 + PW_EMPTY, k)) return;
            unsigned char* st = smem + (unsigned)s * in_bytes;
            S* sv = reinterpret_cast<S*>(st + WD_BYTES + g.grec_stride);
            int c0, nc;
            tile_cells(g, t, c0, nc);
            // own cells: contiguous rows (bulk copies when 16-byte aligned), halo list: gathered
            const bool bulk_own = ((size_t)c0 * sizeof(S)) % 16 == 0;
            if (lane == 0) {
                nc_of[s] = nc;
                pipe_fence();
                *reinterpret_cast<volatile int*>(&stage_tile[s]) = k;
                ac::mbar_arrive_expect_tx(&full[s], (uint32_t)WD_BYTES + (uint32_t)g.grec_stride +
                                                        (bulk_own ? 5u * (uint32_t)T * (uint32_t)sizeof(S) : 0u));
                ac::bulk_g2s(st, g.t_wd + (size_t)t * (WD_BYTES / 8), (uint32_t)WD_BYTES, &full[s]);
                ac::bulk_g2s(st + WD_BYTES, g.t_grec + (size_t)t * g.grec_stride, (uint32_t)g.grec_stride, &full[s]);
                if (bulk_own)
                    for (int v = 0; v < 5; v++)
                        ac::bulk_g2s(sv + (size_t)v * svs, A.rc.val[v] + c0, (uint32_t)T * (uint32_t)sizeof(S), &full[s]);
            }
            __syncwarp();
            if (!bulk_own)
                for (int q = lane; q < 5 * T; q += 32) {
                    const int v = q / T, i = q % T;
                    ac::cp_async_small<(int)sizeof(S)>(sv + (size_t)v * svs + i, A.rc.val[v] + c0 + i);
                }
#pragma unroll
            for (int q = 0; q < PIPE_HALO_REGS; q++) {
                const int h = lane + 32 * q;
                if (h < n_halo) {
#pragma unroll
                    for (int v = 0; v < 5; v++) ac::cp_async_small<(int)sizeof(S)>(sv + (size_t)v * svs + T + h, A.rc.val[v] + ids[q]);
                }
            }
            ac::cp_async_mbar_arrive_noinc(&full[s]);
#pragma unroll
            for (int q = 0; q < PIPE_HALO_REGS; q++) ids[q] = ids_n[q];
            n_halo = n_halo_n;
        }
        return;
    }
    // ---- consumers ----
    constexpr int NV = DIM == 3 ? 3 : 2, N_TASKS = T / 32 + NV * (IBIS_PIPE_G_SPLITVEL ? T / 32 : 1);
    const int total = n_my * N_TASKS;
    const unsigned mT = pipe_magic(N_TASKS), mI = pipe_magic(nsi), mO = pipe_magic(nso);
    for (;;) {
        const int item = pipe_next_item(&q_head, lane);
        if (item >= total) break;
        const int k = pipe_div(item, mT), r = item - k * N_TASKS;
        const int ui = pipe_div(k, mI);
        const int s = k - ui * nsi, o = k - pipe_div(k, mO) * nso;
        if (!pipe_spin_ge(&stage_tile[s], k, A.pipe_dbg, 16 + PW_STAGE, k)) return;
        if (!pipe_wait(&full[s], ui & 1, A.pipe_dbg, 16 + PW_FULL, k)) return;
        // the staging buffer is free once the stores of the tile nso tiles back have read it
        if (k >= nso && !pipe_spin_ge(&out_done[o], k - nso, A.pipe_dbg, 16 + PW_OUT, k)) return;
        unsigned char* st = smem + (unsigned)s * in_bytes;
        const GRec R = grec_view(st + WD_BYTES, NF, T);
        S* stg = reinterpret_cast<S*>(out_base + (unsigned)o * out_bytes);
        const double* wd = reinterpret_cast<const double*>(st);
        const S* sv = reinterpret_cast<const S*>(st + WD_BYTES + g.grec_stride);
        grad_item<DIM, NF, T, S>(A, wd, R, sv, svs, stg, r, nc_of[s], lane);
        ac::fence_proxy_async();                           // the staging writes -> visible to the TMA unit
        __syncwarp();
        int d = 0;
        if (lane == 0) {
            pipe_fence();
            d = atomicAdd(&done_tasks[s], 1);
        }
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d == N_TASKS - 1) {
            // the tile's last item: its side records are complete in the staging buffer.  Their destination moves next
            // to them, the input stage goes back to the loaders, and the store warp takes over the buffer
            pipe_fence();
            if (lane == 0) {
                int* pl = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(stg) + off_dst);
                pl[0] = R.hdr[5];
                pl[1] = R.hdr[3];
                done_tasks[s] = 0;
                pipe_fence();
                *reinterpret_cast<volatile int*>(&stage_free[s]) = k;
                *reinterpret_cast<volatile int*>(&out_ready[o]) = k;
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
    // staging buffers: four when they fit (items + drain of a tile overlap the next tiles'), the rest input stages
    int nso = 4;
    while (nso > 1 && (size_t)nso * out_b + 2 * in_b > PIPE_SMEM_BUDGET) nso--;
    int nsi = (size_t)nso * out_b < PIPE_SMEM_BUDGET ? (int)((PIPE_SMEM_BUDGET - (size_t)nso * out_b) / in_b) : 0;
    if (nsi > PIPE_MAX_STAGES) nsi = PIPE_MAX_STAGES;
    if (nsi < 2 || nso < 2 || A.g.halo_cap > 32 * PIPE_HALO_REGS) {
        if (c->last_error.empty()) c->last_error = "pass G': a tile does not fit shared memory";
        return;
    }
    const size_t smem = (size_t)nsi * in_b + (size_t)nso * out_b;
    auto kernel = k_grad_pipe<DIM, NF, T, S>;
    if (!set_smem(c, kernel, smem)) return;
    const int grid = n_list < c->n_sm ? n_list : c->n_sm;
    kernel<<<grid, 32 * (IBIS_PIPE_G_WARPS + IBIS_PIPE_G_LOADERS + 1), smem, c->stream>>>(A, list, n_list, nsi, nso, c->pipe_prefetch);
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
    const size_t in_b = flux_in_bytes<S>(A.g.p_cap, A.g.freca_stride);
    const size_t fb_b = flux_fb_bytes<S>(A.g.tile_cap, A.g.frec_stride, A.g.tile, A.dt_out != nullptr);
    // three input stages cover the load latency; the rest of shared memory holds flux buffers
    // (IBIS_B200_F_STAGES: sweep knob)
    int nsi = c->pipe_f_stages > 0 ? c->pipe_f_stages : 3;
    int nfb = PIPE_SMEM_BUDGET > (size_t)nsi * in_b ? (int)((PIPE_SMEM_BUDGET - (size_t)nsi * in_b) / fb_b) : 0;
    if (nfb < 3) {
        nsi = 2;
        nfb = PIPE_SMEM_BUDGET > (size_t)nsi * in_b ? (int)((PIPE_SMEM_BUDGET - (size_t)nsi * in_b) / fb_b) : 0;
    }
    if (nfb > PIPE_MAX_STAGES) {
        nfb = PIPE_MAX_STAGES;
        const int more = (int)((PIPE_SMEM_BUDGET - (size_t)nfb * fb_b) / in_b);
        nsi = more > PIPE_MAX_STAGES ? PIPE_MAX_STAGES : more;
    }
    if (nfb < 2) {
        if (c->last_error.empty()) c->last_error = "pass F': a tile does not fit shared memory";
        return;
    }
    const size_t smem = (size_t)nsi * in_b + (size_t)nfb * fb_b;
    auto kernel = k_flux_pipe<DIM, FLUX, S>;
    if (!set_smem(c, kernel, smem)) return;
    const int grid = A.g.n_tiles < c->n_sm ? A.g.n_tiles : c->n_sm;
    kernel<<<grid, 32 * (IBIS_PIPE_F_WARPS + IBIS_PIPE_F_LOADERS), smem, c->stream>>>(A, nsi, nfb, c->pipe_prefetch);
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
