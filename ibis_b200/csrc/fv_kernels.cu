// fv_kernels.cu -- sm_100a kernels of the finite-volume residual:
//   k_bc            ghost-cell fills of every boundary tag in one launch
//   k_wls_r         least-squares r-factors (setup)
//   k_grad_limit    pass G: all WLS gradients + Barth-Jespersen limiters of a cell, fused
//   k_residual      pass F: reconstruct + rotate + flux + rotate back + surface integral,
//                   owner-computes per cell (no face temporaries, no atomics)
//   k_face_flux / k_face_viscous / k_integral   general (viscous) path through a face array
//   k_local_dt_min  local time step + block min
// One thread per cell / face, SoA loads coalesced along the cell index, neighbour gathers
// served by L1/L2.  Built with --fmad=false: arithmetic is bit-for-bit the reference's.
#include <cuda_runtime.h>

#include "async_copy.h"
#include "context.h"

// tuning knobs (compile-time; tools/build_variants.sh builds alternatives for sweeps)
#ifndef IBIS_G_BLOCK
#define IBIS_G_BLOCK 256     // threads per block of pass G
#endif
#ifndef IBIS_G_MINB
#define IBIS_G_MINB 2        // min resident blocks per SM of pass G (register cap)
#endif
#ifndef IBIS_F_MAXT
#define IBIS_F_MAXT 256      // max threads per block of pass F (>= tile size)
#endif
#ifndef IBIS_F_MINB
#define IBIS_F_MINB 2        // min resident blocks per SM of pass F (register cap)
#endif

#ifndef IBIS_GS_MINB
#define IBIS_GS_MINB 2       // min resident blocks per SM of pass G'
#endif
#ifndef IBIS_GS_TPC
#define IBIS_GS_TPC 2        // threads per cell of pass G' (1 | 2)
#endif
#ifndef IBIS_GS_STAGE
#define IBIS_GS_STAGE 1      // stage the static cell tables in shared memory by TMA
#endif
#ifndef IBIS_PREFETCH_AHEAD
#define IBIS_PREFETCH_AHEAD 296   // a block asks the TMA unit to pull the data of the tile this many blocks ahead
#endif                            // (2 resident blocks x 148 SMs: the tile that takes this block's place) into L2; 0 = off
#ifndef IBIS_COPY_TMA
#define IBIS_COPY_TMA 1      // 1: tile transfers as TMA bulk copies; 0: 16-byte cp.async / st.global by all threads
#endif
#ifndef IBIS_F2_THREADS
#define IBIS_F2_THREADS 256  // threads per block of pass F' (>= tile size)
#endif
#ifndef IBIS_F2_MINB
#define IBIS_F2_MINB 2       // resident blocks per SM of pass F'
#endif

namespace ib {

#define CUDA_CHECK_LAUNCH(c)                                                             \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess && (c)->last_error.empty())                               \
            (c)->last_error = std::string("kernel launch: ") + cudaGetErrorString(e__);  \
        (c)->launches++;                                                                 \
    } while (0)

static inline int nblocks(int64_t n, int bs) { return n > 0 ? (int)((n + bs - 1) / bs) : 1; }   // an empty grid launches one idle block

void prof_begin(ibis_b200_ctx* c, int kind) {
    if (!c->prof_on) return;
    auto& ev = c->prof_ev[kind];
    if (c->prof_used[kind] + 2 > ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        ev.push_back(a);
        ev.push_back(b);
    }
    cudaEventRecord((cudaEvent_t)ev[c->prof_used[kind]], c->stream);
}
void prof_end(ibis_b200_ctx* c, int kind) {
    if (!c->prof_on) return;
    cudaEventRecord((cudaEvent_t)c->prof_ev[kind][c->prof_used[kind] + 1], c->stream);
    c->prof_used[kind] += 2;
}
int prof_read(ibis_b200_ctx* c, double* ms, int64_t* count) {
    cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 4; k++) {
        double t = 0.0;
        for (size_t i = 0; i + 1 < c->prof_used[k]; i += 2) {
            float e = 0.f;
            cudaEventElapsedTime(&e, (cudaEvent_t)c->prof_ev[k][i], (cudaEvent_t)c->prof_ev[k][i + 1]);
            t += e;
        }
        ms[k] = t;
        count[k] = (int64_t)(c->prof_used[k] / 2);
        c->prof_used[k] = 0;
    }
    return 0;
}

template <typename S>
__global__ void __launch_bounds__(128) k_bc(const FvArgs<S> A, int phase) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < A.g.n_bfaces) bc_body(A, k, phase);
}

__global__ void __launch_bounds__(256) k_wls_r(const GridDev g, double* r_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < g.n_grad) wls_r_body(g, r_out, c);
}

template <int DIM, int NV, int NF, typename S>
__global__ void __launch_bounds__(IBIS_G_BLOCK, IBIS_G_MINB) k_grad_limit(const FvArgs<S> A, const GradVars<S, NV> V,
                                                     const int do_limit) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < A.g.n_grad) grad_limit_body<DIM, NV, NF>(A, V, c, do_limit != 0);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_cell_static(const GridDev g, double* w_out, double* d_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < g.n_grad) cell_static_body<DIM>(g, w_out, d_out, c);
}

// pass F, tiled: blockDim.x == A.g.tile; dynamic shared memory = 12 * tile_cap * sizeof(S)
template <int DIM, int FLUX, bool O2, typename S>
__global__ void __launch_bounds__(IBIS_F_MAXT, IBIS_F_MINB) k_residual_tile(const FvArgs<S> A) {
    extern __shared__ double smem[];
    S* st = reinterpret_cast<S*>(smem);
    const int tile = blockIdx.x;
    const int c = tile * A.g.tile + threadIdx.x;
    const bool has_cell = c < A.g.n_owned;
    if (has_cell) tile_recon_own_body<DIM, O2>(A, c, st);
    for (int e = A.g.out_off[tile] + threadIdx.x; e < A.g.out_off[tile + 1]; e += blockDim.x)
        tile_recon_outside_body<DIM, O2>(A, e, st);
    __syncthreads();
    const int n = tile_num_slots(A.g, tile);
    for (int s = threadIdx.x; s < n; s += blockDim.x) tile_flux_body<DIM, FLUX>(A, tile, s, st);
    __syncthreads();
    if (has_cell) tile_cell_body<DIM>(A, c, st, 1, A.g.tile_cap);
}

__device__ __forceinline__ double warp_min(double v);

// block-cooperative copies of a contiguous range (bytes: multiple of 16, both ends 16-byte aligned)
__device__ __forceinline__ void coop_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes) {
    for (uint32_t o = threadIdx.x * 16u; o < bytes; o += blockDim.x * 16u)
        ac::cp_async_16(static_cast<unsigned char*>(dst_smem) + o, static_cast<const unsigned char*>(src_gmem) + o);
}
__device__ __forceinline__ void coop_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    for (uint32_t o = threadIdx.x * 16u; o < bytes; o += blockDim.x * 16u)
        *reinterpret_cast<double2*>(static_cast<unsigned char*>(dst_gmem) + o) =
            *reinterpret_cast<const double2*>(static_cast<const unsigned char*>(src_smem) + o);
}

// pass G' of the sides path: one block per tile (owned tiles, then the tiles of the inner halo
// ring), TPC threads per cell (fv_bodies.h: grad_sides_*; TPC = 2: half 0 = thermodynamic variables,
// half 1 = velocity).  Data movement of a block:
//   in   (STAGE) the tile's slice of the static tables (WLS weights, face offsets; 36 KB for 128
//        hexes): bulk async copies (TMA) into shared memory, in flight while the threads gather;
//   out  the side states of the tile go to a staging buffer in shared memory and leave as a handful
//        of bulk async stores (TMA): the tile-owned interface numbering (layout.h) makes the sides a
//        tile produces contiguous ranges of the global side arrays.
template <int DIM, int NF, int TS, int TPC, bool STAGE, typename S>
__global__ void __launch_bounds__(TPC * TS, IBIS_GS_MINB) k_grad_sides(const FvArgs<S> A, const int* __restrict__ tile_list) {
    constexpr int NT = NF * 3;                                  // arrays per table
    constexpr int WS = STAGE ? TS : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    double* tab = reinterpret_cast<double*>(smem_raw);          // [2 * NT][TS] when STAGE
    S* stg = reinterpret_cast<S*>(smem_raw + (STAGE ? (size_t)2 * NT * TS * sizeof(double) : 0));
    const GridDev& g = A.g;
    const int tile = tile_list ? tile_list[blockIdx.x] : blockIdx.x;
    const int* ti = g.tile_info + TILE_INFO * tile;
    int c0, nc;
    tile_cells(g, tile, c0, nc);                               // no load on the way to the gathers
    const int nR = ti[4], n_ent = ti[7], n_own_l = ti[8];
    const int lane_cell = threadIdx.x % TS, half = threadIdx.x / TS;
    const int c = c0 + lane_cell;
    const bool bulk_in = IBIS_COPY_TMA && STAGE && (c0 & 1) == 0;   // 16-byte aligned table slices
    const bool async_in = !IBIS_COPY_TMA && STAGE && (c0 & 1) == 0;
    if (STAGE) {
        if (async_in) {
            for (int arr = 0; arr < 2 * NT; arr++) {
                const int a = arr % NT;
                coop_g2s(tab + arr * TS, (arr < NT ? g.cf_w : g.cf_d) + (long)a * g.cstride + c0, TS * (uint32_t)sizeof(double));
            }
        } else if (bulk_in) {
            if (threadIdx.x == 0) {
                ac::mbar_init(&bar, 2 * NT);
                ac::fence_barrier_init();
            }
            __syncthreads();
            if (threadIdx.x < 2 * NT) {
                const int a = threadIdx.x % NT;
                const double* src = (threadIdx.x < NT ? g.cf_w : g.cf_d) + (long)a * g.cstride + c0;
                ac::mbar_arrive_expect_tx(&bar, TS * (uint32_t)sizeof(double));
                ac::bulk_g2s(tab + threadIdx.x * TS, src, TS * (uint32_t)sizeof(double), &bar);
            }
        } else {
            for (int q = threadIdx.x; q < 2 * NT * TS; q += blockDim.x) {
                const int arr = q / TS, t = q % TS, a = arr % NT;
                tab[q] = (arr < NT ? g.cf_w : g.cf_d)[(long)a * g.cstride + c0 + t];
            }
            __syncthreads();
        }
    }
    if (IBIS_PREFETCH_AHEAD > 0 && (int)blockIdx.x + IBIS_PREFETCH_AHEAD < (int)gridDim.x && threadIdx.x >= 64 &&
        threadIdx.x < 64 + 2 * NT + 5 + 3 * NF) {
        // warps 2.. : the data of the tile that will take this block's place, into L2
        const int pt = tile_list ? tile_list[blockIdx.x + IBIS_PREFETCH_AHEAD] : (int)blockIdx.x + IBIS_PREFETCH_AHEAD;
        int pc0, pnc;
        tile_cells(g, pt, pc0, pnc);
        const int q = threadIdx.x - 64;
        if ((pc0 & 1) == 0) {
            if (q < 2 * NT) {
                ac::bulk_prefetch_l2((q < NT ? g.cf_w : g.cf_d) + (long)(q % NT) * g.cstride + pc0, TS * 8u);
            } else if (q < 2 * NT + 5) {
                ac::bulk_prefetch_l2(A.rc.val[q - 2 * NT] + pc0, TS * (uint32_t)sizeof(S));
            } else {
                const int r = q - 2 * NT - 5, j = r % NF, which = r / NF;
                if (which == 0 && (pc0 & 3) == 0) ac::bulk_prefetch_l2(g.cf_nbr + (long)j * g.cstride + pc0, TS * 4u);
                else if (which == 1 && (pc0 & 7) == 0) ac::bulk_prefetch_l2(g.cf_stage + (long)j * g.cstride + pc0, TS * 2u);
                else if (which == 2 && (pc0 & 15) == 0) ac::bulk_prefetch_l2(g.cf_out + (long)j * g.cstride + pc0, TS * 1u);
            }
        }
    }
    const bool active = lane_cell < nc;
    CellTables<WS> T;
    T.w = STAGE ? tab : g.cf_w;
    T.d = STAGE ? tab + NT * TS : g.cf_d;
    T.stride_rt = STAGE ? TS : g.cstride;
    T.idx = STAGE ? lane_cell : c;
    if (TPC == 1) {
        HalfGather<NF, 2, S> G0;
        HalfGather<NF, 3, S> G1;
        if (active) {
            grad_sides_gather<DIM, NF, 0>(A, c, G0);
            grad_sides_gather<DIM, NF, 1>(A, c, G1);
        }
        if (bulk_in) ac::mbar_wait(&bar, 0);
        if (async_in) { ac::cp_async_wait_all(); __syncthreads(); }
        if (active) {
            grad_sides_compute<DIM, NF, 0, WS>(A, G0, T, nR, n_own_l, stg);
            grad_sides_compute<DIM, NF, 1, WS>(A, G1, T, nR, n_own_l, stg);
        }
    } else if (half == 0) {
        HalfGather<NF, 2, S> G;
        if (active) grad_sides_gather<DIM, NF, 0>(A, c, G);
        if (bulk_in) ac::mbar_wait(&bar, 0);
        if (async_in) { ac::cp_async_wait_all(); __syncthreads(); }
        if (active) grad_sides_compute<DIM, NF, 0, WS>(A, G, T, nR, n_own_l, stg);
    } else {
        HalfGather<NF, 3, S> G;
        if (active) grad_sides_gather<DIM, NF, 1>(A, c, G);
        if (bulk_in) ac::mbar_wait(&bar, 0);
        if (async_in) { ac::cp_async_wait_all(); __syncthreads(); }
        if (active) grad_sides_compute<DIM, NF, 1, WS>(A, G, T, nR, n_own_l, stg);
    }
#if IBIS_COPY_TMA
    ac::fence_proxy_async();                                    // staging writes -> visible to the TMA unit
    __syncthreads();
    if (threadIdx.x < 2 + n_ent) {
        const TileXfer x = tile_store_xfer(g, tile, threadIdx.x, n_own_l);
        if (x.count > 0) {
            S* dst = (x.left ? A.sideL : A.sideR) + (long)x.f0 * 6;
            ac::bulk_s2g(dst, stg + (long)x.stage * 6, (uint32_t)x.count * 6 * (uint32_t)sizeof(S));
        }
        ac::bulk_commit_wait_read();                            // shared memory must outlive the reads
    }
#else
    __syncthreads();
    for (int q = 0; q < 2 + n_ent; q++) {
        const TileXfer x = tile_store_xfer(g, tile, q, n_own_l);
        S* dst = (x.left ? A.sideL : A.sideR) + (long)x.f0 * 6;
        coop_s2g(dst, stg + (long)x.stage * 6, (uint32_t)x.count * 6 * (uint32_t)sizeof(S));
    }
#endif
}

// pass F' of the sides path: one block per owned tile.  Left states, right states and frames of the
// tile's interfaces arrive as TMA bulk copies of contiguous ranges (layout.h) on one mbarrier (splitting
// the fetch over two barriers so that the block starts on the first half measured 5 % slower).  Each flux
// overwrites the left state of its own interface; then one thread per cell sums its faces (owner
// computes, no atomics).  Several resident blocks per SM overlap each other's fetches; a persistent
// two-stage pipeline (one 512-thread block per SM) measured SLOWER (profiles/r01_summary.md, r01g)
// and was dropped.
template <typename S>
__device__ __forceinline__ void flux_tile_issue(const FvArgs<S>& A, int tile, S* stL, S* stR, double* stG,
                                                uint64_t* bar) {
    // called by warp 0: lane 0 arms the barrier with the byte count, lanes issue the copies
    const GridDev& g = A.g;
    const int* ti = g.tile_info + TILE_INFO * tile;
    const int n_ent = ti[7], n_slots = ti[9];
    if (threadIdx.x == 0) ac::mbar_arrive_expect_tx(bar, (uint32_t)n_slots * (uint32_t)frec<S>::bytes);
    __syncwarp();
    const int n_x = 1 + n_ent;
    for (int q = threadIdx.x; q < 3 * n_x; q += 32) {
        const TileXfer x = tile_load_xfer(g, tile, q / 3);
        if (x.count <= 0) continue;
        const int which = q % 3;
        const uint32_t bytes = (uint32_t)x.count * (which == 2 ? 80u : 6u * (uint32_t)sizeof(S));
        if (which == 0) ac::bulk_g2s(stL + (long)x.slot * 6, A.sideL + (long)x.f0 * 6, bytes, bar);
        else if (which == 1) ac::bulk_g2s(stR + (long)x.slot * 6, A.sideR + (long)x.f0 * 6, bytes, bar);
        else ac::bulk_g2s(stG + (long)x.slot * 10, g.f_geom + (long)x.f0 * 10, bytes, bar);
    }
}

// the transfers of another tile as L2 prefetches (called by one warp other than warp 0)
template <typename S>
__device__ __forceinline__ void flux_tile_prefetch(const FvArgs<S>& A, int tile) {
    const GridDev& g = A.g;
    const int n_x = 1 + g.tile_info[TILE_INFO * tile + 7];
    for (int q = threadIdx.x & 31; q < 3 * n_x; q += 32) {
        const TileXfer x = tile_load_xfer(g, tile, q / 3);
        if (x.count <= 0) continue;
        const int which = q % 3;
        const uint32_t bytes = (uint32_t)x.count * (which == 2 ? 80u : 6u * (uint32_t)sizeof(S));
        if (which == 0) ac::bulk_prefetch_l2(A.sideL + (long)x.f0 * 6, bytes);
        else if (which == 1) ac::bulk_prefetch_l2(A.sideR + (long)x.f0 * 6, bytes);
        else ac::bulk_prefetch_l2(g.f_geom + (long)x.f0 * 10, bytes);
    }
}

// the same transfers with 16-byte cp.async issued by every thread of the block
template <typename S>
__device__ __forceinline__ void flux_tile_issue_coop(const FvArgs<S>& A, int tile, S* stL, S* stR, double* stG) {
    const GridDev& g = A.g;
    const int n_x = 1 + g.tile_info[TILE_INFO * tile + 7];
    for (int q = 0; q < n_x; q++) {
        const TileXfer x = tile_load_xfer(g, tile, q);
        coop_g2s(stL + (long)x.slot * 6, A.sideL + (long)x.f0 * 6, (uint32_t)x.count * 6u * (uint32_t)sizeof(S));
        coop_g2s(stR + (long)x.slot * 6, A.sideR + (long)x.f0 * 6, (uint32_t)x.count * 6u * (uint32_t)sizeof(S));
        coop_g2s(stG + (long)x.slot * 10, g.f_geom + (long)x.f0 * 10, (uint32_t)x.count * 80u);
    }
}

template <int DIM, int FLUX, typename S>
__global__ void __launch_bounds__(IBIS_F2_THREADS, IBIS_F2_MINB) k_flux_tile(const FvArgs<S> A, const int n_stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const GridDev& g = A.g;
    const int cap = g.tile_cap;
    S* stL = reinterpret_cast<S*>(smem_raw);
    S* stR = stL + (size_t)cap * 6;
    double* stG = reinterpret_cast<double*>(stR + (size_t)cap * 6);
    const int tile = blockIdx.x;
#if IBIS_COPY_TMA
    if (threadIdx.x == 0) {
        ac::mbar_init(&bar, 1);
        ac::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) flux_tile_issue<S>(A, tile, stL, stR, stG, &bar);
    else if (IBIS_PREFETCH_AHEAD > 0 && threadIdx.x >= 64 && threadIdx.x < 96 && tile + IBIS_PREFETCH_AHEAD < g.n_tiles)
        flux_tile_prefetch<S>(A, tile + IBIS_PREFETCH_AHEAD);
#else
    flux_tile_issue_coop<S>(A, tile, stL, stR, stG);
#endif
    const int* ti = g.tile_info + TILE_INFO * tile;
    int c0, nc;
    tile_cells(g, tile, c0, nc);
    const int n = ti[9];
    const int c = c0 + threadIdx.x;
    const bool has_cell = threadIdx.x < nc;
    CellFaces CF;
    if (has_cell) tile_cell_load(g, c, CF);
#if IBIS_COPY_TMA
    ac::mbar_wait(&bar, 0);
#else
    ac::cp_async_wait_all();
    __syncthreads();
    (void)bar;
#endif
    for (int s = threadIdx.x; s < n; s += blockDim.x)
        flux_from_sides<DIM, FLUX>(A, stL + (long)s * 6, stR + (long)s * 6, stG + (long)s * 10, stL + (long)s * 6, 1);
    __syncthreads();
    if (has_cell) tile_cell_sum<DIM>(A, c, CF, stL, 6, 1);
    if (A.dt_out) {                         // estimate_dt fused into this evaluation (uniform branch)
        double d = has_cell ? tile_cell_dt<S>(A, c, CF, stG) : INFINITY;
        d = warp_min(d);
        __shared__ double dmin[32];
        if ((threadIdx.x & 31) == 0) dmin[threadIdx.x >> 5] = d;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); w++) d = dmin[w] < d ? dmin[w] : d;
            // order-preserving key of a double (negative local time steps of inverted cells included: the
            // reference's min picks them, finite_volume.cpp:180): flip all bits of negatives, the sign of
            // the others; NaN never wins a `<` in the reference and is skipped here
            if (d == d) {
                const unsigned long long b = (unsigned long long)__double_as_longlong(d);
                atomicMin(A.dt_out, b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull));
            }
        }
    }
    (void)n_stages;
}

template <int DIM, int FLUX, bool O2, typename S>
__global__ void __launch_bounds__(128) k_face_flux(const FvArgs<S> A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < A.g.n_faces) face_flux_body<DIM, FLUX, O2>(A, f);
}

// moving grid: interface-velocity-relative flux (convective_flux.cpp:105-140,166-176)
template <int DIM, int FLUX, bool O2, typename S>
__global__ void __launch_bounds__(128) k_face_flux_moving(const FvArgs<S> A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < A.g.n_faces) face_flux_body<DIM, FLUX, O2, S, true>(A, f);
}

__global__ void __launch_bounds__(256) k_face_vel(const GridDev g, const int* fverts, const double* vvel, double* face_vel) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < g.n_faces) face_vel_body(g, fverts, vvel, face_vel, f);
}

// GridBlock::compute_geometric_data after a vertex move (grid.h:169-178): 0 interfaces, 1 centroids,
// 2 volumes, 3 ghost-cell centres
template <int OP>
__global__ void __launch_bounds__(256) k_geom_update(const GridDev g, const GeomUpdate U, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (OP == 0) geom_face_body(g, U, i);
    else if (OP == 1) geom_centroid_body(g, U, i);
    else if (OP == 2) geom_volume_body(g, U, i);
    else geom_ghost_body(g, U, i);
}
__global__ void __launch_bounds__(256) k_move_vertices(double* pos, const double* vel, const double scale, const long n3) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) pos[i] = pos[i] + vel[i] * scale;          // add_scaled_vector (util/vector3.cpp:39-47)
}

template <typename S>
struct CqPtrs { S* p[5]; };
template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_gcl(const FvArgs<S> A, const CqPtrs<S> cq) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < A.g.n_owned) gcl_body<DIM>(A, cq.p, c);
}

template <int DIM, typename S>
__global__ void __launch_bounds__(128) k_face_viscous(const FvArgs<S> A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < A.g.n_faces) face_viscous_body<DIM>(A, f);
}

template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_integral(const FvArgs<S> A) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < A.g.n_owned) integral_body<DIM>(A, c);
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w < v) ? w : v;
    }
    return v;
}

// local dt of every owned cell, min over the block -> partial[blockIdx.x]
template <typename S>
__global__ void __launch_bounds__(256) k_local_dt_min(const FvArgs<S> A, double* partial) {
    __shared__ double sm[8];
    double v = INFINITY;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < A.g.n_owned;
         c += gridDim.x * blockDim.x) {
        const double d = local_dt_body(A, c);
        v = (d < v) ? d : v;                                   // Kokkos::min(dt_local, dt_utd)
    }
    v = warp_min(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : INFINITY;
        v = warp_min(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) k_final_min(const double* partial, int n, double* out) {
    __shared__ double sm[8];
    double v = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v = (partial[i] < v) ? partial[i] : v;
    v = warp_min(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : INFINITY;
        v = warp_min(v);
        if (threadIdx.x == 0) out[0] = v;
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
template <typename S>
void launch_bc(ibis_b200_ctx* c, const FvArgs<S>& A, int phase) {
    if (A.g.n_bfaces == 0) return;
    k_bc<S><<<nblocks(A.g.n_bfaces, 128), 128, 0, c->stream>>>(A, phase);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_bc<double>(ibis_b200_ctx*, const FvArgs<double>&, int);
template void launch_bc<dual>(ibis_b200_ctx*, const FvArgs<dual>&, int);

void launch_wls_r(ibis_b200_ctx* c, double* r_out) {
    if (c->g.n_grad == 0) return;
    k_wls_r<<<nblocks(c->g.n_grad, 256), 256, 0, c->stream>>>(c->g, r_out);
    CUDA_CHECK_LAUNCH(c);
}

void launch_cell_static(ibis_b200_ctx* c, double* w_out, double* d_out) {
    if (c->g.n_grad == 0) return;
    if (c->g.dim == 3) k_cell_static<3><<<nblocks(c->g.n_grad, 256), 256, 0, c->stream>>>(c->g, w_out, d_out);
    else k_cell_static<2><<<nblocks(c->g.n_grad, 256), 256, 0, c->stream>>>(c->g, w_out, d_out);
    CUDA_CHECK_LAUNCH(c);
}

template <int DIM, int NV, typename S>
static void launch_grad_limit_nf(ibis_b200_ctx* c, const FvArgs<S>& A, const GradVars<S, NV>& V, int do_limit) {
    const int nb = nblocks(A.g.n_grad, IBIS_G_BLOCK);
    switch (A.g.max_faces) {
        case 6: k_grad_limit<DIM, NV, 6, S><<<nb, IBIS_G_BLOCK, 0, c->stream>>>(A, V, do_limit); break;
        case 5: k_grad_limit<DIM, NV, 5, S><<<nb, IBIS_G_BLOCK, 0, c->stream>>>(A, V, do_limit); break;
        case 4: k_grad_limit<DIM, NV, 4, S><<<nb, IBIS_G_BLOCK, 0, c->stream>>>(A, V, do_limit); break;
        default: k_grad_limit<DIM, NV, 3, S><<<nb, IBIS_G_BLOCK, 0, c->stream>>>(A, V, do_limit); break;
    }
}

template <typename S>
void launch_grad_limit(ibis_b200_ctx* c, const FvArgs<S>& A, bool viscous_pass) {
    if (A.g.n_grad == 0) return;
    const long cs = A.g.cstride;
    if (viscous_pass) {
        // ViscousFlux::compute_viscous_gradient (viscous_flux.cpp:122-130): T, vx, vy, vz
        const int vars[4] = {V_TVISC, V_VX, V_VY, V_VZ};
        const int src[4] = {F_T, F_VX, F_VY, F_VZ};
        if (A.g.dim == 3) {
            GradVars<S, 4> V;
            for (int i = 0; i < 4; i++) {
                V.src[i] = A.fs[src[i]];
                for (int k = 0; k < 3; k++) V.g[i][k] = A.grad + (vars[i] * 3 + k) * cs;
                V.lim[i] = A.lim + vars[i] * cs;
            }
            launch_grad_limit_nf<3, 4, S>(c, A, V, 0);
        } else {
            GradVars<S, 3> V;                                  // vz == 0 in 2-D
            for (int i = 0; i < 3; i++) {
                V.src[i] = A.fs[src[i]];
                for (int k = 0; k < 3; k++) V.g[i][k] = A.grad + (vars[i] * 3 + k) * cs;
                V.lim[i] = A.lim + vars[i] * cs;
            }
            launch_grad_limit_nf<2, 3, S>(c, A, V, 0);
        }
    } else {
        // compute_convective_gradient + compute_limiters (convective_flux.cpp:181-209,423-462)
        const int vars[5] = {V_RHO, V_TH2, V_VX, V_VY, V_VZ};
        const int src[5] = {F_RHO, A.th2, F_VX, F_VY, F_VZ};
        prof_begin(c, 0);
        if (A.g.dim == 3) {
            GradVars<S, 5> V;
            for (int i = 0; i < 5; i++) {
                V.src[i] = A.fs[src[i]];
                for (int k = 0; k < 3; k++) V.g[i][k] = A.grad + (vars[i] * 3 + k) * cs;
                V.lim[i] = A.lim + vars[i] * cs;
            }
            launch_grad_limit_nf<3, 5, S>(c, A, V, A.limiter_on);
        } else {
            GradVars<S, 4> V;
            for (int i = 0; i < 4; i++) {
                V.src[i] = A.fs[src[i]];
                for (int k = 0; k < 3; k++) V.g[i][k] = A.grad + (vars[i] * 3 + k) * cs;
                V.lim[i] = A.lim + vars[i] * cs;
            }
            launch_grad_limit_nf<2, 4, S>(c, A, V, A.limiter_on);
        }
        prof_end(c, 0);
    }
    CUDA_CHECK_LAUNCH(c);
}
template void launch_grad_limit<double>(ibis_b200_ctx*, const FvArgs<double>&, bool);
template void launch_grad_limit<dual>(ibis_b200_ctx*, const FvArgs<dual>&, bool);

#define DISPATCH_FLUX_O2(KERNEL, DIMV, GRID, BS, SMEM)                                                        \
    switch (c->cfg.flux_calculator) {                                                                          \
        case FLUX_HANEL: LAUNCH_O2(KERNEL, DIMV, FLUX_HANEL, GRID, BS, SMEM) break;                            \
        case FLUX_AUSMDV: LAUNCH_O2(KERNEL, DIMV, FLUX_AUSMDV, GRID, BS, SMEM) break;                          \
        case FLUX_LDFSS2: LAUNCH_O2(KERNEL, DIMV, FLUX_LDFSS2, GRID, BS, SMEM) break;                          \
        default: LAUNCH_O2(KERNEL, DIMV, FLUX_RUSANOV, GRID, BS, SMEM) break;                                  \
    }
#define LAUNCH_O2(KERNEL, DIMV, FLUXV, GRID, BS, SMEM)                                                        \
    if (A.order2) { launch_one(c, KERNEL<DIMV, FLUXV, true, S>, GRID, BS, SMEM, A); }                         \
    else { launch_one(c, KERNEL<DIMV, FLUXV, false, S>, GRID, BS, SMEM, A); }

template <typename K, typename S>
static void launch_one(ibis_b200_ctx* c, K kernel, int grid, int bs, size_t smem, const FvArgs<S>& A) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess && c->last_error.empty())
            c->last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    }
    kernel<<<grid, bs, smem, c->stream>>>(A);
}

template <typename S>
void launch_residual(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_tiles == 0) return;
    const size_t smem = (size_t)12 * A.g.tile_cap * sizeof(S);
    prof_begin(c, 1);
    if (A.g.dim == 3) {
        DISPATCH_FLUX_O2(k_residual_tile, 3, A.g.n_tiles, A.g.tile, smem)
    } else {
        DISPATCH_FLUX_O2(k_residual_tile, 2, A.g.n_tiles, A.g.tile, smem)
    }
    prof_end(c, 1);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_residual<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_residual<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <typename K, typename S, typename... Extra>
static void launch_tiles(ibis_b200_ctx* c, K kernel, int grid, int bs, size_t smem, const FvArgs<S>& A, Extra... extra) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess && c->last_error.empty())
        c->last_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    kernel<<<grid, bs, smem, c->stream>>>(A, extra...);
}

template <int DIM, int TS, typename S>
static void launch_grad_sides_ts(ibis_b200_ctx* c, const FvArgs<S>& A, const int* list, int grid) {
    constexpr int TPC = IBIS_GS_TPC;
    constexpr bool STAGE = IBIS_GS_STAGE != 0;
    const int nf = A.g.max_faces < 3 ? 3 : A.g.max_faces;
    const size_t smem = (STAGE ? (size_t)2 * nf * 3 * TS * sizeof(double) : 0) + (size_t)A.g.stage_cap * 6 * sizeof(S);
    switch (nf) {
        case 6: launch_tiles(c, k_grad_sides<DIM, 6, TS, TPC, STAGE, S>, grid, TPC * TS, smem, A, list); break;
        case 5: launch_tiles(c, k_grad_sides<DIM, 5, TS, TPC, STAGE, S>, grid, TPC * TS, smem, A, list); break;
        case 4: launch_tiles(c, k_grad_sides<DIM, 4, TS, TPC, STAGE, S>, grid, TPC * TS, smem, A, list); break;
        default: launch_tiles(c, k_grad_sides<DIM, 3, TS, TPC, STAGE, S>, grid, TPC * TS, smem, A, list); break;
    }
}

// which: 0 = every tile, 1 = the early tiles (no halo cell, no ghost among their neighbours), 2 = the late ones
template <typename S>
void launch_grad_sides(ibis_b200_ctx* c, const FvArgs<S>& A, int which) {
    const int* list = which == 1 ? A.g.tiles_early : which == 2 ? A.g.tiles_late : nullptr;
    const int grid = which == 1 ? A.g.n_early : which == 2 ? A.g.n_late : A.g.n_tiles_all;
    if (grid == 0) return;
    prof_begin(c, 0);
    if (A.g.dim == 3) {
        if (A.g.tile == 64) launch_grad_sides_ts<3, 64, S>(c, A, list, grid);
        else if (A.g.tile == 128) launch_grad_sides_ts<3, 128, S>(c, A, list, grid);
        else launch_grad_sides_ts<3, 256, S>(c, A, list, grid);
    } else {
        if (A.g.tile == 64) launch_grad_sides_ts<2, 64, S>(c, A, list, grid);
        else if (A.g.tile == 128) launch_grad_sides_ts<2, 128, S>(c, A, list, grid);
        else launch_grad_sides_ts<2, 256, S>(c, A, list, grid);
    }
    prof_end(c, 0);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_grad_sides<double>(ibis_b200_ctx*, const FvArgs<double>&, int);
template void launch_grad_sides<dual>(ibis_b200_ctx*, const FvArgs<dual>&, int);

template <int DIM, typename S>
static void launch_flux_tile_dim(ibis_b200_ctx* c, const FvArgs<S>& A, int grid, int bs, size_t smem, int n_stages) {
    switch (c->cfg.flux_calculator) {
        case FLUX_HANEL: launch_tiles(c, k_flux_tile<DIM, FLUX_HANEL, S>, grid, bs, smem, A, n_stages); break;
        case FLUX_AUSMDV: launch_tiles(c, k_flux_tile<DIM, FLUX_AUSMDV, S>, grid, bs, smem, A, n_stages); break;
        case FLUX_LDFSS2: launch_tiles(c, k_flux_tile<DIM, FLUX_LDFSS2, S>, grid, bs, smem, A, n_stages); break;
        default: launch_tiles(c, k_flux_tile<DIM, FLUX_RUSANOV, S>, grid, bs, smem, A, n_stages); break;
    }
}

template <typename S>
void launch_flux_tile(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_tiles == 0) return;
    const size_t stage = (size_t)A.g.tile_cap * frec<S>::bytes;
    const int bs = A.g.tile > IBIS_F2_THREADS ? A.g.tile : IBIS_F2_THREADS;
    const int n_stages = 1;
    const size_t smem = stage;
    const int grid = A.g.n_tiles;
    prof_begin(c, 1);
    if (A.g.dim == 3) launch_flux_tile_dim<3, S>(c, A, grid, bs, smem, n_stages);
    else launch_flux_tile_dim<2, S>(c, A, grid, bs, smem, n_stages);
    prof_end(c, 1);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_flux_tile<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_flux_tile<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

// the fused dt reduction of pass F' works on order-preserving integer keys (see k_flux_tile)
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

template <typename S>
void launch_face_flux(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_faces == 0) return;
    if (A.g.dim == 3) {
        DISPATCH_FLUX_O2(k_face_flux, 3, nblocks(A.g.n_faces, 128), 128, 0)
    } else {
        DISPATCH_FLUX_O2(k_face_flux, 2, nblocks(A.g.n_faces, 128), 128, 0)
    }
    CUDA_CHECK_LAUNCH(c);
}
template void launch_face_flux<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_face_flux<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <typename S>
void launch_face_flux_moving(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_faces == 0) return;
    if (A.g.dim == 3) {
        DISPATCH_FLUX_O2(k_face_flux_moving, 3, nblocks(A.g.n_faces, 128), 128, 0)
    } else {
        DISPATCH_FLUX_O2(k_face_flux_moving, 2, nblocks(A.g.n_faces, 128), 128, 0)
    }
    CUDA_CHECK_LAUNCH(c);
}
template void launch_face_flux_moving<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_face_flux_moving<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

void launch_geom_update(ibis_b200_ctx* c, const GeomUpdate& U) {
    const GridDev& g = c->g;
    if (g.n_faces) k_geom_update<0><<<nblocks(g.n_faces, 256), 256, 0, c->stream>>>(g, U, g.n_faces);
    if (g.n_cells) k_geom_update<1><<<nblocks(g.n_cells, 256), 256, 0, c->stream>>>(g, U, g.n_cells);
    if (g.n_cells) k_geom_update<2><<<nblocks(g.n_cells, 256), 256, 0, c->stream>>>(g, U, g.n_cells);
    if (g.n_bfaces) k_geom_update<3><<<nblocks(g.n_bfaces, 256), 256, 0, c->stream>>>(g, U, g.n_bfaces);
    CUDA_CHECK_LAUNCH(c);
}
void launch_move_vertices(ibis_b200_ctx* c, double* pos, const double* vel, double scale, int64_t n3) {
    if (n3 == 0) return;
    k_move_vertices<<<nblocks(n3, 256), 256, 0, c->stream>>>(pos, vel, scale, (long)n3);
    CUDA_CHECK_LAUNCH(c);
}

void launch_face_vel(ibis_b200_ctx* c, const int* fverts, const double* vvel, double* face_vel) {
    if (c->g.n_faces == 0) return;
    k_face_vel<<<nblocks(c->g.n_faces, 256), 256, 0, c->stream>>>(c->g, fverts, vvel, face_vel);
    CUDA_CHECK_LAUNCH(c);
}

template <typename S>
void launch_gcl(ibis_b200_ctx* c, const FvArgs<S>& A, S* const* cq) {
    if (A.g.n_owned == 0) return;
    CqPtrs<S> q;
    for (int k = 0; k < 5; k++) q.p[k] = cq[k];
    if (A.g.dim == 3)
        k_gcl<3, S><<<nblocks(A.g.n_owned, 256), 256, 0, c->stream>>>(A, q);
    else
        k_gcl<2, S><<<nblocks(A.g.n_owned, 256), 256, 0, c->stream>>>(A, q);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_gcl<double>(ibis_b200_ctx*, const FvArgs<double>&, double* const*);
template void launch_gcl<dual>(ibis_b200_ctx*, const FvArgs<dual>&, dual* const*);

template <typename S>
void launch_face_viscous(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.dim == 3)
        k_face_viscous<3, S><<<nblocks(A.g.n_faces, 128), 128, 0, c->stream>>>(A);
    else
        k_face_viscous<2, S><<<nblocks(A.g.n_faces, 128), 128, 0, c->stream>>>(A);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_face_viscous<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_face_viscous<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <typename S>
void launch_integral(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.dim == 3)
        k_integral<3, S><<<nblocks(A.g.n_owned, 256), 256, 0, c->stream>>>(A);
    else
        k_integral<2, S><<<nblocks(A.g.n_owned, 256), 256, 0, c->stream>>>(A);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_integral<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_integral<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <typename S>
void launch_local_dt_min(ibis_b200_ctx* c, const FvArgs<S>& A, double* d_out) {
    const int nb = reduction_blocks(A.g.n_owned);
    k_local_dt_min<S><<<nb, 256, 0, c->stream>>>(A, c->d_partials);
    CUDA_CHECK_LAUNCH(c);
    k_final_min<<<1, 256, 0, c->stream>>>(c->d_partials, nb, d_out);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_local_dt_min<double>(ibis_b200_ctx*, const FvArgs<double>&, double*);
template void launch_local_dt_min<dual>(ibis_b200_ctx*, const FvArgs<dual>&, double*);

}  // namespace ib
