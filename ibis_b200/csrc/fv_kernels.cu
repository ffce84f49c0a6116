// fv_kernels.cu -- sm_100a kernels of the finite-volume residual:
//   k_bc            ghost-cell fills of every boundary tag in one launch
//   k_wls_r         least-squares r-factors (setup)
//   k_grad_limit    pass G: all WLS gradients + Barth-Jespersen limiters of a cell, fused
//   k_residual_tile pass F: reconstruct + rotate + flux + rotate back + surface integral per tile of
//                   cells, owner-computes per cell (no face temporaries, no atomics)
//   k_face_flux / k_face_viscous / k_integral   general (viscous) path through a face array
//   k_local_dt_min  local time step + block min
// One thread per cell / face, SoA loads coalesced along the cell index, neighbour gathers
// served by L1/L2.  Built with --fmad=false: arithmetic is bit-for-bit the reference's.
// These are the kernels of the general ("gather") path -- first order, viscous, flux boundary actions, moving
// grids.  The inviscid second-order evaluation runs the persistent pipelined kernels of fv_pipe.cu.
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
