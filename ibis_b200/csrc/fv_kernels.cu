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

#ifndef IBIS_GS_BLOCK
#define IBIS_GS_BLOCK 128    // threads per block of pass G' (gradients + limiters + interface sides)
#endif
#ifndef IBIS_GS_MINB
#define IBIS_GS_MINB 4
#endif
#ifndef IBIS_F2_THREADS
#define IBIS_F2_THREADS 256  // threads per block of pass F' (>= tile size)
#endif
#ifndef IBIS_F2_MINB
#define IBIS_F2_MINB 2
#endif

namespace ib {

#define CUDA_CHECK_LAUNCH(c)                                                             \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess && (c)->last_error.empty())                               \
            (c)->last_error = std::string("kernel launch: ") + cudaGetErrorString(e__);  \
        (c)->launches++;                                                                 \
    } while (0)

static inline int nblocks(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

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

// pass G' of the sides path: one thread per cell with a face list (owned + inner halo ring).
// The block's slice of the static tables (WLS weights, face offsets: 36 KB for hexes) arrives in
// shared memory by bulk async copies (TMA) while the threads issue their neighbour gathers, so the
// only registers tied up while waiting on HBM are the gathered values themselves.
template <int DIM, int NF, typename S>
__global__ void __launch_bounds__(IBIS_GS_BLOCK, IBIS_GS_MINB) k_grad_sides(const FvArgs<S> A) {
    constexpr int B = IBIS_GS_BLOCK;
    constexpr int NT = NF * 3;                                  // arrays per table
    __shared__ __align__(128) double tab[2 * NT * B];
    __shared__ uint64_t bar;
    const int c0 = blockIdx.x * B;
    const int c = c0 + threadIdx.x;
    if (threadIdx.x == 0) {
        ac::mbar_init(&bar, 1);
        ac::fence_barrier_init();
        ac::mbar_arrive_expect_tx(&bar, 2 * NT * B * (uint32_t)sizeof(double));
    }
    __syncthreads();
    if (threadIdx.x < 2 * NT) {
        const int a = threadIdx.x % NT;
        const double* src = (threadIdx.x < NT ? A.g.cf_w : A.g.cf_d) + (long)a * A.g.cstride + c0;
        ac::bulk_g2s(tab + threadIdx.x * B, src, B * (uint32_t)sizeof(double), &bar);
    }
    CellGather<NF, S> G;
    const bool active = c < A.g.n_grad;
    if (active) grad_sides_gather<DIM, NF>(A, c, G);
    ac::mbar_wait(&bar, 0);
    if (active) {
        CellTables<B> T;
        T.w = tab;
        T.d = tab + NT * B;
        T.stride_rt = B;
        T.idx = threadIdx.x;
        grad_sides_compute<DIM, NF, B>(A, G, T);
    }
}

template <typename S>
__global__ void __launch_bounds__(256) k_frec_geom(const GridDev g, S* sides) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < g.n_faces) frec_geom_body(g, sides, f);
}

// pass F' of the sides path: blockDim.x >= A.g.tile; dynamic shared memory = tile_cap records.
// The records of the tile's home interfaces are one contiguous range: one bulk async copy (TMA);
// the foreign ones (homed in an earlier tile) are gathered with 16-byte cp.async.  Each flux
// overwrites the left side of its own record; then one thread per cell sums its faces.
template <int DIM, int FLUX, typename S>
__global__ void __launch_bounds__(IBIS_F2_THREADS, IBIS_F2_MINB) k_flux_tile(const FvArgs<S> A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    S* st = reinterpret_cast<S*>(smem_raw);
    constexpr int RS = frec<S>::stride, RB = frec<S>::bytes;
    const GridDev& g = A.g;
    const int tile = blockIdx.x;
    const int h0 = g.home0[tile], nh = g.home0[tile + 1] - h0;
    const int fo = g.foreign_off[tile], n = nh + (g.foreign_off[tile + 1] - fo);
    if (threadIdx.x == 0) {
        ac::mbar_init(&bar, 1);
        ac::fence_barrier_init();
        if (nh > 0) {
            ac::mbar_arrive_expect_tx(&bar, (uint32_t)nh * RB);
            ac::bulk_g2s(st, A.sides + (long)h0 * RS, (uint32_t)nh * RB, &bar);
        }
    }
    for (int s = nh + threadIdx.x; s < n; s += blockDim.x) {
        const int f = g.foreign[fo + (s - nh)];
        const unsigned char* src = reinterpret_cast<const unsigned char*>(A.sides + (long)f * RS);
        unsigned char* dst = reinterpret_cast<unsigned char*>(st + (long)s * RS);
#pragma unroll
        for (int q = 0; q < RB / 16; q++) ac::cp_async_16(dst + q * 16, src + q * 16);
    }
    const int c = tile * g.tile + threadIdx.x;
    const bool has_cell = threadIdx.x < g.tile && c < g.n_owned;
    CellFaces CF;
    if (has_cell) tile_cell_load(g, c, CF);
    ac::cp_async_wait_all();
    __syncthreads();                       // barrier initialised; everybody's cp.async data landed
    if (nh > 0) ac::mbar_wait(&bar, 0);
    for (int s = threadIdx.x; s < n; s += blockDim.x) {
        S* rec = st + (long)s * RS;
        flux_from_record<DIM, FLUX>(A, rec, rec, 1);
    }
    __syncthreads();
    if (has_cell) tile_cell_sum<DIM>(A, c, CF, st, RS, 1);
}

template <int DIM, int FLUX, bool O2, typename S>
__global__ void __launch_bounds__(128) k_face_flux(const FvArgs<S> A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < A.g.n_faces) face_flux_body<DIM, FLUX, O2>(A, f);
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

template <int DIM, typename S>
static void launch_grad_sides_nf(ibis_b200_ctx* c, const FvArgs<S>& A) {
    const int nb = nblocks(A.g.n_grad, IBIS_GS_BLOCK);
    switch (A.g.max_faces) {
        case 6: k_grad_sides<DIM, 6, S><<<nb, IBIS_GS_BLOCK, 0, c->stream>>>(A); break;
        case 5: k_grad_sides<DIM, 5, S><<<nb, IBIS_GS_BLOCK, 0, c->stream>>>(A); break;
        case 4: k_grad_sides<DIM, 4, S><<<nb, IBIS_GS_BLOCK, 0, c->stream>>>(A); break;
        default: k_grad_sides<DIM, 3, S><<<nb, IBIS_GS_BLOCK, 0, c->stream>>>(A); break;
    }
}

template <typename S>
void launch_grad_sides(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_grad == 0) return;
    prof_begin(c, 0);
    if (A.g.dim == 3) launch_grad_sides_nf<3, S>(c, A);
    else launch_grad_sides_nf<2, S>(c, A);
    prof_end(c, 0);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_grad_sides<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_grad_sides<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <int DIM, typename S>
static void launch_flux_tile_dim(ibis_b200_ctx* c, const FvArgs<S>& A, int bs, size_t smem) {
    switch (c->cfg.flux_calculator) {
        case FLUX_HANEL: launch_one(c, k_flux_tile<DIM, FLUX_HANEL, S>, A.g.n_tiles, bs, smem, A); break;
        case FLUX_AUSMDV: launch_one(c, k_flux_tile<DIM, FLUX_AUSMDV, S>, A.g.n_tiles, bs, smem, A); break;
        case FLUX_LDFSS2: launch_one(c, k_flux_tile<DIM, FLUX_LDFSS2, S>, A.g.n_tiles, bs, smem, A); break;
        default: launch_one(c, k_flux_tile<DIM, FLUX_RUSANOV, S>, A.g.n_tiles, bs, smem, A); break;
    }
}

template <typename S>
void launch_flux_tile(ibis_b200_ctx* c, const FvArgs<S>& A) {
    if (A.g.n_tiles == 0) return;
    const size_t smem = (size_t)A.g.tile_cap * frec<S>::bytes;
    const int bs = A.g.tile > IBIS_F2_THREADS ? A.g.tile : IBIS_F2_THREADS;
    prof_begin(c, 1);
    if (A.g.dim == 3) launch_flux_tile_dim<3, S>(c, A, bs, smem);
    else launch_flux_tile_dim<2, S>(c, A, bs, smem);
    prof_end(c, 1);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_flux_tile<double>(ibis_b200_ctx*, const FvArgs<double>&);
template void launch_flux_tile<dual>(ibis_b200_ctx*, const FvArgs<dual>&);

template <typename S>
void launch_frec_geom(ibis_b200_ctx* c, S* sides) {
    if (c->g.n_faces == 0) return;
    k_frec_geom<S><<<nblocks(c->g.n_faces, 256), 256, 0, c->stream>>>(c->g, sides);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_frec_geom<double>(ibis_b200_ctx*, double*);
template void launch_frec_geom<dual>(ibis_b200_ctx*, dual*);

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
