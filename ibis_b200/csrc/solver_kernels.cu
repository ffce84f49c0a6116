// solver_kernels.cu -- state-vector kernels of the explicit and steady solvers:
// primitive/conserved conversion, Runge-Kutta axpy, L2 norms, bad-cell count, layout
// conversion at the ABI boundary, BLAS-1 on device vectors with two-stage deterministic
// reductions (warp shuffles, then one block over the per-block partials), Krylov-basis
// gemv, JFNK pack/unpack, halo pack/unpack.
#include <cuda_runtime.h>

#include "context.h"
#include "solver_bodies.h"

namespace ib {

#define CUDA_CHECK_LAUNCH(c)                                                             \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess && (c)->last_error.empty())                               \
            (c)->last_error = std::string("kernel launch: ") + cudaGetErrorString(e__);  \
        (c)->launches++;                                                                 \
    } while (0)

static inline int nblocks(int64_t n, int bs) { return n > 0 ? (int)((n + bs - 1) / bs) : 1; }   // an empty grid launches one idle block

// grid for reductions: a fixed multiple of the 148 SMs, grid-stride loops
int reduction_blocks(int64_t n) {
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    return (int)nb;
}

template <typename S>
struct Comps7 { S* p[7]; };
template <typename S>
struct Comps5 { S* p[5]; };

template <typename S> static Comps7<S> c7(S* const* a) { Comps7<S> r; for (int i = 0; i < 7; i++) r.p[i] = a[i]; return r; }
template <typename S> static Comps5<S> c5(S* const* a) { Comps5<S> r; for (int i = 0; i < 5; i++) r.p[i] = a[i]; return r; }

template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_cons_to_prim(Comps5<S> cq, Comps7<S> fs, Gas<S> gas, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cons_to_prim_body<DIM>(cq.p, fs.p, gas, i);
}

template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_prim_to_cons(Comps5<S> cq, Comps7<S> fs, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prim_to_cons_body<DIM>(cq.p, fs.p, i);
}

// dt_eff = d_dt ? scale * d_dt[0] : dt   (the device-resident time step of rk_step)
template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_apply_dt(Comps5<S> u0, Comps5<S> u1, Comps5<S> k, double dt,
                                                   const double* d_dt, double scale, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const double dte = d_dt ? scale * d_dt[0] : dt;
    if (i < n) apply_dt_body<DIM>(u0.p, u1.p, k.p, dte, i);
}

template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_rk_update(Comps5<S> u_in, RkTerms<S> T, const double* d_dt, Comps5<S> u_out,
                                                    Comps7<S> fs, Gas<S> gas, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rk_update_body<DIM>(u_in.p, T, d_dt[0], u_out.p, fs.p, gas, i);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NQ quantities held per thread; result valid in thread 0
template <int NQ>
__device__ __forceinline__ void block_sum(double* v, double* sm /* [NQ*8] */) {
#pragma unroll
    for (int q = 0; q < NQ; q++) v[q] = warp_sum(v[q]);
    if ((threadIdx.x & 31) == 0)
        for (int q = 0; q < NQ; q++) sm[q * 8 + (threadIdx.x >> 5)] = v[q];
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[q * 8 + threadIdx.x] : 0.0;
            v[q] = warp_sum(t);
        }
    }
}

template <int DIM, typename S>
__global__ void __launch_bounds__(256) k_l2_partial(Comps5<S> cq, long n, double* partial) {
    __shared__ double sm[6 * 8];
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double t[6];
        l2_body<DIM>(cq.p, i, t);
#pragma unroll
        for (int q = 0; q < 6; q++) acc[q] += t[q];
    }
    block_sum<6>(acc, sm);
    if (threadIdx.x == 0)
        for (int q = 0; q < 6; q++) partial[blockIdx.x * 8 + q] = acc[q];
}

// sums NQ columns of partial[nb][8]; post: 0 none, 1 sqrt of each, 2 sqrt and out[1] = 1/sqrt
template <int NQ>
__global__ void __launch_bounds__(256) k_final_sum(const double* partial, int nb, double* out, int post) {
    __shared__ double sm[NQ * 8];
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) acc[q] = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
#pragma unroll
        for (int q = 0; q < NQ; q++) acc[q] += partial[i * 8 + q];
    block_sum<NQ>(acc, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) out[q] = post ? sqrt(acc[q]) : acc[q];
        if (post == 2) out[1] = 1.0 / out[0];
    }
}

template <typename S>
__global__ void __launch_bounds__(256) k_count_bad_partial(Comps7<S> fs, long n, double* partial) {
    __shared__ double sm[8];
    double acc[1] = {0.0};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        if (bad_cell(mem<S>::load(fs.p[F_T], i), mem<S>::load(fs.p[F_RHO], i))) acc[0] += 1.0;
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x * 8] = acc[0];
}

// RungeKutta::estimate_dt (runge_kutta.cpp:108-119) without plot clamping
__global__ void k_rk_dt(const double* d_dt_stable, double cfl, double dt_prev, double t_remaining,
                        double* d_dt) {
    const double a = cfl * d_dt_stable[0];
    // dt_prev = NaN: the step the device took last (d_dt still holds it; DBL_MAX before the first step)
    const double b = 1.5 * (dt_prev != dt_prev ? d_dt[0] : dt_prev);
    const double dt_startup = (b < a) ? b : a;                 // Ibis::min(a, b)
    d_dt[0] = (t_remaining < dt_startup) ? t_remaining : dt_startup;
}

// ---- ABI layout conversion: rows of ncomp_ref doubles (AoS) <-> internal components ----
struct CompMap { int m[8]; };

template <typename S>
__global__ void __launch_bounds__(256) k_pack_in(const double* re_, const double* du_, Comps7<S> comps,
                                                  CompMap map, int ncomp_ref, long n, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;     // row in reference numbering
    if (i >= n) return;
    const long p = perm ? perm[i] : i;                              // internal cell id
    for (int q = 0; q < ncomp_ref; q++) {
        const double r = re_[i * ncomp_ref + q];
        const double d = du_ ? du_[i * ncomp_ref + q] : 0.0;
        S v;
        if constexpr (sizeof(S) == sizeof(double)) v = r; else v = dual(r, d);
        mem<S>::store(comps.p[map.m[q]], p, v);
    }
}

template <typename S>
__global__ void __launch_bounds__(256) k_flow_read(const double* stage, long n, int has_vz, Comps7<S> fs, Gas<S> gas,
                                                    const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flow_read_body<S>(stage, stage + n, stage + 2 * n, stage + 3 * n, has_vz ? stage + 4 * n : nullptr, fs.p, gas, i,
                      perm ? perm[i] : i);
}
template <typename S>
__global__ void __launch_bounds__(256) k_flow_write(Comps7<S> fs, double* stage, long n, int has_vz, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flow_write_body<S>(fs.p, stage, stage + n, stage + 2 * n, stage + 3 * n, has_vz ? stage + 4 * n : nullptr, i,
                       perm ? perm[i] : i);
}

template <typename S>
__global__ void __launch_bounds__(256) k_pack_out(Comps7<S> comps, CompMap map, int ncomp_ref, long n,
                                                   double* re_, double* du_, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
    for (int q = 0; q < ncomp_ref; q++) {
        const S v = mem<S>::load(comps.p[map.m[q]], p);
        if (re_) re_[i * ncomp_ref + q] = re(v);
        if (du_) du_[i * ncomp_ref + q] = du(v);
    }
}

__global__ void __launch_bounds__(256) k_gather_debug(const double* src, long stride, int ncomp, long n,
                                                       double* out, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
    for (int q = 0; q < ncomp; q++) out[i * ncomp + q] = src[q * stride + p];
}

// ---- vectors ----
__global__ void __launch_bounds__(256) k_dot_partial(const double* a, const double* b, long n, double* partial) {
    __shared__ double sm[8];
    double acc[1] = {0.0};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        acc[0] += a[i] * b[i];
    block_sum<1>(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x * 8] = acc[0];
}

// y += x * (sign * s)   with s from the host or from device memory
__global__ void __launch_bounds__(256) k_axpy(double* y, const double* x, double s, const double* d_s,
                                               double sign, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const double f = sign * (d_s ? d_s[0] : s);
    if (i < n) y[i] += x[i] * f;
}

__global__ void __launch_bounds__(256) k_scale(const double* x, double* r, double s, const double* d_s, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const double f = d_s ? d_s[0] : s;
    if (i < n) r[i] = x[i] * f;
}

__global__ void __launch_bounds__(256) k_sub(const double* a, const double* b, double* r, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = a[i] - b[i];
}

__global__ void __launch_bounds__(256) k_fill(double* p, double v, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// out(row) = sum_col basis(row, col) * y(col), columns in order (dense_linear_algebra.h:244-262)
__global__ void __launch_bounds__(256) k_gemv(const double* basis, long n, long stride, int m,
                                               const double* y, double* out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double dot = 0.0;
    for (int col = 0; col < m; col++) dot += basis[col * stride + i] * y[col];
    out[i] = dot;
}

// ---- JFNK (solvers/steady_state.cpp:60-68,102-110,142-149; solvers/jfnk.cpp:69-76) ----
template <int NC>
__global__ void __launch_bounds__(256) k_jfnk_set_dual(Comps5<dual> cq, const double* vec, Comps5<dual> tmp, long n,
                                                        const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // cell in reference numbering
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
#pragma unroll
    for (int q = 0; q < NC; q++) {
        const int k = (NC == 4 && q == 3) ? 4 : q;            // 2-D: mass mx my energy
        const dual v = mem<dual>::load(cq.p[k], p);
        mem<dual>::store(tmp.p[k], p, dual(v.re, vec[i * NC + q]));
    }
}

template <int NC>
__global__ void __launch_bounds__(256) k_jfnk_set_vector(Comps5<dual> res, const double* vec, double dt_star,
                                                          double* result, long n, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
#pragma unroll
    for (int q = 0; q < NC; q++) {
        const int k = (NC == 4 && q == 3) ? 4 : q;
        result[i * NC + q] = 1 / dt_star * vec[i * NC + q] - mem<dual>::load(res.p[k], p).du;
    }
}

template <int NC>
__global__ void __launch_bounds__(256) k_jfnk_eval_rhs(Comps5<dual> res, double* rhs, long n, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
#pragma unroll
    for (int q = 0; q < NC; q++) {
        const int k = (NC == 4 && q == 3) ? 4 : q;
        rhs[i * NC + q] = mem<dual>::load(res.p[k], p).re;
    }
}

template <int NC>
__global__ void __launch_bounds__(256) k_jfnk_apply_update(Comps5<dual> cq, const double* dU, long n, const int* perm) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long p = perm ? perm[i] : i;
#pragma unroll
    for (int q = 0; q < NC; q++) {
        const int k = (NC == 4 && q == 3) ? 4 : q;
        const dual v = mem<dual>::load(cq.p[k], p);
        mem<dual>::store(cq.p[k], p, dual(v.re + dU[i * NC + q], 0.0));
    }
}

// ---- halo pack / unpack: buf[k][var] <-> fs[var][cells[k]] (per-peer segments are contiguous) ----
template <typename S>
__global__ void __launch_bounds__(256) k_halo_pack(Comps7<S> fs, const int* cells, long n, S* buf) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cidx = cells[i];
#pragma unroll
    for (int q = 0; q < 7; q++) mem<S>::store(buf, i * 7 + q, mem<S>::load(fs.p[q], cidx));
}

template <typename S>
__global__ void __launch_bounds__(256) k_halo_unpack(Comps7<S> fs, const int* cells, long n, const S* buf) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cidx = cells[i];
#pragma unroll
    for (int q = 0; q < 7; q++) mem<S>::store(fs.p[q], cidx, mem<S>::load(buf, i * 7 + q));
}

// ---- peer-memory halo exchange (NVLink P2P, no NCCL on the data path) ----
// push: one thread per sent cell gathers its 7 rows and stores them, coalesced along the cell
// index, into the PEER's receive area; every block fences its remote stores at system scope and the
// last block to finish publishes the exchange number to the peer's flag (release).
template <typename S>
__global__ void __launch_bounds__(256) k_halo_push(Comps7<S> fs, const int* cells, long n, S* remote,
                                                   unsigned long long* remote_flag, unsigned* counter,
                                                   unsigned long long iter) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int cidx = cells[i];
#pragma unroll
        for (int q = 0; q < 7; q++) mem<S>::store(remote, q * n + i, mem<S>::load(fs.p[q], cidx));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0u;                                   // ready for the next exchange
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote_flag), "l"(iter) : "memory");
        }
    }
}

// pull: wait until the peer has published this exchange (acquire), then scatter the segment; the
// area is read past L1 (ld.global.cg): it was written by another GPU.
// The wait is bounded (about 10 s of SM clocks): a peer that died, returned early on an error or made a different
// number of evaluations never publishes -- the block then records the stall in the context's abort word (the one the
// pipelined kernels use; abi.cu::check_pipe_stall turns it into ECOMM at the next synchronise) and leaves the halo as it is.
constexpr long long HALO_TIMEOUT_CLOCKS = 20000000000ll;
template <typename S>
__global__ void __launch_bounds__(256) k_halo_pull(Comps7<S> fs, const int* cells, long n, const S* area,
                                                   const unsigned long long* flag, unsigned long long iter, unsigned int* dbg) {
    __shared__ int gave_up;
    if (threadIdx.x == 0) {
        unsigned long long seen;
        const long long t0 = clock64();
        int stalled = 0;
        for (unsigned it = 1;; it++) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
            if (seen >= iter) break;
            if ((it & 1023u) == 0u && (*reinterpret_cast<volatile unsigned int*>(dbg) != 0u || clock64() - t0 > HALO_TIMEOUT_CLOCKS)) {
                if (atomicCAS(dbg, 0u, 1u) == 0u) {
                    dbg[1] = 64u; dbg[2] = blockIdx.x; dbg[3] = 0u; dbg[4] = (unsigned)iter; dbg[5] = (unsigned)seen;
                    __threadfence_system();
                }
                stalled = 1;
                break;
            }
        }
        gave_up = stalled;
    }
    __syncthreads();
    if (gave_up) return;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cidx = cells[i];
#pragma unroll
    for (int q = 0; q < 7; q++) {
        const double* src = reinterpret_cast<const double*>(area + (q * n + i));
        if (sizeof(S) == sizeof(double)) {
            reinterpret_cast<double*>(fs.p[q])[cidx] = __ldcg(src);
        } else {
            const double2 t = __ldcg(reinterpret_cast<const double2*>(src));
            *reinterpret_cast<double2*>(fs.p[q] + cidx) = t;
        }
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
template <typename S> static Gas<S> gas_of(const ibis_b200_ctx* c) {
    Gas<S> g;
    g.R = S(c->cfg.R); g.Cv = S(c->cfg.Cv); g.Cp = S(c->cfg.Cp); g.gamma = S(c->cfg.gamma);
    return g;
}

template <typename S>
void launch_cons_to_prim(ibis_b200_ctx* c, S* const* cq, S* const* fs, int64_t n) {
    if (c->dim == 3) k_cons_to_prim<3, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(cq), c7(fs), gas_of<S>(c), n);
    else k_cons_to_prim<2, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(cq), c7(fs), gas_of<S>(c), n);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_cons_to_prim<double>(ibis_b200_ctx*, double* const*, double* const*, int64_t);
template void launch_cons_to_prim<dual>(ibis_b200_ctx*, dual* const*, dual* const*, int64_t);

template <typename S>
void launch_prim_to_cons(ibis_b200_ctx* c, S* const* cq, S* const* fs, int64_t n) {
    if (c->dim == 3) k_prim_to_cons<3, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(cq), c7(fs), n);
    else k_prim_to_cons<2, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(cq), c7(fs), n);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_prim_to_cons<double>(ibis_b200_ctx*, double* const*, double* const*, int64_t);
template void launch_prim_to_cons<dual>(ibis_b200_ctx*, dual* const*, dual* const*, int64_t);

template <typename S>
void launch_apply_dt(ibis_b200_ctx* c, S* const* u0, S* const* u1, S* const* k, double dt,
                     const double* d_dt, double scale, int64_t n) {
    if (c->dim == 3) k_apply_dt<3, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(u0), c5(u1), c5(k), dt, d_dt, scale, n);
    else k_apply_dt<2, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(u0), c5(u1), c5(k), dt, d_dt, scale, n);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_apply_dt<double>(ibis_b200_ctx*, double* const*, double* const*, double* const*, double, const double*, double, int64_t);
template void launch_apply_dt<dual>(ibis_b200_ctx*, dual* const*, dual* const*, dual* const*, double, const double*, double, int64_t);

template <typename S>
void launch_rk_update(ibis_b200_ctx* c, S* const* u_in, const RkTerms<S>& T, const double* d_dt, S* const* u_out,
                      S* const* fs, int64_t n) {
    Comps5<S> uo;
    for (int q = 0; q < 5; q++) uo.p[q] = u_out ? u_out[q] : nullptr;
    if (c->dim == 3) k_rk_update<3, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(u_in), T, d_dt, uo, c7(fs), gas_of<S>(c), n);
    else k_rk_update<2, S><<<nblocks(n, 256), 256, 0, c->stream>>>(c5(u_in), T, d_dt, uo, c7(fs), gas_of<S>(c), n);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_rk_update<double>(ibis_b200_ctx*, double* const*, const RkTerms<double>&, const double*, double* const*, double* const*, int64_t);

template <typename S>
void launch_l2_norms(ibis_b200_ctx* c, S* const* cq, int64_t n, double* d_out6) {
    const int nb = reduction_blocks(n);
    if (c->dim == 3) k_l2_partial<3, S><<<nb, 256, 0, c->stream>>>(c5(cq), n, c->d_partials);
    else k_l2_partial<2, S><<<nb, 256, 0, c->stream>>>(c5(cq), n, c->d_partials);
    CUDA_CHECK_LAUNCH(c);
    // sums only; the square roots are taken after the cross-rank reduction
    k_final_sum<6><<<1, 256, 0, c->stream>>>(c->d_partials, nb, d_out6, 0);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_l2_norms<double>(ibis_b200_ctx*, double* const*, int64_t, double*);
template void launch_l2_norms<dual>(ibis_b200_ctx*, dual* const*, int64_t, double*);

template <typename S>
void launch_count_bad(ibis_b200_ctx* c, S* const* fs, int64_t n, double* d_out) {
    const int nb = reduction_blocks(n);
    k_count_bad_partial<S><<<nb, 256, 0, c->stream>>>(c7(fs), n, c->d_partials);
    CUDA_CHECK_LAUNCH(c);
    k_final_sum<1><<<1, 256, 0, c->stream>>>(c->d_partials, nb, d_out, 0);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_count_bad<double>(ibis_b200_ctx*, double* const*, int64_t, double*);
template void launch_count_bad<dual>(ibis_b200_ctx*, dual* const*, int64_t, double*);

void launch_rk_dt(ibis_b200_ctx* c, const double* d_dt_stable, double cfl, double dt_prev,
                  double t_remaining, double* d_dt) {
    k_rk_dt<<<1, 1, 0, c->stream>>>(d_dt_stable, cfl, dt_prev, t_remaining, d_dt);
    CUDA_CHECK_LAUNCH(c);
}

template <typename S>
void launch_pack_in(ibis_b200_ctx* c, const double* d_re, const double* d_du, S* const* comps,
                    const int* comp_map, int ncomp_ref, int64_t n) {
    Comps7<S> cc;
    for (int i = 0; i < 7; i++) cc.p[i] = comps[i];
    CompMap m;
    for (int i = 0; i < 8; i++) m.m[i] = i < ncomp_ref ? comp_map[i] : 0;
    k_pack_in<S><<<nblocks(n, 256), 256, 0, c->stream>>>(d_re, d_du, cc, m, ncomp_ref, n, c->d_ref2int);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_pack_in<double>(ibis_b200_ctx*, const double*, const double*, double* const*, const int*, int, int64_t);
template void launch_pack_in<dual>(ibis_b200_ctx*, const double*, const double*, dual* const*, const int*, int, int64_t);

template <typename S>
void launch_flow_read(ibis_b200_ctx* c, const double* stage, int64_t n, int has_vz, S* const* fs) {
    if (n <= 0) return;
    k_flow_read<S><<<nblocks(n, 256), 256, 0, c->stream>>>(stage, n, has_vz, c7(fs), gas_of<S>(c), c->d_ref2int);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_flow_read<double>(ibis_b200_ctx*, const double*, int64_t, int, double* const*);
template void launch_flow_read<dual>(ibis_b200_ctx*, const double*, int64_t, int, dual* const*);
template <typename S>
void launch_flow_write(ibis_b200_ctx* c, S* const* fs, double* stage, int64_t n, int has_vz) {
    if (n <= 0) return;
    k_flow_write<S><<<nblocks(n, 256), 256, 0, c->stream>>>(c7(fs), stage, n, has_vz, c->d_ref2int);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_flow_write<double>(ibis_b200_ctx*, double* const*, double*, int64_t, int);
template void launch_flow_write<dual>(ibis_b200_ctx*, dual* const*, double*, int64_t, int);

template <typename S>
void launch_pack_out(ibis_b200_ctx* c, S* const* comps, const int* comp_map, int ncomp_ref,
                     int64_t n, double* d_re, double* d_du) {
    Comps7<S> cc;
    for (int i = 0; i < 7; i++) cc.p[i] = comps[i];
    CompMap m;
    for (int i = 0; i < 8; i++) m.m[i] = i < ncomp_ref ? comp_map[i] : 0;
    k_pack_out<S><<<nblocks(n, 256), 256, 0, c->stream>>>(cc, m, ncomp_ref, n, d_re, d_du, c->d_ref2int);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_pack_out<double>(ibis_b200_ctx*, double* const*, const int*, int, int64_t, double*, double*);
template void launch_pack_out<dual>(ibis_b200_ctx*, dual* const*, const int*, int, int64_t, double*, double*);

__global__ void k_dual_kat(const double* in, double* out) { dual_kat_body(in, out); }
// operands for the division self-test: a 64-bit mix, shaped so that every exponent (denormals, zeros, infinities and
// NaNs included) and both signs turn up, and so that equal and nearly equal operands are frequent
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double test_operand(unsigned long long r) {
    const unsigned kind = (unsigned)(r >> 60);
    if (kind == 0) return __longlong_as_double((long long)(r & 0x8000000000000000ull));             // +-0
    if (kind == 1) return __longlong_as_double((long long)(r & 0x800fffffffffffffull));              // denormal
    if (kind == 2) return __longlong_as_double((long long)((r & 0x8000000000000000ull) | 0x7ff0000000000000ull | ((r & 1) ? (r & 0xfffff) : 0)));
    if (kind < 8) {                                                                                  // moderate magnitudes
        const unsigned long long e = 1023 - 40 + ((r >> 52) & 0x7f) % 80;
        return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | (e << 52)));
    }
    return __longlong_as_double((long long)(r & ~(0xfull << 60) | ((r >> 4) & (0x7ull << 60))));    // any exponent
}
__device__ __forceinline__ bool same_bits(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b) || (a != a && b != b);
}
__global__ void __launch_bounds__(256) k_division_check(long long n, unsigned long long seed, unsigned long long* bad) {
    unsigned long long local = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long r = mix64(seed + 4 * (unsigned long long)i);
        const double a1 = test_operand(r), b1 = test_operand(mix64(r + 1)), a2 = test_operand(mix64(r + 2));
        double b2 = test_operand(mix64(r + 3));
        if ((r & 0x300) == 0) b2 = a2;                       // equal operands
        double q1, q2;
        div_pair(a1, b1, a2, b2, q1, q2);
        local += !same_bits(q1, a1 / b1) + !same_bits(q2, a2 / b2);
        div_shared(a1, a2, b1, q1, q2);
        local += !same_bits(q1, a1 / b1) + !same_bits(q2, a2 / b1);
        local += !same_bits(div_z(a1, b2), a1 / b2);
    }
    if (local) atomicAdd(bad, local);
}
__global__ void k_u64_to_double(unsigned long long* p) { *reinterpret_cast<double*>(p) = (double)*p; }
void launch_division_check(ibis_b200_ctx* c, int64_t n, uint64_t seed, double* d_out) {
    unsigned long long* bad = reinterpret_cast<unsigned long long*>(d_out);
    cudaMemsetAsync(bad, 0, 8, c->stream);
    k_division_check<<<148 * 8, 256, 0, c->stream>>>((long long)n, (unsigned long long)seed, bad);
    k_u64_to_double<<<1, 1, 0, c->stream>>>(bad);
    CUDA_CHECK_LAUNCH(c);
}

void launch_dual_kat(ibis_b200_ctx* c, const double* in, double* out) {
    k_dual_kat<<<1, 1, 0, c->stream>>>(in, out);
    CUDA_CHECK_LAUNCH(c);
}

void launch_gather_debug(ibis_b200_ctx* c, const double* src, int ncomp, int64_t n, double* out) {
    k_gather_debug<<<nblocks(n, 256), 256, 0, c->stream>>>(src, c->cstride, ncomp, n, out, c->d_ref2int);
    CUDA_CHECK_LAUNCH(c);
}

__global__ void k_dot_ordered(const double* a, const double* b, long n, double* out, int post) {
    double acc = 0.0;
    for (long i = 0; i < n; i++) acc += a[i] * b[i];
    out[0] = post ? sqrt(acc) : acc;
    if (post == 2) out[1] = 1.0 / out[0];
}

void launch_dot(ibis_b200_ctx* c, const double* a, const double* b, int64_t n, double* d_out, int post) {
    if (c->ordered_dot && c->n_ranks <= 1) {
        k_dot_ordered<<<1, 1, 0, c->stream>>>(a, b, n, d_out, post);
        CUDA_CHECK_LAUNCH(c);
        return;
    }
    const int nb = reduction_blocks(n);
    k_dot_partial<<<nb, 256, 0, c->stream>>>(a, b, n, c->d_partials);
    CUDA_CHECK_LAUNCH(c);
    k_final_sum<1><<<1, 256, 0, c->stream>>>(c->d_partials, nb, d_out, post);
    CUDA_CHECK_LAUNCH(c);
}

__global__ void k_scalar_post(double* out, int post) {
    out[0] = sqrt(out[0]);
    if (post == 2) out[1] = 1.0 / out[0];
}
void launch_scalar_post(ibis_b200_ctx* c, double* d_out, int post) {
    if (!post) return;
    k_scalar_post<<<1, 1, 0, c->stream>>>(d_out, post);
    CUDA_CHECK_LAUNCH(c);
}

void launch_axpy(ibis_b200_ctx* c, double* y, const double* x, double s, const double* d_s,
                 double sign, int64_t n) {
    k_axpy<<<nblocks(n, 256), 256, 0, c->stream>>>(y, x, s, d_s, sign, n);
    CUDA_CHECK_LAUNCH(c);
}

void launch_scale(ibis_b200_ctx* c, const double* x, double* r, double s, const double* d_s, int64_t n) {
    k_scale<<<nblocks(n, 256), 256, 0, c->stream>>>(x, r, s, d_s, n);
    CUDA_CHECK_LAUNCH(c);
}

void launch_sub(ibis_b200_ctx* c, const double* a, const double* b, double* r, int64_t n) {
    k_sub<<<nblocks(n, 256), 256, 0, c->stream>>>(a, b, r, n);
    CUDA_CHECK_LAUNCH(c);
}

void launch_gemv(ibis_b200_ctx* c, const double* basis, int64_t n, int64_t stride, int m,
                 const double* d_y, double* out) {
    k_gemv<<<nblocks(n, 256), 256, 0, c->stream>>>(basis, n, stride, m, d_y, out);
    CUDA_CHECK_LAUNCH(c);
}

void launch_fill(ibis_b200_ctx* c, double* p, double v, int64_t n) {
    if (n <= 0) return;
    k_fill<<<nblocks(n, 256), 256, 0, c->stream>>>(p, v, n);
    CUDA_CHECK_LAUNCH(c);
}

#define JFNK_DISPATCH(KERNEL, ...)                                                          \
    if (c->dim == 3) KERNEL<5><<<nblocks(n, 256), 256, 0, c->stream>>>(__VA_ARGS__);        \
    else KERNEL<4><<<nblocks(n, 256), 256, 0, c->stream>>>(__VA_ARGS__);                    \
    CUDA_CHECK_LAUNCH(c);

void launch_jfnk_set_dual(ibis_b200_ctx* c, dual* const* cq, const double* vec, dual* const* cq_tmp) {
    const long n = c->n_cells;
    JFNK_DISPATCH(k_jfnk_set_dual, c5(cq), vec, c5(cq_tmp), n, c->d_ref2int)
}
void launch_jfnk_set_vector(ibis_b200_ctx* c, dual* const* res, const double* vec, double dt_star, double* result) {
    const long n = c->n_cells;
    JFNK_DISPATCH(k_jfnk_set_vector, c5(res), vec, dt_star, result, n, c->d_ref2int)
}
void launch_jfnk_eval_rhs(ibis_b200_ctx* c, dual* const* res, double* rhs) {
    const long n = c->n_cells;
    JFNK_DISPATCH(k_jfnk_eval_rhs, c5(res), rhs, n, c->d_ref2int)
}
void launch_jfnk_apply_update(ibis_b200_ctx* c, dual* const* cq, const double* dU) {
    const long n = c->n_cells;
    JFNK_DISPATCH(k_jfnk_apply_update, c5(cq), dU, n, c->d_ref2int)
}

template <typename S>
void launch_halo_pack(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, S* buf) {
    if (n <= 0) return;
    k_halo_pack<S><<<nblocks(n, 256), 256, 0, c->stream>>>(c7(fs), cells, n, buf);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_halo_pack<double>(ibis_b200_ctx*, double* const*, const int*, int64_t, double*);
template void launch_halo_pack<dual>(ibis_b200_ctx*, dual* const*, const int*, int64_t, dual*);

template <typename S>
void launch_halo_unpack(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, const S* buf) {
    if (n <= 0) return;
    k_halo_unpack<S><<<nblocks(n, 256), 256, 0, c->stream>>>(c7(fs), cells, n, buf);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_halo_unpack<double>(ibis_b200_ctx*, double* const*, const int*, int64_t, const double*);
template void launch_halo_unpack<dual>(ibis_b200_ctx*, dual* const*, const int*, int64_t, const dual*);

template <typename S>
void launch_halo_push(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, S* remote,
                      unsigned long long* remote_flag, unsigned* counter, unsigned long long iter) {
    // an empty segment still publishes the exchange number (one block, no cells)
    k_halo_push<S><<<n > 0 ? nblocks(n, 256) : 1, 256, 0, c->stream>>>(c7(fs), cells, n, remote, remote_flag, counter, iter);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_halo_push<double>(ibis_b200_ctx*, double* const*, const int*, int64_t, double*, unsigned long long*, unsigned*, unsigned long long);
template void launch_halo_push<dual>(ibis_b200_ctx*, dual* const*, const int*, int64_t, dual*, unsigned long long*, unsigned*, unsigned long long);

template <typename S>
void launch_halo_pull(ibis_b200_ctx* c, S* const* fs, const int* cells, int64_t n, const S* area,
                      const unsigned long long* flag, unsigned long long iter) {
    k_halo_pull<S><<<n > 0 ? nblocks(n, 256) : 1, 256, 0, c->stream>>>(c7(fs), cells, n, area, flag, iter, c->d_pipe_dbg);
    CUDA_CHECK_LAUNCH(c);
}
template void launch_halo_pull<double>(ibis_b200_ctx*, double* const*, const int*, int64_t, const double*, const unsigned long long*, unsigned long long);
template void launch_halo_pull<dual>(ibis_b200_ctx*, dual* const*, const int*, int64_t, const dual*, const unsigned long long*, unsigned long long);

}  // namespace ib
