// abi.cu -- the extern "C" surface of libibis_b200.so (include/ibis_b200.h): context and
// buffer management, conversion of the reference GridBlock arrays into the device SoA
// layout, and the orchestration of one residual evaluation / one Runge-Kutta step on the
// context's stream.
#include <math.h>
#include <stdio.h>
#include <time.h>

#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "comm.h"
#include "context.h"
#include <limits>

#include "layout.h"

using namespace ib;

namespace {

thread_local std::string g_last_error;

// Argument errors (EINVAL, EHANDLE) are the caller's and leave the context usable: they are reported by the call
// that made them and by ibis_b200_last_error, nothing else.  Device, memory and communication failures are sticky:
// last_error stays set and every later call that launches work returns ECUDA (the stream's state is unknown).
int fail(ibis_b200_ctx* c, int code, const std::string& msg) {
    if (c) {
        c->last_message = msg;
        if (code == IBIS_B200_ECUDA || code == IBIS_B200_ENOMEM || code == IBIS_B200_ECOMM) c->last_error = msg;
    }
    g_last_error = msg;
    return code;
}

#define GPU_TRY(c, expr)                                                                  \
    do {                                                                                  \
        int e__ = (expr);                                                                 \
        if (e__ != 0)                                                                     \
            return fail(c, IBIS_B200_ECUDA, std::string(#expr) + ": " + gpu::error_string(e__)); \
    } while (0)

}  // namespace
namespace ib {
int fail_global(int code, const char* msg) { return fail(nullptr, code, msg); }
}  // namespace ib
namespace {

int64_t pad_to(int64_t n, int64_t m) { return (n + m - 1) / m * m; }

// IBIS_B200_TIMING=1: wall time of the stages of ibis_b200_create on stderr
struct StageTimer {
    bool on;
    double t0;
    static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    StageTimer() : on(getenv("IBIS_B200_TIMING") != nullptr), t0(now()) {}
    void lap(const char* what) {
        if (!on) return;
        const double t = now();
        fprintf(stderr, "[ibis_b200 create] %-28s %8.3f s\n", what, t - t0);
        t0 = t;
    }
};

template <typename T>
int upload_vec(ibis_b200_ctx* c, const std::vector<T>& h, T** d) {
    void* p = nullptr;
    if (gpu::dev_malloc(&p, h.size() * sizeof(T)) != 0) return IBIS_B200_ENOMEM;
    c->owned_allocs.push_back(p);
    if (!h.empty() && gpu::h2d(p, h.data(), h.size() * sizeof(T), c->stream) != 0) return IBIS_B200_ECUDA;
    if (gpu::stream_sync(c->stream) != 0) return IBIS_B200_ECUDA;   // h dies with the caller
    *d = static_cast<T*>(p);
    return 0;
}

template <typename T>
int dev_alloc_zero(ibis_b200_ctx* c, int64_t n, T** d) {
    void* p = nullptr;
    if (gpu::dev_malloc(&p, (size_t)n * sizeof(T)) != 0) return IBIS_B200_ENOMEM;
    c->owned_allocs.push_back(p);
    if (n > 0 && gpu::memset_async(p, 0, (size_t)n * sizeof(T), c->stream) != 0) return IBIS_B200_ECUDA;
    *d = static_cast<T*>(p);
    return 0;
}

Buffer* get_buf(ibis_b200_ctx* c, int h, BufKind kind) {
    if (h <= 0 || h >= (int)c->bufs.size()) return nullptr;
    Buffer* b = &c->bufs[h];
    if (b->kind != kind) return nullptr;
    return b;
}

template <typename S>
void comp_ptrs(const Buffer* b, S** out /* [7] */) {
    for (int i = 0; i < 7; i++)
        out[i] = i < b->ncomp ? static_cast<S*>(b->d) + (int64_t)i * b->stride : nullptr;
}

int new_buffer(ibis_b200_ctx* c, BufKind kind, bool is_dual, int ncomp, int64_t n, int64_t stride, int* handle) {
    Buffer b;
    b.kind = kind;
    b.is_dual = is_dual;
    b.ncomp = ncomp;
    b.n = n;
    b.stride = stride;
    const size_t bytes = (size_t)ncomp * stride * (is_dual ? 16 : 8);
    if (gpu::dev_malloc(&b.d, bytes) != 0) return fail(c, IBIS_B200_ENOMEM, "out of device memory");
    GPU_TRY(c, gpu::memset_async(b.d, 0, bytes, c->stream));
    for (size_t i = 1; i < c->bufs.size(); i++)
        if (c->bufs[i].kind == BUF_FREE) {
            c->bufs[i] = b;
            *handle = (int)i;
            return 0;
        }
    c->bufs.push_back(b);
    *handle = (int)c->bufs.size() - 1;
    return 0;
}

int ensure_stage(ibis_b200_ctx* c, int64_t elems) {
    if (elems <= c->stage_elems) return 0;
    if (c->d_stage) gpu::dev_free(c->d_stage);
    void* p = nullptr;
    if (gpu::dev_malloc(&p, (size_t)elems * sizeof(double)) != 0) {
        c->d_stage = nullptr;
        c->stage_elems = 0;
        return fail(c, IBIS_B200_ENOMEM, "out of device memory (staging)");
    }
    c->d_stage = static_cast<double*>(p);
    c->stage_elems = elems;
    return 0;
}

template <typename S>
FvArgs<S> make_args(ibis_b200_ctx* c, Buffer* fs, Buffer* dudt, bool allow_reconstruction) {
    FvArgs<S> A;
    A.g = c->g;
    S* p[7];
    comp_ptrs<S>(fs, p);
    for (int i = 0; i < 7; i++) A.fs[i] = p[i];
    if (dudt) {
        comp_ptrs<S>(dudt, p);
        for (int i = 0; i < 5; i++) A.dudt[i] = p[i];
    } else {
        for (int i = 0; i < 5; i++) A.dudt[i] = nullptr;
    }
    if (sizeof(S) == sizeof(double)) {
        A.grad = reinterpret_cast<S*>(c->grad_r);
        A.lim = reinterpret_cast<S*>(c->lim_r);
        A.flux = reinterpret_cast<S*>(c->flux_r);
        A.sidesP = reinterpret_cast<S*>(c->sidesP_r);
    } else {
        A.grad = reinterpret_cast<S*>(c->grad_d);
        A.lim = reinterpret_cast<S*>(c->lim_d);
        A.flux = reinterpret_cast<S*>(c->flux_d);
        A.sidesP = reinterpret_cast<S*>(c->sidesP_d);
    }
    {   // the five reconstructed variables of pass F with their gradient / limiter arrays
        const int src[5] = {F_RHO, (c->cfg.thermo_interpolator == THERMO_RHO_P) ? F_P
                                   : (c->cfg.thermo_interpolator == THERMO_RHO_T) ? F_T : F_E,
                            F_VX, F_VY, F_VZ};
        for (int v = 0; v < 5; v++) {
            A.rc.val[v] = A.fs[src[v]];
            for (int k = 0; k < 3; k++) A.rc.g[v][k] = A.grad ? A.grad + (int64_t)(v * 3 + k) * c->cstride : nullptr;
            A.rc.lim[v] = A.lim ? A.lim + (int64_t)v * c->cstride : nullptr;
        }
    }
    A.dt_out = nullptr;
    A.pipe_dbg = c->d_pipe_dbg;
    A.face_T = c->face_T;
    A.face_vel = nullptr;              // set by the moving-grid evaluation only
    A.face_rho = c->face_rho;
    A.gas.R = S(c->cfg.R);
    A.gas.Cv = S(c->cfg.Cv);
    A.gas.Cp = S(c->cfg.Cp);
    A.gas.gamma = S(c->cfg.gamma);
    A.mu0 = S(c->cfg.mu0);
    A.T0 = S(c->cfg.T0);
    A.Ts = S(c->cfg.Ts);
    A.Pr = S(c->cfg.Pr);
    A.eps = c->cfg.limiter_epsilon;
    A.delta = c->cfg.ldfss_delta;
    A.signal_factor = c->cfg.viscous_signal_factor;
    A.thermo = c->cfg.thermo_interpolator;
    A.th2 = (A.thermo == THERMO_RHO_P) ? F_P : (A.thermo == THERMO_RHO_T) ? F_T : F_E;
    A.limiter_on = c->cfg.limiter == IBIS_B200_LIMITER_BARTH_JESPERSEN;
    A.order2 = allow_reconstruction && c->cfg.reconstruction_order == 2;
    A.viscous = c->cfg.viscous != 0;
    return A;
}

// CubicSpline::CubicSpline (util/cubic_spline.cpp:9-89): second derivatives of the natural spline through
// (x_i, y_i) by the reference's assembly and Thomas sweep, operation for operation.  The result is indexed
// by knot, n entries with y''_0 = y''_{n-1} = 0 (see spline_eval in fv_bodies.h for the reference's two
// out-of-bounds entries).
std::vector<double> spline_second_derivatives(const double* x, const double* y, int n) {
    const int ne = n - 2;
    std::vector<double> diag(ne, 0.0), upper(ne, 0.0), rhs(ne, 0.0), ydd(n, 0.0);
    for (int i = 0; i <= ne; i++) {
        const double dxi = x[i + 1] - x[i], dyi = y[i + 1] - y[i];
        if (i != ne) {
            diag[i] += dxi;
            rhs[i] -= dyi / dxi;
        }
        if (i == 0) continue;
        rhs[i - 1] += dyi / dxi;
        diag[i - 1] += dxi;
        upper[i - 1] += dxi;        // upper[ne - 1] is never read
    }
    for (int i = 0; i < ne; i++) {
        rhs[i] *= 6;
        diag[i] *= 2;
    }
    for (int i = 1; i < ne; i++) {
        const double w = upper[i - 1] / diag[i - 1];
        diag[i] = diag[i] - w * upper[i - 1];
        rhs[i] = rhs[i] - w * rhs[i - 1];
    }
    ydd[n - 2] = rhs[ne - 1] / diag[ne - 1];
    for (int i = ne - 2; i >= 0; i--) ydd[i + 1] = (rhs[i] - upper[i] * ydd[i + 2]) / diag[i];
    return ydd;
}

// lazily allocated FiniteVolume temporaries for one scalar type
// second-order inviscid evaluations go through the sides path (fv_bodies.h)
bool sides_path(const ibis_b200_ctx* c) {
    return c->use_sides && c->cfg.reconstruction_order == 2 && !c->cfg.viscous && !c->has_flux_actions;
}

// dual-number evaluations (16-byte scalars) take the pipelined sides kernels too unless IBIS_B200_DUAL_SIDES=0
// sends them through the gather path (what round 1 did: its one-tile-per-block sides kernels were slower on duals)
template <typename S>
bool sides_path_for(const ibis_b200_ctx* c) { return sides_path(c) && (sizeof(S) == sizeof(double) || c->dual_sides); }

template <typename S>
int ensure_fv_temporaries(ibis_b200_ctx* c, bool want_grad = false, bool want_flux = false) {
    const bool need_grad = want_grad || c->cfg.viscous || (c->cfg.reconstruction_order > 1 && !sides_path_for<S>(c));
    if (sides_path_for<S>(c)) {
        S** sp = (sizeof(S) == 8) ? reinterpret_cast<S**>(&c->sidesP_r) : reinterpret_cast<S**>(&c->sidesP_d);
        if (!*sp) {
            int rc = dev_alloc_zero<S>(c, (int64_t)SIDE_REC * std::max<int64_t>(c->p_total, 1), sp);
            if (rc) return fail(c, rc, "out of device memory (interface sides)");
        }
    }
    S** grad = (sizeof(S) == 8) ? reinterpret_cast<S**>(&c->grad_r) : reinterpret_cast<S**>(&c->grad_d);
    S** lim = (sizeof(S) == 8) ? reinterpret_cast<S**>(&c->lim_r) : reinterpret_cast<S**>(&c->lim_d);
    S** flux = (sizeof(S) == 8) ? reinterpret_cast<S**>(&c->flux_r) : reinterpret_cast<S**>(&c->flux_d);
    if (need_grad && !*grad) {
        int rc = dev_alloc_zero<S>(c, (int64_t)NVARS * 3 * c->cstride, grad);
        if (rc) return fail(c, rc, "out of device memory (gradients)");
        rc = dev_alloc_zero<S>(c, (int64_t)NVARS * c->cstride, lim);
        if (rc) return fail(c, rc, "out of device memory (limiters)");
    }
    if ((c->cfg.viscous || c->has_flux_actions || want_flux) && !*flux) {
        int rc = dev_alloc_zero<S>(c, (int64_t)5 * c->fstride, flux);
        if (rc) return fail(c, rc, "out of device memory (face flux)");
    }
    if (c->cfg.viscous && !c->face_T) {
        int rc = dev_alloc_zero<double>(c, c->fstride, &c->face_T);
        if (rc) return fail(c, rc, "out of device memory (face_fs)");
        rc = dev_alloc_zero<double>(c, c->fstride, &c->face_rho);
        if (rc) return fail(c, rc, "out of device memory (face_fs)");
    }
    return 0;
}

// FiniteVolume<T>::compute_dudt (finite_volume.cpp:57-96)
// dt_bits: when not null and the evaluation takes the sides path, the minimum local time step of the
// owned cells is left there (as the bit pattern of a double; the caller initialises it to +inf)
// cq != nullptr: the moving-grid overload (finite_volume.cpp:57-85) -- interface velocities from the vertex
// velocities last set, interface-relative fluxes through the stored face-flux array, GCL term with cq
template <typename S>
int compute_dudt_impl(ibis_b200_ctx* c, Buffer* fs, Buffer* dudt, bool allow_reconstruction, double* dt_bits = nullptr,
                      Buffer* cq = nullptr) {
    const bool moving = cq != nullptr;
    if (!moving && c->geometry_moved)
        return fail(c, IBIS_B200_EINVAL, "the vertices have moved: only ibis_b200_compute_dudt_moving is valid on this context");
    int rc = ensure_fv_temporaries<S>(c, moving && c->cfg.reconstruction_order > 1, moving);
    if (rc) return rc;
    FvArgs<S> A = make_args<S>(c, fs, dudt, allow_reconstruction);
    A.dt_out = reinterpret_cast<unsigned long long*>(dt_bits);
    if (moving) {
        prof_begin(c, 2);
        rc = comm::halo_exchange<S>(c, A.fs);
        prof_end(c, 2);
        if (rc) return rc;
        launch_bc(c, A, IBIS_B200_PHASE_PRE_RECONSTRUCTION);
        launch_face_vel(c, c->d_fverts, c->d_vertex_vel, c->d_face_vel);      // grid.compute_grid_motion (grid.h:562-565)
        A.face_vel = c->d_face_vel;
        if (A.order2) launch_grad_limit(c, A, false);
        launch_face_flux_moving(c, A);
        if (c->has_flux_actions) launch_bc(c, A, IBIS_B200_PHASE_POST_CONVECTIVE_FLUX);
        if (A.viscous) {
            if (c->has_previscous) launch_bc(c, A, IBIS_B200_PHASE_PRE_VISCOUS_GRAD);
            launch_grad_limit(c, A, true);
            launch_face_viscous(c, A);
        }
        launch_integral(c, A);
        S* q[7];
        comp_ptrs<S>(cq, q);
        launch_gcl(c, A, q);
        if (!c->last_error.empty()) return IBIS_B200_ECUDA;
        return 0;
    }
    if (sizeof(S) == sizeof(double)) c->last_fs = (int)(fs - c->bufs.data());
    if (A.order2 && sides_path_for<S>(c) && comm::halo_split(c)) {
        // overlap: this rank's boundary cells go out to the peers, the tiles that touch neither a halo
        // cell nor a ghost run pass G' meanwhile, then the peers' data is awaited and scattered, the
        // ghosts filled and the remaining tiles (incl. the inner halo ring) done
        rc = comm::halo_exchange_begin<S>(c, A.fs);
        if (rc) return rc;
        launch_grad_sides(c, A, 1);
        prof_begin(c, 2);
        rc = comm::halo_exchange_end<S>(c, A.fs);
        prof_end(c, 2);
        if (rc) return rc;
        launch_bc(c, A, IBIS_B200_PHASE_PRE_RECONSTRUCTION);
        launch_grad_sides(c, A, 2);
        launch_flux_tile(c, A);
        rc = comm::halo_exchange_join(c);              // the pushes (side stream) have read the flow state
        if (rc) return rc;
        if (!c->last_error.empty()) return IBIS_B200_ECUDA;
        return 0;
    }
    prof_begin(c, 2);
    rc = comm::halo_exchange<S>(c, A.fs);
    prof_end(c, 2);
    if (rc) return rc;
    launch_bc(c, A, IBIS_B200_PHASE_PRE_RECONSTRUCTION);
    if (A.order2 && sides_path_for<S>(c)) {
        launch_grad_sides(c, A, 0);
        launch_flux_tile(c, A);
        if (!c->last_error.empty()) return IBIS_B200_ECUDA;
        return 0;
    }
    if (A.order2) launch_grad_limit(c, A, false);
    if (A.viscous || c->has_flux_actions) {
        // the reference's own sequence over a stored face-flux array (finite_volume.cpp:62-81)
        launch_face_flux(c, A);
        if (c->has_flux_actions) launch_bc(c, A, IBIS_B200_PHASE_POST_CONVECTIVE_FLUX);
        if (A.viscous) {
            if (c->has_previscous) launch_bc(c, A, IBIS_B200_PHASE_PRE_VISCOUS_GRAD);
            launch_grad_limit(c, A, true);
            launch_face_viscous(c, A);
        }
        launch_integral(c, A);
    } else {
        launch_residual(c, A);
    }
    if (!c->last_error.empty()) return IBIS_B200_ECUDA;
    return 0;
}

// after a synchronise: did a pipelined kernel give up on a stage barrier (fv_pipe.cu::pipe_wait)?
int check_pipe_stall(ibis_b200_ctx* c) {
    if (!c->h_pipe_dbg || c->h_pipe_dbg[0] == 0) return 0;
    const unsigned int* d = c->h_pipe_dbg;
    char msg[256];
    if (d[1] == 64u) {
        snprintf(msg, sizeof msg, "halo exchange %u: a peer never published its cells (it had published %u); the ranks have made "
                                  "different numbers of evaluations or a peer has failed", d[4], d[5]);
        return fail(c, IBIS_B200_ECOMM, msg);
    }
    snprintf(msg, sizeof msg, "pipelined kernel stalled and was aborted: barrier kind %u, block %u, warp %u, local tile %u, parity %u",
             d[1], d[2], d[3], d[4], d[5]);
    return fail(c, IBIS_B200_ECUDA, msg);
}
int sync_checked(ibis_b200_ctx* c) {
    if (c->d_pipe_dbg) GPU_TRY(c, gpu::d2h(c->h_pipe_dbg, c->d_pipe_dbg, 8 * sizeof(unsigned int), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return check_pipe_stall(c);
}

int fetch_scalars(ibis_b200_ctx* c, int first, int count, double* out) {
    GPU_TRY(c, gpu::d2h(c->h_scalars + first, c->d_scalars + first, count * sizeof(double), c->stream));
    int rc = sync_checked(c);
    if (rc) return rc;
    for (int i = 0; i < count; i++) out[i] = c->h_scalars[first + i];
    return 0;
}

// slots in d_scalars
enum { SC_DT_STABLE = 0, SC_DT = 1, SC_BAD = 2, SC_DOT = 4, SC_NORM = 6, SC_L2 = 8, SC_H = 16, SC_Y = 16 };
constexpr int N_SCALARS = 16 + 1024;   // h / y of up to 1023 Krylov vectors

}  // namespace

extern "C" {

const char* ibis_b200_last_error(const ibis_b200_ctx* c) {
    if (!c) return g_last_error.c_str();
    return !c->last_error.empty() ? c->last_error.c_str() : c->last_message.c_str();
}

int ibis_b200_create(ibis_b200_ctx** out, const ibis_b200_grid_desc* gd_ref, const ibis_b200_config* cfg, int device) {
    if (!out || !gd_ref || !cfg) return fail(nullptr, IBIS_B200_EINVAL, "null argument");
    // internal renumbering (layout.h); IBIS_B200_RENUMBER=0 keeps the reference numbering
    const char* env_renumber = getenv("IBIS_B200_RENUMBER");
    const bool morton = !(env_renumber && env_renumber[0] == '0');
    StageTimer timer;
    Renumbered R;
    {
        const char* err = nullptr;
        int rrc = renumber_grid(gd_ref, morton, R, &err);
        if (rrc) return fail(nullptr, rrc, err ? err : "renumbering failed");
    }
    timer.lap("renumber_grid (Morton)");
    // tiles and the tile-owned interface numbering (layout.h)
    TileLayout TL;
    {
        int T = 64;
        if (const char* e = getenv("IBIS_B200_TILE")) T = atoi(e);
        if (T != 64 && T != 128) T = 64;
        const char* err = nullptr;
        int rrc = regroup_faces(R, T, TL, &err);
        if (rrc) return fail(nullptr, rrc, err ? err : "interface regrouping failed");
    }
    timer.lap("regroup_faces (tiles)");
    const ibis_b200_grid_desc* gd = &R.desc;
    if (gd->dim != 2 && gd->dim != 3) return fail(nullptr, IBIS_B200_EINVAL, "dim must be 2 or 3");
    if (cfg->thermo_interpolator == IBIS_B200_THERMO_P_T)
        return fail(nullptr, IBIS_B200_EINVAL,
                    "thermo_interpolator p_T is broken in the reference (convective_flux.cpp:53-55) and unsupported");
    if (cfg->reconstruction_order != 1 && cfg->reconstruction_order != 2)
        return fail(nullptr, IBIS_B200_EINVAL, "Invalid reconstruction order");
    if (cfg->flux_calculator < 0 || cfg->flux_calculator > 3)
        return fail(nullptr, IBIS_B200_EINVAL, "Unknown flux calculator");
    const int64_t n_total = gd->n_cells + gd->n_ghost;
    if (n_total >= (int64_t)1 << 31 || gd->n_faces >= (int64_t)1 << 31)
        return fail(nullptr, IBIS_B200_EINVAL, "grid too large for 32-bit ids on one GPU");

    ibis_b200_ctx* c = new (std::nothrow) ibis_b200_ctx();
    if (!c) return fail(nullptr, IBIS_B200_ENOMEM, "out of host memory");
    c->bufs.resize(1);
    int rc = gpu::set_device(device);
    if (rc != 0) { delete c; return fail(nullptr, IBIS_B200_ECUDA, std::string("cudaSetDevice: ") + gpu::error_string(rc)); }
    gpu::get_device(&c->device);
    c->n_sm = gpu::sm_count(c->device);
    rc = gpu::stream_create(&c->stream);
    if (rc != 0) { delete c; return fail(nullptr, IBIS_B200_ECUDA, std::string("cudaStreamCreate: ") + gpu::error_string(rc)); }

    c->cfg = *cfg;
    if (const char* e = getenv("IBIS_B200_ORDERED_DOT")) c->ordered_dot = e[0] == '1';
    if (const char* e = getenv("IBIS_B200_SIDES")) c->use_sides = e[0] != '0';
    if (const char* e = getenv("IBIS_B200_DUAL_SIDES")) c->dual_sides = e[0] != '0';
    if (const char* e = getenv("IBIS_B200_RK_UNFUSED")) c->rk_unfused = e[0] == '1';
    if (const char* e = getenv("IBIS_B200_PREFETCH")) c->pipe_prefetch = std::max(0, atoi(e));
    if (const char* e = getenv("IBIS_B200_F_STAGES")) c->pipe_f_stages = std::max(0, atoi(e));
    c->bc_actions.assign(cfg->bc_actions, cfg->bc_actions + cfg->n_bc_actions);
    c->cfg.bc_actions = nullptr;
    c->dim = gd->dim;
    c->n_cons = gd->dim + 2;
    c->n_tags = gd->n_tags;
    c->n_cells = gd->n_cells;
    c->n_owned = gd->n_owned > 0 ? gd->n_owned : gd->n_cells;
    c->n_ghost = gd->n_ghost;
    c->n_total = n_total;
    c->n_faces = gd->n_faces;
    c->cstride = pad_to(std::max<int64_t>(n_total, 1) + 256, 256);   // whole-tile bulk copies that start at any cell stay in bounds
    c->fstride = pad_to(std::max<int64_t>(gd->n_faces, 1), 32);
    const int64_t cs = c->cstride, fs = c->fstride;

    // ---- cell -> face ELL ----
    int max_faces = 0;
    for (int64_t i = 0; i < gd->n_cells; i++)
        max_faces = std::max<int>(max_faces, (int)(gd->cell_face_offsets[i + 1] - gd->cell_face_offsets[i]));
    if (max_faces > MAXF) { ibis_b200_destroy(c); return fail(nullptr, IBIS_B200_EINVAL, "cell with more than 6 faces"); }
    if (max_faces == 0) max_faces = 1;
    std::vector<int> cf_face((size_t)max_faces * cs, -1), cf_nbr((size_t)max_faces * cs, -1);
    std::vector<signed char> cf_out((size_t)MAXF * cs, 0);            // MAXF wide, see GridDev
    // cells with their full face list come first; halo cells of the outer layer carry none
    // (renumber_grid has checked that they form a prefix)
    const int64_t n_grad = R.n_grad;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < gd->n_cells; i++) {
        const int64_t o = gd->cell_face_offsets[i];
        const int nf = (int)(gd->cell_face_offsets[i + 1] - o);
        for (int j = 0; j < nf; j++) {
            cf_face[(size_t)j * cs + i] = (int)gd->cell_face_ids[o + j];
            cf_nbr[(size_t)j * cs + i] = (int)gd->cell_neighbours[o + j];
            cf_out[(size_t)j * cs + i] = (signed char)gd->cell_outsigns[o + j];
        }
    }
    std::vector<double> cxyz((size_t)3 * cs, 0.0), vol((size_t)cs, 1.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_total; i++) {
        for (int d = 0; d < 3; d++) cxyz[(size_t)d * cs + i] = gd->cell_centroid[3 * i + d];
        if (i < gd->n_cells) vol[i] = gd->cell_volume[i];
    }

    // ---- faces SoA ----
    std::vector<int> fl(fs, 0), fr(fs, 0);
    std::vector<double> fa(fs, 0.0), fn((size_t)3 * fs, 0.0), ft1((size_t)3 * fs, 0.0), ft2((size_t)3 * fs, 0.0),
        fc((size_t)3 * fs, 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < gd->n_faces; f++) {
        fl[f] = (int)gd->face_left[f];
        fr[f] = (int)gd->face_right[f];
        fa[f] = gd->face_area[f];
        for (int d = 0; d < 3; d++) {
            fn[(size_t)d * fs + f] = gd->face_norm[3 * f + d];
            ft1[(size_t)d * fs + f] = gd->face_tan1[3 * f + d];
            ft2[(size_t)d * fs + f] = gd->face_tan2[3 * f + d];
            fc[(size_t)d * fs + f] = gd->face_centre[3 * f + d];
        }
    }
    // ---- boundary faces and action tables ----
    const int64_t n_bf = gd->n_tags > 0 ? gd->tag_offsets[gd->n_tags] : 0;
    std::vector<int> bf_face(std::max<int64_t>(n_bf, 1), 0), bf_tag(std::max<int64_t>(n_bf, 1), 0);
    for (int t = 0; t < gd->n_tags; t++)
        for (int64_t k = gd->tag_offsets[t]; k < gd->tag_offsets[t + 1]; k++) {
            bf_face[k] = (int)gd->tag_faces[k];
            bf_tag[k] = t;
        }
    std::vector<BcAction> acts;
    std::vector<double> bc_tab;
    std::vector<int> act_off((size_t)N_PHASES * (gd->n_tags + 1), 0);
    for (int phase = 0; phase < N_PHASES; phase++) {
        for (int t = 0; t < gd->n_tags; t++) {
            act_off[(size_t)phase * (gd->n_tags + 1) + t] = (int)acts.size();
            for (const auto& a : c->bc_actions)
                if (a.tag == t && a.phase == phase) {
                    // the reference has one factory per kind (boundary.cpp:449-492): a ghost-cell action
                    // in a flux list, or the reverse, is "Unknown ... action" there
                    const bool flux_action = a.type == IBIS_B200_BC_CONSTANT_FLUX;
                    if (a.type < 0 || a.type > IBIS_B200_BC_CONSTANT_FLUX ||
                        flux_action != (phase == IBIS_B200_PHASE_POST_CONVECTIVE_FLUX)) {
                        ibis_b200_destroy(c);
                        return fail(nullptr, IBIS_B200_EINVAL, phase == IBIS_B200_PHASE_POST_CONVECTIVE_FLUX
                                                                   ? "Unknown flux action" : "Unknown boundary action");
                    }
                    BcAction b;
                    b.type = a.type;
                    b.pad = 0;
                    for (int q = 0; q < 7; q++) b.params[q] = a.params[q];
                    if (a.type == IBIS_B200_BC_BOUNDARY_LAYER_PROFILE) {
                        const int64_t n = (int64_t)a.params[1], off = (int64_t)a.params[2];
                        std::string why;
                        if (n < 4) why = "Insufficient points for cubic spline";        // cubic_spline.cpp:12-14
                        else if (!cfg->bc_tables || off < 0 || off + 3 * n > cfg->n_bc_tables)
                            why = "boundary_layer_profile: table outside bc_tables";
                        if (!why.empty()) {
                            ibis_b200_destroy(c);
                            return fail(nullptr, IBIS_B200_EINVAL, why.c_str());
                        }
                        const double* x = cfg->bc_tables + off;
                        b.params[2] = (double)bc_tab.size();
                        bc_tab.insert(bc_tab.end(), x, x + 3 * n);
                        for (int q = 1; q <= 2; q++) {
                            const std::vector<double> dd = spline_second_derivatives(x, x + q * n, (int)n);
                            bc_tab.insert(bc_tab.end(), dd.begin(), dd.end());
                        }
                    }
                    acts.push_back(b);
                    if (phase == IBIS_B200_PHASE_PRE_VISCOUS_GRAD) c->has_previscous = true;
                    if (phase == IBIS_B200_PHASE_POST_CONVECTIVE_FLUX) c->has_flux_actions = true;
                }
        }
        act_off[(size_t)phase * (gd->n_tags + 1) + gd->n_tags] = (int)acts.size();
    }
    if (bc_tab.empty()) bc_tab.push_back(0.0);
    if (acts.empty()) acts.push_back(BcAction());

    c->ref2int = R.ref2int;
    c->face_ref2int.swap(R.face_ref2int);
    GridDev& g = c->g;
    g.dim = gd->dim;
    g.n_owned = (int)c->n_owned;
    g.n_grad = (int)n_grad;
    g.n_cells = (int)gd->n_cells;
    g.n_total = (int)n_total;
    g.n_faces = (int)gd->n_faces;
    g.n_bfaces = (int)n_bf;
    g.max_faces = max_faces;
    g.cstride = cs;
    g.fstride = fs;
    g.n_tags = gd->n_tags;
    c->n_bfaces = n_bf;
    int* di = nullptr;
    signed char* dc = nullptr;
    double* dd = nullptr;
    BcAction* da = nullptr;
#define UP(vec, ptr, field)                                                        \
    rc = upload_vec(c, vec, &ptr);                                                 \
    if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); } \
    g.field = ptr;
    UP(cf_face, di, cf_face) UP(cf_nbr, di, cf_nbr) UP(cf_out, dc, cf_out)
    UP(vol, dd, volume)
    UP(cxyz, dd, cx)
    g.cy = g.cx + cs;
    g.cz = g.cx + 2 * cs;
    UP(fl, di, f_left) UP(fr, di, f_right) UP(fa, dd, f_area)
    UP(fn, dd, f_n) UP(ft1, dd, f_t1) UP(ft2, dd, f_t2) UP(fc, dd, f_c)
    {
        std::vector<double> fgeom((size_t)10 * std::max<int64_t>(gd->n_faces, 1), 0.0);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < gd->n_faces; f++) {
            double* q = &fgeom[(size_t)10 * f];
            for (int d = 0; d < 3; d++) {
                q[d] = gd->face_norm[3 * f + d];
                q[3 + d] = gd->face_tan1[3 * f + d];
                q[6 + d] = gd->face_tan2[3 * f + d];
            }
            q[9] = gd->face_area[f];
        }
        UP(fgeom, dd, f_geom)
    }
    UP(bf_face, di, bf_face) UP(bf_tag, di, bf_tag)
    UP(acts, da, act) UP(act_off, di, act_off) UP(bc_tab, dd, bc_tab)
    if (morton) {
        rc = upload_vec(c, c->ref2int, &c->d_ref2int);
        if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
    }
#undef UP
    timer.lap("SoA build + upload");
    // ---- tiles (layout.h): slot / staging tables of the tiled kernels ----
    {
        const int T = TL.T;
        const int nt = TL.nt_all, n_tiles = TL.nt_owned;
        // flux slots of a tile: its home interfaces [0, nh), then the interfaces of its entries
        std::vector<int> foreign_off(nt + 1, 0);
        for (int b = 0; b < nt; b++) {
            int nf = 0;
            for (int e = TL.ent_off[b]; e < TL.ent_off[b + 1]; e++) nf += TL.entries[e].count;
            foreign_off[b + 1] = foreign_off[b] + nf;
        }
        std::vector<int> foreign(std::max(foreign_off[nt], 1), 0);
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nt; b++) {
            int k = foreign_off[b];
            for (int e = TL.ent_off[b]; e < TL.ent_off[b + 1]; e++)
                for (int q = 0; q < TL.entries[e].count; q++) foreign[k++] = TL.entries[e].f_start + q;
        }
        // per cell face: flux slot in the cell's tile, and the staging index of the cell's own side
        std::vector<unsigned short> cf_slot((size_t)MAXF * cs, 0xFFFF), cf_stage((size_t)MAXF * cs, 0xFFFF);
        bool bad_table = false;
#pragma omp parallel for schedule(static) reduction(|| : bad_table)
        for (int b = 0; b < nt; b++) {
            const int h0 = TL.h0[b], nh = TL.h0[b + 1] - h0, nR = TL.nR[b];
            const int n_own_l = TL.n_own_left(b);
            for (int64_t i = TL.c0[b]; i < TL.c0[b + 1]; i++)
                for (int j = 0; j < max_faces; j++) {
                    const int f = cf_face[(size_t)j * cs + i];
                    if (f < 0) continue;
                    const bool is_left = cf_out[(size_t)j * cs + i] > 0;
                    int slot, stage;
                    if (TL.home_tile[f] == b) {
                        slot = f - h0;
                        stage = is_left ? slot - nR : n_own_l + slot;
                        if (stage < 0) bad_table = true;
                    } else {
                        slot = TL.slot_other[f];
                        stage = -1;
                        for (int e = TL.ent_off[b]; e < TL.ent_off[b + 1]; e++) {
                            const TileEntry& te = TL.entries[e];
                            if (f >= te.f_start && f < te.f_start + te.count) stage = te.stage_start + (f - te.f_start);
                        }
                        if (slot < 0 || stage < 0) bad_table = true;
                    }
                    if (slot >= 0xFFFF || stage >= 0xFFFF) bad_table = true;
                    cf_slot[(size_t)j * cs + i] = (unsigned short)slot;
                    cf_stage[(size_t)j * cs + i] = (unsigned short)stage;
                }
        }
        const int cap = TL.cap_slots;
        if (bad_table || cap >= 0x7FFF || (size_t)cap * frec<dual>::bytes > 220 * 1024 ||
            (size_t)TL.cap_stage * 6 * sizeof(dual) > 150 * 1024) {
            ibis_b200_destroy(c);
            return fail(nullptr, IBIS_B200_EINVAL, "tile has too many interfaces for shared memory");
        }
        // tile descriptors for the kernels: {c0, nc, h0, nh, nR, n0, ent_off, n_ent, n_own_l, n_slots, -, -}
        std::vector<int> tile_info((size_t)TILE_INFO * std::max(nt, 1), 0);
        for (int b = 0; b < nt; b++) {
            int* q = &tile_info[(size_t)TILE_INFO * b];
            q[0] = TL.c0[b]; q[1] = TL.c0[b + 1] - TL.c0[b]; q[2] = TL.h0[b]; q[3] = TL.h0[b + 1] - TL.h0[b];
            q[4] = TL.nR[b]; q[5] = TL.n0[b]; q[6] = TL.ent_off[b]; q[7] = TL.ent_off[b + 1] - TL.ent_off[b];
            q[8] = TL.n_own_left(b);
            q[9] = q[3];
            for (int e = TL.ent_off[b]; e < TL.ent_off[b + 1]; e++) q[9] += TL.entries[e].count;
            if (q[7] <= TILE_INLINE_ENT)
                for (int e = 0; e < q[7]; e++) {
                    const TileEntry& te = TL.entries[TL.ent_off[b] + e];
                    int* w = q + 12 + 5 * e;
                    w[0] = te.f_start; w[1] = te.count; w[2] = te.section; w[3] = te.slot_start; w[4] = te.stage_start;
                }
        }
        std::vector<int> tile_ent((size_t)5 * std::max<size_t>(TL.entries.size(), 1), 0);
        for (size_t e = 0; e < TL.entries.size(); e++) {
            int* q = &tile_ent[5 * e];
            q[0] = TL.entries[e].f_start; q[1] = TL.entries[e].count; q[2] = TL.entries[e].section;
            q[3] = TL.entries[e].slot_start; q[4] = TL.entries[e].stage_start;
        }
        // early tiles: owned tiles none of whose cells has a neighbour outside the owned range
        std::vector<int> tiles_early, tiles_late;
        for (int b = 0; b < nt; b++) {
            bool early = b < n_tiles;
            for (int64_t i = TL.c0[b]; early && i < TL.c0[b + 1]; i++)
                for (int j = 0; j < max_faces; j++) {
                    const int nb = cf_nbr[(size_t)j * cs + i];
                    if (cf_face[(size_t)j * cs + i] >= 0 && nb >= c->n_owned) { early = false; break; }
                }
            (early ? tiles_early : tiles_late).push_back(b);
        }
        g.n_early = (int)tiles_early.size();
        g.n_late = (int)tiles_late.size();
        if (tiles_early.empty()) tiles_early.push_back(0);
        if (tiles_late.empty()) tiles_late.push_back(0);
        std::vector<int> home0(TL.h0.begin(), TL.h0.end());
        // sides of the (owned) tiles' interfaces that belong to cells outside the tile: gather path
        std::vector<int> out_off(n_tiles + 1, 0);
        auto face_of_slot = [&](int b, int slot) {
            const int nh = home0[b + 1] - home0[b];
            return slot < nh ? home0[b] + slot : foreign[foreign_off[b] + (slot - nh)];
        };
#pragma omp parallel for schedule(static)
        for (int b = 0; b < n_tiles; b++) {
            const int64_t c0 = TL.c0[b], c1 = TL.c0[b + 1];
            const int ns = (home0[b + 1] - home0[b]) + (foreign_off[b + 1] - foreign_off[b]);
            int n = 0;
            for (int sl = 0; sl < ns; sl++) {
                const int f = face_of_slot(b, sl);
                if (!(fl[f] >= c0 && fl[f] < c1)) n++;
                if (!(fr[f] >= c0 && fr[f] < c1)) n++;
            }
            out_off[b + 1] = n;
        }
        for (int b = 0; b < n_tiles; b++) out_off[b + 1] += out_off[b];
        const int64_t n_out = std::max(out_off[n_tiles], 1);
        std::vector<unsigned short> out_slot(n_out, 0);
        std::vector<int> out_cell(n_out, 0);
        std::vector<double> out_d((size_t)3 * n_out, 0.0);
        auto set_out = [&](int k, int slot_side, int f, int cell) {
            out_slot[k] = (unsigned short)slot_side;
            out_cell[k] = cell;
            for (int d = 0; d < 3; d++)     // face centre - centroid, as reconstruct_side computes it
                out_d[(size_t)d * n_out + k] = fc[(size_t)d * fs + f] - cxyz[(size_t)d * cs + cell];
        };
#pragma omp parallel for schedule(static)
        for (int b = 0; b < n_tiles; b++) {
            const int64_t c0 = TL.c0[b], c1 = TL.c0[b + 1];
            const int ns = (home0[b + 1] - home0[b]) + (foreign_off[b + 1] - foreign_off[b]);
            int k = out_off[b];
            for (int sl = 0; sl < ns; sl++) {
                const int f = face_of_slot(b, sl);
                if (!(fl[f] >= c0 && fl[f] < c1)) set_out(k++, sl, f, fl[f]);
                if (!(fr[f] >= c0 && fr[f] < c1)) set_out(k++, sl | 0x8000, f, fr[f]);
            }
        }
        g.tile = T;
        g.n_tiles = n_tiles;
        g.n_tiles_all = nt;
        g.tile_cap = cap;
        g.stage_cap = TL.cap_stage;
        unsigned short* du16 = nullptr;
#define UP2(vec, ptr, field)                                                        \
    rc = upload_vec(c, vec, &ptr);                                                  \
    if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); } \
    g.field = ptr;
        UP2(home0, di, home0) UP2(foreign_off, di, foreign_off) UP2(foreign, di, foreign) UP2(cf_slot, du16, cf_slot)
        UP2(cf_stage, du16, cf_stage) UP2(tile_info, di, tile_info) UP2(tile_ent, di, tile_ent)
        UP2(tiles_early, di, tiles_early) UP2(tiles_late, di, tiles_late)
        UP2(out_off, di, out_off) UP2(out_slot, du16, out_slot) UP2(out_cell, di, out_cell) UP2(out_d, dd, out_d)
        g.n_out = n_out;
        // ---- packed per-tile records of the pipelined sides path (fv_pipe.cu; layout in fv_bodies.h::GridDev) ----
        g.t_wd = nullptr; g.t_grec = nullptr; g.t_freca = nullptr; g.t_frec = nullptr; g.t_pinfo = nullptr; g.t_pull = nullptr;
        g.grec_stride = g.freca_stride = g.frec_stride = g.halo_cap = g.pull_cap = g.pull_stride = g.p_cap = 0;
        g.pipe_nf = max_faces < 3 ? 3 : max_faces;
        // the kernels recompute the interface normal as cross(t1, t2), the reference's own formula
        // (grid/interface.h:250): arrays that do not satisfy it bit for bit keep to the gather path
        bool normals_exact = true;
#pragma omp parallel for schedule(static) reduction(&& : normals_exact)
        for (int64_t f = 0; f < gd->n_faces; f++) {
            const double* a = gd->face_tan1 + 3 * f;
            const double* b = gd->face_tan2 + 3 * f;
            const double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            if (memcmp(n, gd->face_norm + 3 * f, sizeof n) != 0) normals_exact = false;
        }
        if (!normals_exact) c->use_sides = false;
        c->p_total = 0;
        if (sides_path(c)) {
            const int NFr = g.pipe_nf;
            const int capp = (int)pad_to(cap, 4);
            g.tile_cap = capp;
            auto tile_of_cell = [&](int64_t cell) -> int {
                if (cell < c->n_owned) return (int)(cell / T);
                if (cell < n_grad) return n_tiles + (int)((cell - c->n_owned) / T);
                return -1;
            };
            auto slot_in_cell = [&](int64_t cell, int f) -> int {
                for (int j = 0; j < max_faces; j++)
                    if (cf_face[(size_t)j * cs + cell] == f) return j;
                return -1;
            };
            // neighbours of a tile's cells outside the tile (sorted, without duplicates) and the size of the tile's
            // staging image (its cells' sides + the ghost sides of its boundary interfaces)
            std::vector<std::vector<int>> halos(nt);
            std::vector<int> n_stage(nt, 0);
#pragma omp parallel for schedule(dynamic, 64)
            for (int b = 0; b < nt; b++) {
                const int64_t c0 = TL.c0[b], c1 = TL.c0[b + 1];
                const int nR = TL.nR[b], n_own_l = TL.n_own_left(b);
                std::vector<int>& h = halos[b];
                int ns = 0;
                for (int64_t i = c0; i < c1; i++)
                    for (int j = 0; j < max_faces; j++) {
                        if (cf_face[(size_t)j * cs + i] < 0) continue;
                        const int nb = cf_nbr[(size_t)j * cs + i];
                        if (nb < c0 || nb >= c1) h.push_back(nb);
                        const int st = cf_stage[(size_t)j * cs + i];
                        ns = std::max(ns, st + 1);
                        if (nb >= gd->n_cells) ns = std::max(ns, far_stage_index(cf_out[(size_t)j * cs + i] > 0, st, nR, n_own_l) + 1);
                    }
                std::sort(h.begin(), h.end());
                h.erase(std::unique(h.begin(), h.end()), h.end());
                n_stage[b] = (int)pad_to(ns, SIDE_EVEN);
            }
            g.stage_cap = (int)pad_to(g.stage_cap, SIDE_EVEN);
            int halo_cap = 4;
            for (int b = 0; b < nt; b++) halo_cap = std::max<int>(halo_cap, (int)halos[b].size());
            halo_cap = (int)pad_to(halo_cap, 4);
            if (T + halo_cap > NBR_LOCAL_MASK) { ibis_b200_destroy(c); return fail(nullptr, IBIS_B200_EINVAL, "tile halo too large"); }
            g.halo_cap = halo_cap;
            // the loader of pass G' keeps a tile's halo ids in registers (fv_pipe.cu::PIPE_HALO_REGS): grids whose
            // tiles scatter further than that keep to the gather path
            if (halo_cap > 256) c->use_sides = false;
          if (c->use_sides) {

            // side-state arrays: every tile's staging image goes to its own array (one bulk store of pass G'); pass F'
            // assembles the image of an owned tile's interface sides from its own array and from runs of its
            // neighbours' arrays (the far sides of its surface interfaces)
            std::vector<int> pinfo((size_t)2 * std::max(nt, 1), 0);
            int64_t ptotal = 0;
            for (int b = 0; b < nt; b++) {
                if (ptotal + n_stage[b] >= ((int64_t)1 << 31)) { ibis_b200_destroy(c); return fail(nullptr, IBIS_B200_EINVAL, "too many side records for 32-bit ids"); }
                pinfo[2 * b] = (int)ptotal;
                pinfo[2 * b + 1] = n_stage[b];
                ptotal += n_stage[b];
            }
            c->p_total = ptotal;
            struct Foreign { int a, a_stage, slot, right; };
            struct Pull { int src, count, dst; };
            std::vector<std::vector<Pull>> pulls(std::max(n_tiles, 1));
            std::vector<int> n_image(std::max(n_tiles, 1), 0);
            std::vector<int> lidx((size_t)capp * std::max(n_tiles, 1), 0), ridx((size_t)capp * std::max(n_tiles, 1), 0);
            bool bad_sides = false;
#pragma omp parallel for schedule(dynamic, 64) reduction(|| : bad_sides)
            for (int b = 0; b < n_tiles; b++) {
                const int nR = TL.nR[b], n_own_l = TL.n_own_left(b);
                const int ns = (home0[b + 1] - home0[b]) + (foreign_off[b + 1] - foreign_off[b]);
                std::vector<Foreign> F;
                for (int sl = 0; sl < ns; sl++) {
                    const int f = face_of_slot(b, sl);
                    const int64_t cl[2] = {fl[f], fr[f]};
                    for (int side = 0; side < 2; side++) {
                        const int64_t cell = cl[side], other = cl[1 - side];
                        int idx = -1;
                        if (cell >= gd->n_cells) {                          // ghost: the far side the tile itself stages
                            const int j = slot_in_cell(other, f);
                            if (j < 0 || tile_of_cell(other) != b) { bad_sides = true; continue; }
                            idx = far_stage_index(cf_out[(size_t)j * cs + other] > 0, cf_stage[(size_t)j * cs + other], nR, n_own_l);
                        } else if (tile_of_cell(cell) == b) {
                            const int j = slot_in_cell(cell, f);
                            if (j < 0) { bad_sides = true; continue; }
                            idx = cf_stage[(size_t)j * cs + cell];
                        } else {
                            const int a = tile_of_cell(cell);
                            const int j = a >= 0 ? slot_in_cell(cell, f) : -1;
                            if (j < 0) { bad_sides = true; continue; }
                            F.push_back({a, (int)cf_stage[(size_t)j * cs + cell], sl, side});
                        }
                        (side == 0 ? lidx : ridx)[(size_t)capp * b + sl] = idx;
                    }
                }
                std::sort(F.begin(), F.end(), [](const Foreign& x, const Foreign& y) {
                    return x.a != y.a ? x.a < y.a : x.a_stage != y.a_stage ? x.a_stage < y.a_stage : x.slot < y.slot;
                });
                // runs of consecutive records of one neighbour, widened to even boundaries
                int image = n_stage[b];
                for (size_t q = 0; q < F.size();) {
                    size_t e = q + 1;
                    while (e < F.size() && F[e].a == F[q].a && F[e].a_stage == F[q].a_stage + (int)(e - q)) e++;
                    const int s0 = F[q].a_stage / SIDE_EVEN * SIDE_EVEN;
                    const int s1 = (int)pad_to(F[q].a_stage + (int)(e - q), SIDE_EVEN);      // <= n_stage[a], which is even
                    pulls[b].push_back({pinfo[2 * F[q].a] + s0, s1 - s0, image});
                    for (size_t k = q; k < e; k++)
                        (F[k].right ? ridx : lidx)[(size_t)capp * b + F[k].slot] = image + (F[k].a_stage - s0);
                    image += s1 - s0;
                    q = e;
                }
                n_image[b] = image;
            }
            if (bad_sides) { ibis_b200_destroy(c); return fail(nullptr, IBIS_B200_EINVAL, "inconsistent tile tables"); }
            int p_cap = 4, pull_cap = 1;
            for (int b = 0; b < n_tiles; b++) {
                p_cap = std::max(p_cap, n_image[b]);
                pull_cap = std::max<int>(pull_cap, (int)pulls[b].size());
            }
            g.p_cap = p_cap;
            if (p_cap >= 0xFFFF) { ibis_b200_destroy(c); return fail(nullptr, IBIS_B200_EINVAL, "tile has too many side records"); }
            // a loader lane issues one run: tiles that need more than a warp's worth keep to the gather path
            if (pull_cap > 30) c->use_sides = false;
            g.pull_cap = pull_cap;
            g.pull_stride = 8 + 3 * pull_cap;
            std::vector<int> tpull((size_t)g.pull_stride * std::max(n_tiles, 1), 0);
            for (int b = 0; b < n_tiles; b++) {
                int* r = &tpull[(size_t)g.pull_stride * b];
                r[0] = pinfo[2 * b]; r[1] = pinfo[2 * b + 1];
                r[2] = (home0[b + 1] - home0[b]) + (foreign_off[b + 1] - foreign_off[b]);
                r[3] = (int)pulls[b].size(); r[4] = n_image[b];
                for (size_t k = 0; k < pulls[b].size(); k++) {
                    r[8 + 3 * k] = pulls[b][k].src; r[9 + 3 * k] = pulls[b][k].count; r[10 + 3 * k] = pulls[b][k].dst;
                }
            }
            std::vector<std::vector<Pull>>().swap(pulls);
            g.grec_stride = (int)pad_to(grec_bytes(NFr, T, halo_cap), 128);
            g.freca_stride = (int)pad_to(freca_bytes(capp), 128);
            g.frec_stride = (int)pad_to(frec_bytes(T), 128);
            std::vector<unsigned char> grec((size_t)g.grec_stride * std::max(nt, 1), 0);
#pragma omp parallel for schedule(dynamic, 64)
            for (int b = 0; b < nt; b++) {
                unsigned char* rec = &grec[(size_t)g.grec_stride * b];
                unsigned short* nbr = reinterpret_cast<unsigned short*>(rec);
                unsigned short* stg = nbr + (size_t)NFr * T;
                int* hdr = reinterpret_cast<int*>(stg + (size_t)NFr * T);
                int* hl = hdr + 8;
                int has_ghost = 0;
                const int64_t c0 = TL.c0[b], c1 = TL.c0[b + 1];
                const std::vector<int>& h = halos[b];
                for (int i = 0; i < T; i++)
                    for (int j = 0; j < NFr; j++) {
                        const int64_t cell = c0 + i;
                        unsigned short e = (unsigned short)i, st = 0xFFFF;
                        if (cell < c1 && j < max_faces) st = cf_stage[(size_t)j * cs + cell];
                        if (st != 0xFFFF) {
                            const int nb = cf_nbr[(size_t)j * cs + cell];
                            int local;
                            if (nb >= c0 && nb < c1) local = (int)(nb - c0);
                            else local = T + (int)(std::lower_bound(h.begin(), h.end(), nb) - h.begin());
                            e = (unsigned short)local;
                            if (cf_out[(size_t)j * cs + cell] > 0) e |= NBR_LEFT_BIT;
                            if (nb >= gd->n_cells) { e |= NBR_GHOST_BIT; has_ghost = 1; }
                        }
                        nbr[(size_t)j * T + i] = e;
                        stg[(size_t)j * T + i] = st;
                    }
                hdr[0] = (int)h.size(); hdr[1] = TL.nR[b]; hdr[2] = TL.n_own_left(b); hdr[3] = n_stage[b];
                hdr[4] = has_ghost; hdr[5] = pinfo[2 * b]; hdr[6] = hdr[7] = 0;
                for (size_t k = 0; k < h.size(); k++) hl[k] = h[k];
                for (int k = (int)h.size(); k < halo_cap; k++) hl[k] = h.empty() ? (int)c0 : h[0];   // padding: any valid cell
            }
            std::vector<std::vector<int>>().swap(halos);
            std::vector<unsigned char> freca((size_t)g.freca_stride * std::max(n_tiles, 1), 0);
            std::vector<unsigned char> frec((size_t)g.frec_stride * std::max(n_tiles, 1), 0);
#pragma omp parallel for schedule(static)
            for (int b = 0; b < n_tiles; b++) {
                unsigned char* ra = &freca[(size_t)g.freca_stride * b];
                unsigned short* li = reinterpret_cast<unsigned short*>(ra);
                unsigned short* ri = li + capp;
                double* fr6 = reinterpret_cast<double*>(ri + capp);
                const int ns = (home0[b + 1] - home0[b]) + (foreign_off[b + 1] - foreign_off[b]);
                for (int sl = 0; sl < ns; sl++) {
                    li[sl] = (unsigned short)lidx[(size_t)capp * b + sl];
                    ri[sl] = (unsigned short)ridx[(size_t)capp * b + sl];
                    const int f = face_of_slot(b, sl);
                    for (int d = 0; d < 3; d++) {
                        fr6[(size_t)sl * 6 + d] = gd->face_tan1[3 * (size_t)f + d];
                        fr6[(size_t)sl * 6 + 3 + d] = gd->face_tan2[3 * (size_t)f + d];
                    }
                }
                unsigned char* rec = &frec[(size_t)g.frec_stride * b];
                double* rv = reinterpret_cast<double*>(rec);
                double* sa = rv + T;
                unsigned short* sl = reinterpret_cast<unsigned short*>(sa + (size_t)MAXF * T);
                const int64_t c0 = TL.c0[b], c1 = TL.c0[b + 1];
                for (int i = 0; i < T; i++) {
                    const int64_t cell = c0 + i;
                    rv[i] = cell < c1 ? vol[cell] : 1.0;
                    for (int j = 0; j < MAXF; j++) {
                        unsigned short sv = 0xFFFF;
                        double a = 0.0;
                        if (cell < c1 && j < max_faces && cf_face[(size_t)j * cs + cell] >= 0) {
                            sv = cf_slot[(size_t)j * cs + cell];
                            a = -fa[cf_face[(size_t)j * cs + cell]] * (double)cf_out[(size_t)j * cs + cell];
                        }
                        sl[(size_t)j * T + i] = sv;
                        sa[(size_t)j * T + i] = a;
                    }
                }
            }
            unsigned char* du8 = nullptr;
            rc = upload_vec(c, grec, &du8);
            if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
            g.t_grec = du8;
            rc = upload_vec(c, freca, &du8);
            if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
            g.t_freca = du8;
            rc = upload_vec(c, frec, &du8);
            if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
            g.t_frec = du8;
            rc = upload_vec(c, pinfo, &di);
            if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
            g.t_pinfo = di;
            rc = upload_vec(c, tpull, &di);
            if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "grid upload failed"); }
            g.t_pull = di;
          }
        }
#undef UP2
    }
    timer.lap("tile tables + upload");
    if (timer.on)
        fprintf(stderr, "[ibis_b200 create] tile %d, tiles %d (+%d halo ring), slots/tile <= %d, staged sides/tile <= %d, image records/tile <= %d "
                        "(%lld stored in all), halo list <= %d, pulled runs/tile <= %d, records %d + %d + %d B/tile\n", c->g.tile, c->g.n_tiles,
                c->g.n_tiles_all - c->g.n_tiles, c->g.tile_cap, c->g.stage_cap, c->g.p_cap, (long long)c->p_total, c->g.halo_cap,
                c->g.pull_cap, c->g.grec_stride, c->g.freca_stride, c->g.frec_stride);
    // WLS r-factors on the device (GridBlock::allocate_gradient_weights, grid.h:499-503)
    double* r = nullptr;
    rc = dev_alloc_zero<double>(c, 6 * cs, &r);
    if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "out of device memory"); }
    g.wls_r = r;
    double *cfw = nullptr, *cfd = nullptr;
    rc = dev_alloc_zero<double>(c, (int64_t)max_faces * 3 * cs, &cfw);
    if (!rc) rc = dev_alloc_zero<double>(c, (int64_t)MAXF * 3 * cs, &cfd);           // MAXF wide, see GridDev
    if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "out of device memory"); }
    g.cf_w = cfw;
    g.cf_d = cfd;
    launch_wls_r(c, r);
    launch_cell_static(c, cfw, cfd);
    if (g.t_grec) {
        // the tile-major copy of the WLS tables the pipelined kernel fetches with one bulk copy per tile
        double* twd = nullptr;
        rc = dev_alloc_zero<double>(c, (int64_t)std::max(g.n_tiles_all, 1) * 2 * g.pipe_nf * 3 * g.tile, &twd);
        if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "out of device memory"); }
        g.t_wd = twd;
        launch_pack_pipe_tables(c, twd);
    }

    // reductions
    c->n_red_blocks = 148 * 8;
    rc = dev_alloc_zero<double>(c, (int64_t)c->n_red_blocks * 8, &c->d_partials);
    if (!rc) rc = dev_alloc_zero<double>(c, N_SCALARS, &c->d_scalars);
    if (!rc) {                                      // "the previous step" of the first asynchronous rk_step
        const double big = std::numeric_limits<double>::max();
        rc = gpu::h2d(c->d_scalars + SC_DT, &big, sizeof big, c->stream) != 0 ? IBIS_B200_ECUDA : 0;
        if (!rc) rc = gpu::stream_sync(c->stream) != 0 ? IBIS_B200_ECUDA : 0;
    }
    if (rc) { ibis_b200_destroy(c); return fail(nullptr, rc, "out of device memory"); }
    void* hp = nullptr;
    if (gpu::host_malloc(&hp, N_SCALARS * sizeof(double)) != 0) {
        ibis_b200_destroy(c);
        return fail(nullptr, IBIS_B200_ENOMEM, "pinned allocation failed");
    }
    c->h_scalars = static_cast<double*>(hp);
    rc = dev_alloc_zero<unsigned int>(c, 8, &c->d_pipe_dbg);
    if (rc || gpu::host_malloc(&hp, 8 * sizeof(unsigned int)) != 0) {
        ibis_b200_destroy(c);
        return fail(nullptr, IBIS_B200_ENOMEM, "allocation failed");
    }
    c->h_pipe_dbg = static_cast<unsigned int*>(hp);
    memset(c->h_pipe_dbg, 0, 8 * sizeof(unsigned int));
    rc = gpu::stream_sync(c->stream);
    if (rc != 0 || !c->last_error.empty()) {
        std::string msg = c->last_error.empty() ? gpu::error_string(rc) : c->last_error;
        ibis_b200_destroy(c);
        return fail(nullptr, IBIS_B200_ECUDA, msg);
    }
    timer.lap("device setup kernels");
    *out = c;
    return 0;
}

static void ring_destroy(ibis_b200_ctx* c);      // the snapshot ring, below
int ibis_b200_destroy(ibis_b200_ctx* c) {
    if (!c) return 0;
    gpu::set_device(c->device);
    if (c->stream) gpu::stream_sync(c->stream);
    comm::destroy(c);
    for (auto& b : c->bufs)
        if (b.kind != BUF_FREE && b.d) gpu::dev_free(b.d);
    for (void* p : c->owned_allocs) gpu::dev_free(p);
    if (c->d_stage) gpu::dev_free(c->d_stage);
    if (c->h_scalars) gpu::host_free(c->h_scalars);
    if (c->h_pipe_dbg) gpu::host_free(c->h_pipe_dbg);
    ring_destroy(c);
    if (c->stream) gpu::stream_destroy(c->stream);
    delete c;
    return 0;
}

int ibis_b200_sync(ibis_b200_ctx* c) {
    int rc_ = sync_checked(c);
    if (rc_) return rc_;
    if (!c->last_error.empty()) return IBIS_B200_ECUDA;
    return 0;
}

int ibis_b200_sizes(const ibis_b200_ctx* c, int32_t* dim, int64_t* n_cells, int64_t* n_ghost, int64_t* n_faces) {
    if (!c) return IBIS_B200_EINVAL;
    if (dim) *dim = c->dim;
    if (n_cells) *n_cells = c->n_cells;
    if (n_ghost) *n_ghost = c->n_ghost;
    if (n_faces) *n_faces = c->n_faces;
    return 0;
}
int64_t ibis_b200_launch_count(const ibis_b200_ctx* c) { return c->launches; }
int ibis_b200_profile_enable(ibis_b200_ctx* c, int on) { c->prof_on = on != 0; return 0; }
int ibis_b200_profile_read(ibis_b200_ctx* c, double ms[4], int64_t count[4]) { return prof_read(c, ms, count); }
void* ibis_b200_stream(ibis_b200_ctx* c) { return (void*)c->stream; }

// ---- buffers ----
int ibis_b200_fs_create(ibis_b200_ctx* c, int is_dual, int* h) {
    return new_buffer(c, BUF_FS, is_dual != 0, 7, c->n_total, c->cstride, h);
}
int ibis_b200_cq_create(ibis_b200_ctx* c, int is_dual, int* h) {
    return new_buffer(c, BUF_CQ, is_dual != 0, 5, c->n_total, c->cstride, h);
}
int ibis_b200_vec_create(ibis_b200_ctx* c, int64_t n, int* h) {
    return new_buffer(c, BUF_VEC, false, 1, n, pad_to(std::max<int64_t>(n, 1), 32), h);
}
int ibis_b200_basis_create(ibis_b200_ctx* c, int64_t n, int n_cols, int* h) {
    if (n_cols + 2 > 1024) return fail(c, IBIS_B200_EINVAL, "too many Krylov vectors");
    return new_buffer(c, BUF_BASIS, false, n_cols, n, pad_to(std::max<int64_t>(n, 1), 32), h);
}
int ibis_b200_free(ibis_b200_ctx* c, int h) {
    if (h <= 0 || h >= (int)c->bufs.size() || c->bufs[h].kind == BUF_FREE)
        return fail(c, IBIS_B200_EHANDLE, "bad handle");
    gpu::stream_sync(c->stream);
    gpu::dev_free(c->bufs[h].d);
    c->bufs[h] = Buffer();
    return 0;
}

static const int FS_MAP[7] = {0, 1, 2, 3, 4, 5, 6};
static const int CQ_MAP3[5] = {0, 1, 2, 3, 4};
static const int CQ_MAP2[4] = {0, 1, 2, 4};

static int upload_rows(ibis_b200_ctx* c, Buffer* b, const int* map, int ncomp_ref, const double* re_, const double* du_) {
    if (!re_) return fail(c, IBIS_B200_EINVAL, "null host array");
    const int64_t n = b->n, elems = n * ncomp_ref;
    int rc = ensure_stage(c, (b->is_dual ? 2 : 1) * elems);
    if (rc) return rc;
    GPU_TRY(c, gpu::h2d(c->d_stage, re_, elems * sizeof(double), c->stream));
    const double* d_du = nullptr;
    if (b->is_dual && du_) {
        GPU_TRY(c, gpu::h2d(c->d_stage + elems, du_, elems * sizeof(double), c->stream));
        d_du = c->d_stage + elems;
    }
    if (b->is_dual) {
        dual* p[7];
        comp_ptrs<dual>(b, p);
        launch_pack_in<dual>(c, c->d_stage, d_du, p, map, ncomp_ref, n);
    } else {
        double* p[7];
        comp_ptrs<double>(b, p);
        launch_pack_in<double>(c, c->d_stage, nullptr, p, map, ncomp_ref, n);
    }
    GPU_TRY(c, gpu::stream_sync(c->stream));    // the host arrays may be pageable
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

static int download_rows(ibis_b200_ctx* c, Buffer* b, const int* map, int ncomp_ref, double* re_, double* du_) {
    const int64_t n = b->n, elems = n * ncomp_ref;
    int rc = ensure_stage(c, (b->is_dual ? 2 : 1) * elems);
    if (rc) return rc;
    if (b->is_dual) {
        dual* p[7];
        comp_ptrs<dual>(b, p);
        launch_pack_out<dual>(c, p, map, ncomp_ref, n, c->d_stage, c->d_stage + elems);
    } else {
        double* p[7];
        comp_ptrs<double>(b, p);
        launch_pack_out<double>(c, p, map, ncomp_ref, n, c->d_stage, nullptr);
    }
    if (re_) GPU_TRY(c, gpu::d2h(re_, c->d_stage, elems * sizeof(double), c->stream));
    if (du_) {
        if (b->is_dual) GPU_TRY(c, gpu::d2h(du_, c->d_stage + elems, elems * sizeof(double), c->stream));
    }
    GPU_TRY(c, gpu::stream_sync(c->stream));
    if (du_ && !b->is_dual) memset(du_, 0, elems * sizeof(double));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_fs_upload(ibis_b200_ctx* c, int h, const double* re_, const double* du_) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    return upload_rows(c, b, FS_MAP, 7, re_, du_);
}
int ibis_b200_fs_download(ibis_b200_ctx* c, int h, double* re_, double* du_) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    return download_rows(c, b, FS_MAP, 7, re_, du_);
}
int ibis_b200_flow_read_native(ibis_b200_ctx* c, int h, const double* T, const double* p, const double* vx,
                               const double* vy, const double* vz) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    if (!T || !p || !vx || !vy || (c->dim == 3 && !vz)) return fail(c, IBIS_B200_EINVAL, "null host array");
    const int64_t n = c->n_cells;
    const int nv = vz ? 5 : 4;
    int rc = ensure_stage(c, 5 * n);
    if (rc) return rc;
    const double* src[5] = {T, p, vx, vy, vz};
    for (int q = 0; q < nv; q++) GPU_TRY(c, gpu::h2d(c->d_stage + q * n, src[q], n * sizeof(double), c->stream));
    if (b->is_dual) { dual* f[7]; comp_ptrs<dual>(b, f); launch_flow_read<dual>(c, c->d_stage, n, vz != nullptr, f); }
    else { double* f[7]; comp_ptrs<double>(b, f); launch_flow_read<double>(c, c->d_stage, n, vz != nullptr, f); }
    GPU_TRY(c, gpu::stream_sync(c->stream));    // the host arrays may be pageable
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_flow_write_native(ibis_b200_ctx* c, int h, double* T, double* p, double* vx, double* vy, double* vz) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    if (!T || !p || !vx || !vy || (c->dim == 3 && !vz)) return fail(c, IBIS_B200_EINVAL, "null host array");
    const int64_t n = c->n_cells;
    const int nv = vz ? 5 : 4;
    int rc = ensure_stage(c, 5 * n);
    if (rc) return rc;
    if (b->is_dual) { dual* f[7]; comp_ptrs<dual>(b, f); launch_flow_write<dual>(c, f, c->d_stage, n, vz != nullptr); }
    else { double* f[7]; comp_ptrs<double>(b, f); launch_flow_write<double>(c, f, c->d_stage, n, vz != nullptr); }
    double* dst[5] = {T, p, vx, vy, vz};
    for (int q = 0; q < nv; q++) GPU_TRY(c, gpu::d2h(dst[q], c->d_stage + q * n, n * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
// ---- snapshot ring: native_binary snapshots travel while the device computes ------------------------------------
static int ring_init(ibis_b200_ctx* c) {
    ibis_b200_ctx::SnapshotRing& r = c->ring;
    if (r.ready) return 0;
    const size_t bytes = (size_t)5 * (size_t)std::max<int64_t>(c->n_cells, 1) * sizeof(double);
    GPU_TRY(c, gpu::stream_create(&r.up));
    GPU_TRY(c, gpu::stream_create(&r.down));
    for (int k = 0; k < 2; k++) {
        void* p = nullptr;
        if (gpu::dev_malloc(&p, bytes) != 0) return fail(c, IBIS_B200_ENOMEM, "out of device memory (snapshot ring)");
        r.d_up[k] = static_cast<double*>(p);
        if (gpu::dev_malloc(&p, bytes) != 0) return fail(c, IBIS_B200_ENOMEM, "out of device memory (snapshot ring)");
        r.d_down[k] = static_cast<double*>(p);
        GPU_TRY(c, gpu::event_create(&r.uploaded[k]));
        GPU_TRY(c, gpu::event_create(&r.unpacked[k]));
        GPU_TRY(c, gpu::event_create(&r.packed[k]));
        GPU_TRY(c, gpu::event_create(&r.downloaded[k]));
    }
    r.ready = true;
    return 0;
}
static void ring_destroy(ibis_b200_ctx* c) {
    ibis_b200_ctx::SnapshotRing& r = c->ring;
    if (!r.ready) return;
    for (int k = 0; k < 2; k++) {
        gpu::dev_free(r.d_up[k]);
        gpu::dev_free(r.d_down[k]);
        gpu::event_destroy(r.uploaded[k]);
        gpu::event_destroy(r.unpacked[k]);
        gpu::event_destroy(r.packed[k]);
        gpu::event_destroy(r.downloaded[k]);
    }
    gpu::stream_destroy(r.up);
    gpu::stream_destroy(r.down);
    r.ready = false;
}
int ibis_b200_pinned_alloc(ibis_b200_ctx* c, int64_t bytes, void** out) {
    if (!c || !out || bytes < 0) return fail(c, IBIS_B200_EINVAL, "null argument");
    void* p = nullptr;
    if (gpu::host_malloc(&p, (size_t)std::max<int64_t>(bytes, 8)) != 0) return fail(c, IBIS_B200_ENOMEM, "out of page-locked host memory");
    *out = p;
    return 0;
}
int ibis_b200_pinned_free(ibis_b200_ctx* c, void* p) {
    (void)c;
    if (p) gpu::host_free(p);
    return 0;
}
int ibis_b200_snapshot_upload_begin(ibis_b200_ctx* c, int slot, const double* T, const double* p, const double* vx,
                                    const double* vy, const double* vz) {
    if (slot < 0 || slot > 1) return fail(c, IBIS_B200_EINVAL, "snapshot slot must be 0 or 1");
    if (!T || !p || !vx || !vy || (c->dim == 3 && !vz)) return fail(c, IBIS_B200_EINVAL, "null host array");
    int rc = ring_init(c);
    if (rc) return rc;
    ibis_b200_ctx::SnapshotRing& r = c->ring;
    const int64_t n = c->n_cells;
    // the slot is free once the compute stream has unpacked what it held
    if (r.up_used[slot]) GPU_TRY(c, gpu::stream_wait_event(r.up, r.unpacked[slot]));
    const double* src[5] = {T, p, vx, vy, vz};
    for (int q = 0; q < (vz ? 5 : 4); q++) GPU_TRY(c, gpu::h2d(r.d_up[slot] + q * n, src[q], n * sizeof(double), r.up));
    GPU_TRY(c, gpu::event_record(r.uploaded[slot], r.up));
    return 0;
}
int ibis_b200_snapshot_upload_end(ibis_b200_ctx* c, int slot, int h) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    if (slot < 0 || slot > 1 || !c->ring.ready) return fail(c, IBIS_B200_EINVAL, "snapshot_upload_end without snapshot_upload_begin");
    ibis_b200_ctx::SnapshotRing& r = c->ring;
    const int64_t n = c->n_cells;
    GPU_TRY(c, gpu::stream_wait_event(c->stream, r.uploaded[slot]));
    if (b->is_dual) { dual* f[7]; comp_ptrs<dual>(b, f); launch_flow_read<dual>(c, r.d_up[slot], n, c->dim == 3, f); }
    else { double* f[7]; comp_ptrs<double>(b, f); launch_flow_read<double>(c, r.d_up[slot], n, c->dim == 3, f); }
    GPU_TRY(c, gpu::event_record(r.unpacked[slot], c->stream));
    r.up_used[slot] = true;
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_snapshot_download_begin(ibis_b200_ctx* c, int slot, int h, double* T, double* p, double* vx, double* vy, double* vz) {
    Buffer* b = get_buf(c, h, BUF_FS);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a FlowStates handle");
    if (slot < 0 || slot > 1) return fail(c, IBIS_B200_EINVAL, "snapshot slot must be 0 or 1");
    if (!T || !p || !vx || !vy || (c->dim == 3 && !vz)) return fail(c, IBIS_B200_EINVAL, "null host array");
    int rc = ring_init(c);
    if (rc) return rc;
    ibis_b200_ctx::SnapshotRing& r = c->ring;
    const int64_t n = c->n_cells;
    // the slot is free once its previous download has left the device
    if (r.down_used[slot]) GPU_TRY(c, gpu::stream_wait_event(c->stream, r.downloaded[slot]));
    if (b->is_dual) { dual* f[7]; comp_ptrs<dual>(b, f); launch_flow_write<dual>(c, f, r.d_down[slot], n, c->dim == 3); }
    else { double* f[7]; comp_ptrs<double>(b, f); launch_flow_write<double>(c, f, r.d_down[slot], n, c->dim == 3); }
    GPU_TRY(c, gpu::event_record(r.packed[slot], c->stream));
    GPU_TRY(c, gpu::stream_wait_event(r.down, r.packed[slot]));
    double* dst[5] = {T, p, vx, vy, vz};
    for (int q = 0; q < (vz ? 5 : 4); q++) GPU_TRY(c, gpu::d2h(dst[q], r.d_down[slot] + q * n, n * sizeof(double), r.down));
    GPU_TRY(c, gpu::event_record(r.downloaded[slot], r.down));
    r.down_used[slot] = true;
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_snapshot_download_end(ibis_b200_ctx* c, int slot) {
    if (slot < 0 || slot > 1 || !c->ring.ready || !c->ring.down_used[slot])
        return fail(c, IBIS_B200_EINVAL, "snapshot_download_end without snapshot_download_begin");
    GPU_TRY(c, gpu::event_sync(c->ring.downloaded[slot]));
    return check_pipe_stall(c);
}

int ibis_b200_cq_upload(ibis_b200_ctx* c, int h, const double* re_, const double* du_) {
    Buffer* b = get_buf(c, h, BUF_CQ);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a ConservedQuantities handle");
    return upload_rows(c, b, c->dim == 3 ? CQ_MAP3 : CQ_MAP2, c->n_cons, re_, du_);
}
int ibis_b200_cq_download(ibis_b200_ctx* c, int h, double* re_, double* du_) {
    Buffer* b = get_buf(c, h, BUF_CQ);
    if (!b) return fail(c, IBIS_B200_EHANDLE, "not a ConservedQuantities handle");
    return download_rows(c, b, c->dim == 3 ? CQ_MAP3 : CQ_MAP2, c->n_cons, re_, du_);
}
int ibis_b200_vec_upload(ibis_b200_ctx* c, int h, const double* x) {
    Buffer* b = get_buf(c, h, BUF_VEC);
    if (!b || !x) return fail(c, IBIS_B200_EHANDLE, "not a vector handle");
    GPU_TRY(c, gpu::h2d(b->d, x, b->n * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return 0;
}
int ibis_b200_vec_download(ibis_b200_ctx* c, int h, double* x) {
    Buffer* b = get_buf(c, h, BUF_VEC);
    if (!b || !x) return fail(c, IBIS_B200_EHANDLE, "not a vector handle");
    GPU_TRY(c, gpu::d2h(x, b->d, b->n * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_copy(ibis_b200_ctx* c, int dst, int src) {
    if (dst <= 0 || src <= 0 || dst >= (int)c->bufs.size() || src >= (int)c->bufs.size())
        return fail(c, IBIS_B200_EHANDLE, "bad handle");
    Buffer &d = c->bufs[dst], &s = c->bufs[src];
    if (d.kind != s.kind || d.kind == BUF_FREE || d.is_dual != s.is_dual || d.n != s.n || d.ncomp != s.ncomp)
        return fail(c, IBIS_B200_EHANDLE, "copy between different buffer kinds");
    GPU_TRY(c, gpu::d2d(d.d, s.d, (size_t)s.ncomp * s.stride * (s.is_dual ? 16 : 8), c->stream));
    return 0;
}
int ibis_b200_zero(ibis_b200_ctx* c, int h) {
    if (h <= 0 || h >= (int)c->bufs.size() || c->bufs[h].kind == BUF_FREE) return fail(c, IBIS_B200_EHANDLE, "bad handle");
    Buffer& b = c->bufs[h];
    GPU_TRY(c, gpu::memset_async(b.d, 0, (size_t)b.ncomp * b.stride * (b.is_dual ? 16 : 8), c->stream));
    return 0;
}

// ---- the residual ----
int ibis_b200_compute_dudt(ibis_b200_ctx* c, int fs, int dudt, int allow_reconstruction) {
    Buffer* f = get_buf(c, fs, BUF_FS);
    Buffer* d = get_buf(c, dudt, BUF_CQ);
    if (!f || !d) return fail(c, IBIS_B200_EHANDLE, "compute_dudt: bad handle");
    if (f->is_dual != d->is_dual) return fail(c, IBIS_B200_EINVAL, "compute_dudt: fs and dudt differ in scalar type");
    if (f->is_dual) return compute_dudt_impl<dual>(c, f, d, allow_reconstruction != 0);
    return compute_dudt_impl<double>(c, f, d, allow_reconstruction != 0);
}

static int estimate_dt_device(ibis_b200_ctx* c, Buffer* f) {
    if (c->cfg.viscous && !c->face_T) {
        int rc = ensure_fv_temporaries<double>(c);
        if (rc) return rc;
    }
    if (f->is_dual) {
        FvArgs<dual> A = make_args<dual>(c, f, nullptr, true);
        launch_local_dt_min(c, A, c->d_scalars + SC_DT_STABLE);
    } else {
        FvArgs<double> A = make_args<double>(c, f, nullptr, true);
        launch_local_dt_min(c, A, c->d_scalars + SC_DT_STABLE);
    }
    return comm::allreduce(c, c->d_scalars + SC_DT_STABLE, 1, comm::OP_MIN);
}

int ibis_b200_estimate_dt(ibis_b200_ctx* c, int fs, double* dt) {
    Buffer* f = get_buf(c, fs, BUF_FS);
    if (!f || !dt) return fail(c, IBIS_B200_EHANDLE, "estimate_dt: bad handle");
    int rc = estimate_dt_device(c, f);
    if (rc) return rc;
    return fetch_scalars(c, SC_DT_STABLE, 1, dt);
}

int ibis_b200_count_bad_cells(ibis_b200_ctx* c, int fs, uint64_t* n_bad) {
    Buffer* f = get_buf(c, fs, BUF_FS);
    if (!f || !n_bad) return fail(c, IBIS_B200_EHANDLE, "count_bad_cells: bad handle");
    if (f->is_dual) { dual* p[7]; comp_ptrs<dual>(f, p); launch_count_bad<dual>(c, p, c->n_owned, c->d_scalars + SC_BAD); }
    else { double* p[7]; comp_ptrs<double>(f, p); launch_count_bad<double>(c, p, c->n_owned, c->d_scalars + SC_BAD); }
    int rc = comm::allreduce(c, c->d_scalars + SC_BAD, 1, comm::OP_SUM);
    if (rc) return rc;
    double v = 0;
    rc = fetch_scalars(c, SC_BAD, 1, &v);
    *n_bad = (uint64_t)v;
    return rc;
}

int ibis_b200_cons_to_prim(ibis_b200_ctx* c, int cq, int fs) {
    Buffer* q = get_buf(c, cq, BUF_CQ);
    Buffer* f = get_buf(c, fs, BUF_FS);
    if (!q || !f || q->is_dual != f->is_dual) return fail(c, IBIS_B200_EHANDLE, "cons_to_prim: bad handles");
    if (f->is_dual) { dual *a[7], *b[7]; comp_ptrs<dual>(q, a); comp_ptrs<dual>(f, b); launch_cons_to_prim<dual>(c, a, b, c->n_total); }
    else { double *a[7], *b[7]; comp_ptrs<double>(q, a); comp_ptrs<double>(f, b); launch_cons_to_prim<double>(c, a, b, c->n_total); }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_prim_to_cons(ibis_b200_ctx* c, int cq, int fs) {
    Buffer* q = get_buf(c, cq, BUF_CQ);
    Buffer* f = get_buf(c, fs, BUF_FS);
    if (!q || !f || q->is_dual != f->is_dual) return fail(c, IBIS_B200_EHANDLE, "prim_to_cons: bad handles");
    if (f->is_dual) { dual *a[7], *b[7]; comp_ptrs<dual>(q, a); comp_ptrs<dual>(f, b); launch_prim_to_cons<dual>(c, a, b, c->n_total); }
    else { double *a[7], *b[7]; comp_ptrs<double>(q, a); comp_ptrs<double>(f, b); launch_prim_to_cons<double>(c, a, b, c->n_total); }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

static int apply_dt(ibis_b200_ctx* c, Buffer* u0, Buffer* u1, Buffer* k, double dt, const double* d_dt, double scale) {
    if (u0->is_dual) { dual *a[7], *b[7], *kk[7]; comp_ptrs<dual>(u0, a); comp_ptrs<dual>(u1, b); comp_ptrs<dual>(k, kk);
        launch_apply_dt<dual>(c, a, b, kk, dt, d_dt, scale, c->n_total); }
    else { double *a[7], *b[7], *kk[7]; comp_ptrs<double>(u0, a); comp_ptrs<double>(u1, b); comp_ptrs<double>(k, kk);
        launch_apply_dt<double>(c, a, b, kk, dt, d_dt, scale, c->n_total); }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_apply_time_derivative(ibis_b200_ctx* c, int u0, int u1, int dudt, double dt) {
    Buffer *a = get_buf(c, u0, BUF_CQ), *b = get_buf(c, u1, BUF_CQ), *k = get_buf(c, dudt, BUF_CQ);
    if (!a || !b || !k || a->is_dual != b->is_dual || a->is_dual != k->is_dual)
        return fail(c, IBIS_B200_EHANDLE, "apply_time_derivative: bad handles");
    return apply_dt(c, a, b, k, dt, nullptr, 1.0);
}

static int l2_norms_device(ibis_b200_ctx* c, Buffer* q) {
    if (q->is_dual) { dual* p[7]; comp_ptrs<dual>(q, p); launch_l2_norms<dual>(c, p, c->n_owned, c->d_scalars + SC_L2); }
    else { double* p[7]; comp_ptrs<double>(q, p); launch_l2_norms<double>(c, p, c->n_owned, c->d_scalars + SC_L2); }
    return comm::allreduce(c, c->d_scalars + SC_L2, 6, comm::OP_SUM);
}

int ibis_b200_l2_norms(ibis_b200_ctx* c, int cq, double out[6]) {
    Buffer* q = get_buf(c, cq, BUF_CQ);
    if (!q || !out) return fail(c, IBIS_B200_EHANDLE, "l2_norms: bad handle");
    // the reference sums over all n_total rows; ghost (and halo) rows of a dudt are zero
    int rc = l2_norms_device(c, q);
    if (rc) return rc;
    rc = fetch_scalars(c, SC_L2, 6, out);
    for (int i = 0; i < 6; i++) out[i] = sqrt(out[i]);
    return rc;
}

int ibis_b200_set_face_vertices(ibis_b200_ctx* c, const int64_t* face_vertices, int64_t n_vertices) {
    if (!c || !face_vertices || n_vertices <= 0) return fail(c, IBIS_B200_EINVAL, "null argument");
    if (n_vertices >= (int64_t)1 << 31) return fail(c, IBIS_B200_EINVAL, "too many vertices for 32-bit ids");
    std::vector<int> fv((size_t)4 * c->fstride, -1);
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (int64_t f = 0; f < c->n_faces; f++) {
        const int64_t g = c->face_ref2int[f];
        for (int k = 0; k < 4; k++) {
            const int64_t v = face_vertices[4 * f + k];
            if (v >= n_vertices || (k < 2 && v < 0)) bad = true;
            fv[(size_t)k * c->fstride + g] = v < 0 ? -1 : (int)v;
        }
    }
    if (bad) return fail(c, IBIS_B200_EINVAL, "interface refers to a vertex outside the vertex list");
    if (c->d_fverts) return fail(c, IBIS_B200_EINVAL, "face vertices already set");
    int rc = upload_vec(c, fv, &c->d_fverts);
    if (!rc) rc = dev_alloc_zero<double>(c, 3 * n_vertices, &c->d_vertex_vel);
    if (!rc) rc = dev_alloc_zero<double>(c, 3 * c->fstride, &c->d_face_vel);
    if (rc) return fail(c, rc, "out of device memory (moving grid)");
    c->n_vertices = n_vertices;
    return 0;
}

int ibis_b200_set_vertex_velocities(ibis_b200_ctx* c, const double* vertex_vel) {
    if (!c || !vertex_vel) return fail(c, IBIS_B200_EINVAL, "null argument");
    if (!c->d_fverts) return fail(c, IBIS_B200_EINVAL, "ibis_b200_set_face_vertices has not been called");
    GPU_TRY(c, gpu::h2d(c->d_vertex_vel, vertex_vel, (size_t)3 * c->n_vertices * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));      // vertex_vel is the caller's
    c->have_vertex_vel = true;
    return 0;
}

int ibis_b200_set_cell_vertices(ibis_b200_ctx* c, const int32_t* cell_types, const int64_t* cell_vertices) {
    if (!c || !cell_types || !cell_vertices) return fail(c, IBIS_B200_EINVAL, "null argument");
    if (!c->d_fverts) return fail(c, IBIS_B200_EINVAL, "ibis_b200_set_face_vertices has not been called");
    if (c->d_cverts) return fail(c, IBIS_B200_EINVAL, "cell vertices already set");
    std::vector<int> cv((size_t)8 * c->cstride, -1), ct((size_t)c->cstride, 0);
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (int64_t o = 0; o < c->n_cells; o++) {
        const int64_t i = c->ref2int.empty() ? o : c->ref2int[o];
        const int nv = gb_num_vertices(cell_types[o]);
        if (nv == 0) bad = true;
        ct[i] = cell_types[o];
        for (int k = 0; k < nv; k++) {
            const int64_t v = cell_vertices[8 * o + k];
            if (v < 0 || v >= c->n_vertices) bad = true;
            cv[(size_t)k * c->cstride + i] = (int)v;
        }
    }
    if (bad) return fail(c, IBIS_B200_EINVAL, "cell refers to a vertex outside the vertex list, or unsupported cell type");
    int rc = upload_vec(c, cv, &c->d_cverts);
    if (!rc) rc = upload_vec(c, ct, &c->d_ctype);
    if (!rc) rc = dev_alloc_zero<double>(c, 3 * c->n_vertices, &c->d_vertex_pos);
    if (rc) return fail(c, rc, "out of device memory (moving grid)");
    return 0;
}

// GridBlock::compute_geometric_data (grid.h:169-178) from the vertex positions on the device
static int recompute_geometry(ibis_b200_ctx* c) {
    GeomUpdate U;
    U.fverts = c->d_fverts;
    U.cverts = c->d_cverts;
    U.ctype = c->d_ctype;
    U.pos = c->d_vertex_pos;
    U.f_c = const_cast<double*>(c->g.f_c);
    U.f_area = const_cast<double*>(c->g.f_area);
    U.f_geom = const_cast<double*>(c->g.f_geom);
    U.cxyz = const_cast<double*>(c->g.cx);
    U.volume = const_cast<double*>(c->g.volume);
    launch_geom_update(c, U);
    launch_wls_r(c, const_cast<double*>(c->g.wls_r));                                  // compute_gradient_weights
    launch_cell_static(c, const_cast<double*>(c->g.cf_w), const_cast<double*>(c->g.cf_d));
    c->geometry_moved = true;
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_set_vertex_positions(ibis_b200_ctx* c, const double* positions) {
    if (!c || !positions) return fail(c, IBIS_B200_EINVAL, "null argument");
    if (!c->d_cverts) return fail(c, IBIS_B200_EINVAL, "ibis_b200_set_cell_vertices has not been called");
    GPU_TRY(c, gpu::h2d(c->d_vertex_pos, positions, (size_t)3 * c->n_vertices * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    const bool first = !c->have_vertex_pos;
    c->have_vertex_pos = true;
    // the first call hands over the positions the grid was built from: nothing to recompute
    return first ? 0 : recompute_geometry(c);
}

int ibis_b200_move_vertices(ibis_b200_ctx* c, double scale, int recompute) {
    if (!c) return IBIS_B200_EINVAL;
    if (!c->have_vertex_pos || !c->have_vertex_vel)
        return fail(c, IBIS_B200_EINVAL, "move_vertices needs vertex positions and velocities");
    launch_move_vertices(c, c->d_vertex_pos, c->d_vertex_vel, scale, 3 * c->n_vertices);
    c->geometry_moved = true;
    return recompute ? recompute_geometry(c) : (c->last_error.empty() ? 0 : IBIS_B200_ECUDA);
}

int ibis_b200_get_vertex_positions(ibis_b200_ctx* c, double* positions) {
    if (!c || !positions || !c->have_vertex_pos) return fail(c, IBIS_B200_EINVAL, "no vertex positions");
    GPU_TRY(c, gpu::d2h(positions, c->d_vertex_pos, (size_t)3 * c->n_vertices * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return 0;
}

int ibis_b200_compute_dudt_moving(ibis_b200_ctx* c, int fs, int cq, int dudt, int allow_reconstruction) {
    if (!c) return IBIS_B200_EINVAL;
    if (!c->have_vertex_vel) return fail(c, IBIS_B200_EINVAL, "no vertex velocities: the grid is not moving");
    if (c->n_ranks > 1) return fail(c, IBIS_B200_EINVAL, "moving grids are single-block");
    Buffer* F = get_buf(c, fs, BUF_FS);
    Buffer* Q = get_buf(c, cq, BUF_CQ);
    Buffer* D = get_buf(c, dudt, BUF_CQ);
    if (!F || !Q || !D) return fail(c, IBIS_B200_EHANDLE, "compute_dudt_moving: bad handle");
    if (F->is_dual != D->is_dual || F->is_dual != Q->is_dual)
        return fail(c, IBIS_B200_EINVAL, "compute_dudt_moving: real/dual mismatch");
    return F->is_dual ? compute_dudt_impl<dual>(c, F, D, allow_reconstruction != 0, nullptr, Q)
                      : compute_dudt_impl<double>(c, F, D, allow_reconstruction != 0, nullptr, Q);
}

int ibis_b200_debug_face_velocities(ibis_b200_ctx* c, double* face_vel) {
    if (!c || !face_vel || !c->d_face_vel) return fail(c, IBIS_B200_EINVAL, "no interface velocities");
    std::vector<double> h((size_t)3 * c->fstride);
    GPU_TRY(c, gpu::d2h(h.data(), c->d_face_vel, h.size() * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    for (int64_t f = 0; f < c->n_faces; f++)
        for (int k = 0; k < 3; k++) face_vel[3 * f + k] = h[(size_t)k * c->fstride + c->face_ref2int[f]];
    return 0;
}

int ibis_b200_debug_dual_ops(ibis_b200_ctx* c, const double* in5, double* out44) {
    if (!c || !in5 || !out44) return fail(c, IBIS_B200_EINVAL, "null argument");
    int rc = ensure_stage(c, 64);
    if (rc) return rc;
    GPU_TRY(c, gpu::h2d(c->d_stage, in5, 5 * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    launch_dual_kat(c, c->d_stage, c->d_stage + 8);
    GPU_TRY(c, gpu::d2h(out44, c->d_stage + 8, 44 * sizeof(double), c->stream));
    GPU_TRY(c, gpu::stream_sync(c->stream));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_debug_division_check(ibis_b200_ctx* c, int64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!c || !mismatches || n < 0) return fail(c, IBIS_B200_EINVAL, "null argument");
    launch_division_check(c, n, seed, c->d_scalars + SC_BAD);
    double v = 0;
    int rc = fetch_scalars(c, SC_BAD, 1, &v);
    *mismatches = (uint64_t)v;
    return rc;
}

int ibis_b200_debug_gradients(ibis_b200_ctx* c, int var, double* grad, double* limiter) {
    if (var < 0 || var >= NVARS) return fail(c, IBIS_B200_EINVAL, "no gradients available");
    if (sides_path(c)) {
        // the sides path keeps gradients and limiters in registers: evaluate them again, from the
        // FlowStates of the last real evaluation (ghost rows still hold its boundary fill), with the
        // gradient kernel of the gather path
        Buffer* F = get_buf(c, c->last_fs, BUF_FS);
        if (!F || F->is_dual) return fail(c, IBIS_B200_EINVAL, "no gradients available");
        int rc0 = ensure_fv_temporaries<double>(c, true);
        if (rc0) return rc0;
        FvArgs<double> A = make_args<double>(c, F, nullptr, true);
        launch_grad_limit(c, A, false);
    }
    if (!c->grad_r) return fail(c, IBIS_B200_EINVAL, "no gradients available");
    int rc = ensure_stage(c, 3 * c->n_cells);
    if (rc) return rc;
    if (grad) {
        launch_gather_debug(c, c->grad_r + (int64_t)var * 3 * c->cstride, 3, c->n_cells, c->d_stage);
        GPU_TRY(c, gpu::d2h(grad, c->d_stage, 3 * c->n_cells * sizeof(double), c->stream));
        GPU_TRY(c, gpu::stream_sync(c->stream));
    }
    if (limiter) {
        launch_gather_debug(c, c->lim_r + (int64_t)var * c->cstride, 1, c->n_cells, c->d_stage);
        GPU_TRY(c, gpu::d2h(limiter, c->d_stage, c->n_cells * sizeof(double), c->stream));
        GPU_TRY(c, gpu::stream_sync(c->stream));
    }
    return 0;
}

// ---- fused explicit step: RungeKutta::take_step (solvers/runge_kutta.cpp:131-203) ----
int ibis_b200_rk_step(ibis_b200_ctx* c, int fs, int cq, int n_stages, const double* a, const double* b,
                      double cfl, double dt_prev, double t_remaining, double* dt, double* dt_stable) {
    Buffer* F = get_buf(c, fs, BUF_FS);
    Buffer* U = get_buf(c, cq, BUF_CQ);
    if (!F || !U || F->is_dual || U->is_dual || n_stages < 1 || !b) return fail(c, IBIS_B200_EHANDLE, "rk_step: bad arguments");
    // k_[stages], k_tmp_, flow_tmp_ (runge_kutta.cpp:45-52).  The reference fills the vector
    // k_ with COPIES of one ConservedQuantities (runge_kutta.cpp:49-50); Kokkos Views are
    // reference-counted handles, so all its stages share ONE array.  rk_alias_stages (default)
    // reproduces that; ibis_b200_rk_distinct_stages(ctx, 1) gives every stage its own array.
    while ((int)c->rk_k_own.size() < (c->rk_alias_stages ? 1 : n_stages)) {
        int h = 0;
        int rc = ibis_b200_cq_create(c, 0, &h);
        if (rc) return rc;
        c->rk_k_own.push_back(h);
    }
    c->rk_k.assign(n_stages, c->rk_k_own[0]);
    if (!c->rk_alias_stages)
        for (int i = 0; i < n_stages; i++) c->rk_k[i] = c->rk_k_own[i];
    if (n_stages > 1 && !c->rk_ktmp) {
        int rc = ibis_b200_cq_create(c, 0, &c->rk_ktmp);
        if (!rc) rc = ibis_b200_fs_create(c, 0, &c->rk_fstmp);
        if (rc) return rc;
    }
    F = get_buf(c, fs, BUF_FS);           // bufs may have been reallocated
    U = get_buf(c, cq, BUF_CQ);
    // the dt estimate (:147) reads the same FlowStates as the first evaluation (:143): on the sides
    // path it rides along in pass F' (same frames, same cell-face order, min is exact)
    const bool dt_fused = sides_path(c) && c->cfg.reconstruction_order == 2 && !c->rk_unfused;
    if (dt_fused) launch_dt_reduce_init(c, c->d_scalars + SC_DT_STABLE);
    int rc = compute_dudt_impl<double>(c, F, &c->bufs[c->rk_k[0]], true, dt_fused ? c->d_scalars + SC_DT_STABLE : nullptr);
    if (rc) return rc;
    if (dt_fused) launch_dt_reduce_finish(c, c->d_scalars + SC_DT_STABLE);
    rc = dt_fused ? comm::allreduce(c, c->d_scalars + SC_DT_STABLE, 1, comm::OP_MIN) : estimate_dt_device(c, F);
    if (rc) return rc;
    double* d_dt = c->d_scalars + SC_DT;
    launch_rk_dt(c, c->d_scalars + SC_DT_STABLE, cfl, dt_prev, t_remaining, d_dt);
    int row = 0;    // a(i, j) = a[row(i) + j], row(i) = i(i-1)/2 for i = 1..stages-1
    const bool fused = n_stages <= RK_MAX_TERMS && !c->rk_unfused;
    double *uq[7], *fq[7], *ftq[7];
    comp_ptrs<double>(U, uq);
    comp_ptrs<double>(F, fq);
    for (int i = 1; i < n_stages; i++) {
        Buffer* ktmp = &c->bufs[c->rk_ktmp];
        Buffer* ftmp = &c->bufs[c->rk_fstmp];
        comp_ptrs<double>(ftmp, ftq);
        if (fused) {
            // k_tmp = U + dt a(i,0) k_0 [+ dt a(i,j) k_j ...] and flow_tmp = primitives(k_tmp), :155-177,
            // in one pass; k_tmp itself stays in registers
            RkTerms<double> T;
            T.n = 0;
            for (int j = 0; j < i; j++) {
                if (j >= 1 && a[row + j] < 1e-14) continue;                            // :166
                double* kq[7];
                comp_ptrs<double>(&c->bufs[c->rk_k[j]], kq);
                for (int q = 0; q < 5; q++) T.k[T.n][q] = kq[q];
                T.scale[T.n++] = a[row + j];
            }
            launch_rk_update<double>(c, uq, T, d_dt, nullptr, ftq, c->n_total);
        } else {
            rc = apply_dt(c, U, ktmp, &c->bufs[c->rk_k[0]], 0.0, d_dt, a[row + 0]);       // :155
            if (rc) return rc;
            for (int j = 1; j < i; j++) {
                if (a[row + j] < 1e-14) continue;                                          // :166
                rc = apply_dt(c, ktmp, ktmp, &c->bufs[c->rk_k[j]], 0.0, d_dt, a[row + j]); // :169
                if (rc) return rc;
            }
            double* q[7];
            comp_ptrs<double>(ktmp, q);
            launch_cons_to_prim<double>(c, q, ftq, c->n_total);                          // :177
        }
        rc = compute_dudt_impl<double>(c, ftmp, &c->bufs[c->rk_k[i]], true);          // :181
        if (rc) return rc;
        row += i;
    }
    if (fused) {
        RkTerms<double> T;                                                             // :185-193
        T.n = n_stages;
        for (int i = 0; i < n_stages; i++) {
            double* kq[7];
            comp_ptrs<double>(&c->bufs[c->rk_k[i]], kq);
            for (int q = 0; q < 5; q++) T.k[i][q] = kq[q];
            T.scale[i] = b[i];
        }
        launch_rk_update<double>(c, uq, T, d_dt, uq, fq, c->n_total);
    } else {
        for (int i = 0; i < n_stages; i++) {
            rc = apply_dt(c, U, U, &c->bufs[c->rk_k[i]], 0.0, d_dt, b[i]);                // :185-186
            if (rc) return rc;
        }
        launch_cons_to_prim<double>(c, uq, fq, c->n_total);                              // :193
    }
    // dt == NULL and dt_stable == NULL: nothing returns to the host, the step stays asynchronous on the stream
    // (ibis_b200_rk_last_dt fetches the two values later)
    if (dt || dt_stable) {
        double out[2];
        rc = fetch_scalars(c, SC_DT_STABLE, 2, out);
        if (dt_stable) *dt_stable = out[0];
        if (dt) *dt = out[1];
        if (rc) return rc;
    }
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

int ibis_b200_rk_last_dt(ibis_b200_ctx* c, double* dt, double* dt_stable) {
    if (!c) return fail(c, IBIS_B200_EINVAL, "null argument");
    double out[2];
    int rc = fetch_scalars(c, SC_DT_STABLE, 2, out);
    if (dt_stable) *dt_stable = out[0];
    if (dt) *dt = out[1];
    return rc;
}

int ibis_b200_rk_distinct_stages(ibis_b200_ctx* c, int distinct) {
    c->rk_alias_stages = distinct == 0;
    return 0;
}

int ibis_b200_rk_residual_norms(ibis_b200_ctx* c, double out[6]) {
    if (c->rk_k.empty()) return fail(c, IBIS_B200_EINVAL, "rk_residual_norms before rk_step");
    return ibis_b200_l2_norms(c, c->rk_k[0], out);
}

// ---- JFNK ----
int ibis_b200_jfnk_set_dual(ibis_b200_ctx* c, int cq, int vec, int cq_tmp) {
    Buffer *q = get_buf(c, cq, BUF_CQ), *v = get_buf(c, vec, BUF_VEC), *t = get_buf(c, cq_tmp, BUF_CQ);
    if (!q || !v || !t || !q->is_dual || !t->is_dual || v->n != c->n_cells * c->n_cons)
        return fail(c, IBIS_B200_EHANDLE, "jfnk_set_dual: bad handles");
    dual *a[7], *b[7];
    comp_ptrs<dual>(q, a);
    comp_ptrs<dual>(t, b);
    launch_jfnk_set_dual(c, a, static_cast<double*>(v->d), b);
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_jfnk_set_vector(ibis_b200_ctx* c, int residual, int vec, double dt_star, int result) {
    Buffer *r = get_buf(c, residual, BUF_CQ), *v = get_buf(c, vec, BUF_VEC), *o = get_buf(c, result, BUF_VEC);
    if (!r || !v || !o || !r->is_dual || v->n != c->n_cells * c->n_cons || o->n != v->n)
        return fail(c, IBIS_B200_EHANDLE, "jfnk_set_vector: bad handles");
    dual* a[7];
    comp_ptrs<dual>(r, a);
    launch_jfnk_set_vector(c, a, static_cast<double*>(v->d), dt_star, static_cast<double*>(o->d));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_jfnk_eval_rhs(ibis_b200_ctx* c, int residual, int rhs) {
    Buffer *r = get_buf(c, residual, BUF_CQ), *o = get_buf(c, rhs, BUF_VEC);
    if (!r || !o || !r->is_dual || o->n != c->n_cells * c->n_cons)
        return fail(c, IBIS_B200_EHANDLE, "jfnk_eval_rhs: bad handles");
    dual* a[7];
    comp_ptrs<dual>(r, a);
    launch_jfnk_eval_rhs(c, a, static_cast<double*>(o->d));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_jfnk_apply_update(ibis_b200_ctx* c, int cq, int du_) {
    Buffer *q = get_buf(c, cq, BUF_CQ), *v = get_buf(c, du_, BUF_VEC);
    if (!q || !v || !q->is_dual || v->n != c->n_cells * c->n_cons)
        return fail(c, IBIS_B200_EHANDLE, "jfnk_apply_update: bad handles");
    dual* a[7];
    comp_ptrs<dual>(q, a);
    launch_jfnk_apply_update(c, a, static_cast<double*>(v->d));
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

// ---- vectors ----
// Inner products on a decomposed grid.  A Krylov vector has n_cells * n_cons entries in reference numbering, owned
// cells first (partition.py, box_grid.cpp); the entries of halo cells are copies nobody reads (the dual flow state of
// a halo cell arrives through the halo exchange of compute_dudt).  Dots and norms therefore run over the owned prefix
// and are summed over the ranks before anything is derived from them (square root, reciprocal, the axpy of
// Gram-Schmidt): every rank sees the same Hessenberg entries and residual norms and takes the same decisions.
static int dot_collective(ibis_b200_ctx* c, const double* a, const double* b, int64_t n, double* d_out, int post) {
    if (c->n_ranks <= 1) {
        launch_dot(c, a, b, n, d_out, post);
        return 0;
    }
    const int64_t n_red = (n == c->n_cells * c->n_cons) ? c->n_owned * c->n_cons : n;
    launch_dot(c, a, b, n_red, d_out, 0);
    int rc = comm::allreduce(c, d_out, 1, comm::OP_SUM);
    if (rc) return rc;
    if (post) launch_scalar_post(c, d_out, post);
    return 0;
}
int ibis_b200_vec_dot(ibis_b200_ctx* c, int a, int b, double* out) {
    Buffer *x = get_buf(c, a, BUF_VEC), *y = get_buf(c, b, BUF_VEC);
    if (!x || !y || x->n != y->n || !out) return fail(c, IBIS_B200_EHANDLE, "vec_dot: bad handles");
    int rc = dot_collective(c, static_cast<double*>(x->d), static_cast<double*>(y->d), x->n, c->d_scalars + SC_DOT, 0);
    if (rc) return rc;
    return fetch_scalars(c, SC_DOT, 1, out);
}
int ibis_b200_vec_norm2(ibis_b200_ctx* c, int a, double* out) {
    Buffer* x = get_buf(c, a, BUF_VEC);
    if (!x || !out) return fail(c, IBIS_B200_EHANDLE, "vec_norm2: bad handle");
    int rc = dot_collective(c, static_cast<double*>(x->d), static_cast<double*>(x->d), x->n, c->d_scalars + SC_NORM, 1);
    if (rc) return rc;
    return fetch_scalars(c, SC_NORM, 1, out);
}
int ibis_b200_vec_axpy(ibis_b200_ctx* c, int y, int x, double s) {
    Buffer *a = get_buf(c, y, BUF_VEC), *b = get_buf(c, x, BUF_VEC);
    if (!a || !b || a->n != b->n) return fail(c, IBIS_B200_EHANDLE, "vec_axpy: bad handles");
    launch_axpy(c, static_cast<double*>(a->d), static_cast<double*>(b->d), s, nullptr, 1.0, a->n);
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_vec_scale(ibis_b200_ctx* c, int x, int result, double s) {
    Buffer *a = get_buf(c, x, BUF_VEC), *r = get_buf(c, result, BUF_VEC);
    if (!a || !r || a->n != r->n) return fail(c, IBIS_B200_EHANDLE, "vec_scale: bad handles");
    launch_scale(c, static_cast<double*>(a->d), static_cast<double*>(r->d), s, nullptr, a->n);
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_vec_sub(ibis_b200_ctx* c, int a, int b, int result) {
    Buffer *x = get_buf(c, a, BUF_VEC), *y = get_buf(c, b, BUF_VEC), *r = get_buf(c, result, BUF_VEC);
    if (!x || !y || !r || x->n != y->n || x->n != r->n) return fail(c, IBIS_B200_EHANDLE, "vec_sub: bad handles");
    launch_sub(c, static_cast<double*>(x->d), static_cast<double*>(y->d), static_cast<double*>(r->d), x->n);
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}
int ibis_b200_basis_set_col(ibis_b200_ctx* c, int basis, int col, int vec) {
    Buffer *B = get_buf(c, basis, BUF_BASIS), *v = get_buf(c, vec, BUF_VEC);
    if (!B || !v || col < 0 || col >= B->ncomp || v->n != B->n) return fail(c, IBIS_B200_EHANDLE, "basis_set_col: bad handles");
    GPU_TRY(c, gpu::d2d(static_cast<double*>(B->d) + (int64_t)col * B->stride, v->d, v->n * sizeof(double), c->stream));
    return 0;
}

// modified Gram-Schmidt of gmres.cpp:144-150, all on the stream, one host synchronisation
int ibis_b200_arnoldi_mgs(ibis_b200_ctx* c, int basis, int j, int w, int v, double* h) {
    Buffer *B = get_buf(c, basis, BUF_BASIS), *W = get_buf(c, w, BUF_VEC), *V = get_buf(c, v, BUF_VEC);
    if (!B || !W || !V || j < 0 || j + 1 >= B->ncomp || W->n != B->n || V->n != B->n || !h)
        return fail(c, IBIS_B200_EHANDLE, "arnoldi_mgs: bad arguments");
    double* wd = static_cast<double*>(W->d);
    double* vd = static_cast<double*>(V->d);
    double* hs = c->d_scalars + SC_H;
    for (int i = 0; i <= j; i++) {
        const double* vi = static_cast<double*>(B->d) + (int64_t)i * B->stride;
        int rc = dot_collective(c, wd, vi, W->n, hs + i, 0);         // H(i,j) = dot(w, V_i), summed over the ranks
        if (rc) return rc;
        launch_axpy(c, wd, vi, 0.0, hs + i, -1.0, W->n);             // w += V_i * (-H(i,j))
    }
    int rc = dot_collective(c, wd, wd, W->n, hs + j + 1, 2);         // H(j+1,j) = norm2(w); hs[j+2] = 1/H
    if (rc) return rc;
    launch_scale(c, wd, vd, 0.0, hs + j + 2, W->n);                  // v = w * (1.0 / H(j+1,j))
    GPU_TRY(c, gpu::d2d(static_cast<double*>(B->d) + (int64_t)(j + 1) * B->stride, vd, V->n * sizeof(double), c->stream));
    return fetch_scalars(c, SC_H, j + 2, h);
}

int ibis_b200_basis_gemv(ibis_b200_ctx* c, int basis, int m, const double* y, int out) {
    Buffer *B = get_buf(c, basis, BUF_BASIS), *O = get_buf(c, out, BUF_VEC);
    if (!B || !O || m < 0 || m > B->ncomp || O->n != B->n || (m > 0 && !y)) return fail(c, IBIS_B200_EHANDLE, "basis_gemv: bad arguments");
    for (int i = 0; i < m; i++) c->h_scalars[SC_Y + i] = y[i];
    if (m > 0) GPU_TRY(c, gpu::h2d(c->d_scalars + SC_Y, c->h_scalars + SC_Y, m * sizeof(double), c->stream));
    launch_gemv(c, static_cast<double*>(B->d), B->n, B->stride, m, c->d_scalars + SC_Y, static_cast<double*>(O->d));
    GPU_TRY(c, gpu::stream_sync(c->stream));                         // h_scalars is reused by the next call
    return c->last_error.empty() ? 0 : IBIS_B200_ECUDA;
}

// ---- domain decomposition ----
int ibis_b200_halo_setup(ibis_b200_ctx* c, int n_peers, const int32_t* peer_ranks, const int64_t* send_offsets,
                         const int64_t* send_cells, const int64_t* recv_offsets, const int64_t* recv_cells) {
    return comm::halo_setup(c, n_peers, peer_ranks, send_offsets, send_cells, recv_offsets, recv_cells);
}
int ibis_b200_comm_init(ibis_b200_ctx* c, const void* id, int rank, int n_ranks) {
    return comm::init(c, id, rank, n_ranks);
}
int ibis_b200_nccl_unique_id(void* out) { return comm::unique_id(out); }
int ibis_b200_halo_p2p_export(ibis_b200_ctx* c, void* handle64, int64_t* n_recv) { return comm::p2p_export(c, handle64, n_recv); }
int ibis_b200_halo_p2p_import(ibis_b200_ctx* c, int peer_index, const void* handle64, int64_t peer_n_recv,
                              int64_t my_offset_in_peer, int my_slot_in_peer) {
    return comm::p2p_import(c, peer_index, handle64, peer_n_recv, my_offset_in_peer, my_slot_in_peer);
}
int ibis_b200_halo_p2p_enable(ibis_b200_ctx* c, int on) { return comm::p2p_enable(c, on); }

}  // extern "C"
