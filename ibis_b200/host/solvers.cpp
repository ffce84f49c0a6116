// solvers.cpp -- see solvers.hpp.  Line references are to /root/reference/src.
#include "solvers.hpp"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <limits>

namespace ibis_host {

void check(ibis_b200_ctx* ctx, int rc, const char* what) {
    if (rc != 0) {
        const char* msg = ibis_b200_last_error(ctx);
        throw AbiError(rc, std::string(what) + ": " + (msg ? msg : "?"));
    }
}
#define ABI(call) check(ctx_, (call), #call)

// ========================================================================================
// CFL schedules (solvers/cfl.cpp:11-84)
// ========================================================================================
double LinearSchedule::eval(double t) {
    if (t <= times_[0]) return cfls_[0];
    size_t n = times_.size();
    if (t >= times_[n - 1]) return cfls_[n - 1];
    for (size_t i = 0; i + 1 < times_.size(); i++) {
        if (t >= times_[i] && t < times_[i + 1]) {
            double frac = (t - times_[i]) / (times_[i + 1] - times_[i]);
            return cfls_[i] + frac * (cfls_[i + 1] - cfls_[i]);
        }
    }
    throw std::runtime_error("Shouldn't reach here");
}

double ResidualBasedCfl::eval(double t) {          // cfl.cpp:57-70; stateful
    if (t > threshold_) {
        previous_residual_ = t;
        return previous_cfl_;
    }
    double ratio = previous_residual_ / t;
    double new_cfl = previous_cfl_ * pow(ratio, power_);
    new_cfl = std::min(new_cfl, max_cfl_);
    new_cfl = std::max(start_cfl_, new_cfl);
    previous_residual_ = t;
    previous_cfl_ = new_cfl;
    return new_cfl;
}

std::unique_ptr<CflSchedule> make_cfl_schedule(const ibis_b200_cfl_config& c) {
    switch (c.type) {
        case IBIS_B200_CFL_CONSTANT:
            return std::unique_ptr<CflSchedule>(new ConstantSchedule(c.value));
        case IBIS_B200_CFL_LINEAR_INTERPOLATE:
            if (c.n < 1 || !c.times || !c.cfls) throw std::runtime_error("Unknown CFL schedule");
            return std::unique_ptr<CflSchedule>(new LinearSchedule(
                std::vector<double>(c.times, c.times + c.n), std::vector<double>(c.cfls, c.cfls + c.n)));
        case IBIS_B200_CFL_RESIDUAL_BASED:
            return std::unique_ptr<CflSchedule>(
                new ResidualBasedCfl(c.growth_threshold, c.power, c.start_cfl, c.max_cfl));
        default:
            throw std::runtime_error("Unknown CFL schedule");
    }
}

// ========================================================================================
// Hessenberg least squares on the host (linear_algebra/gmres.cpp:22-74)
// ========================================================================================
// The reference stores H as (m+1) x m and takes (j+2) x (j+2) sub-matrices of it, which at
// j = m-1 runs one past the end; here the storage is (m+2) x (m+2).  The
// arithmetic is the reference's: its dense gemm/gemv with the mostly-identity rotation
// matrix Omega reduce to the two-row updates below (every other term is an exact 0*x or
// 1*x), evaluated in the same order (ascending inner index, starting from 0.0).
HessenbergQR::HessenbergQR(size_t max_iters) : m_(max_iters + 2) {
    H_.assign(m_ * m_, 0.0);
    Q_.assign(m_ * m_, 0.0);
    g_.assign(m_, 0.0);
    hr_.assign(m_, 0.0);
    reset();
}

void HessenbergQR::reset() {
    std::fill(H_.begin(), H_.end(), 0.0);
    std::fill(Q_.begin(), Q_.end(), 0.0);
    for (size_t i = 0; i < m_; i++) Q(i, i) = 1.0;
    std::fill(g_.begin(), g_.end(), 0.0);
}

double HessenbergQR::add_column(size_t j, const double* h) {
    for (size_t i = 0; i < j + 2; i++) H(i, j) = h[i];
    if (j != 0) {
        // apply previous rotations to the new column: h_rotated = Q(0:j+1,0:j+1) h(0:j+1)
        for (size_t r = 0; r < j + 1; r++) {
            double dot = 0.0;
            for (size_t c = 0; c < j + 1; c++) dot += Q(r, c) * H(c, j);
            hr_[r] = dot;
        }
        for (size_t r = 0; r < j + 1; r++) H(r, j) = hr_[r];
    }
    const double denom = sqrt(H(j, j) * H(j, j) + H(j + 1, j) * H(j + 1, j));
    const double si = H(j + 1, j) / denom;
    const double ci = H(j, j) / denom;
    // H_new = Omega H_old on the leading (j+2) x (j+2) block: rows j and j+1 change
    for (size_t c = 0; c < j + 2; c++) {
        const double a = H(j, c), b = H(j + 1, c);
        H(j, c) = (0.0 + ci * a) + si * b;
        H(j + 1, c) = (0.0 + (-si) * a) + ci * b;
    }
    {
        const double a = g_[j], b = g_[j + 1];
        g_[j] = (0.0 + ci * a) + si * b;
        g_[j + 1] = (0.0 + (-si) * a) + ci * b;
    }
    // Q_new = Omega Q_old (j == 0: Q_new = Omega)
    for (size_t c = 0; c < j + 2; c++) {
        const double a = Q(j, c), b = Q(j + 1, c);
        Q(j, c) = (0.0 + ci * a) + si * b;
        Q(j + 1, c) = (0.0 + (-si) * a) + ci * b;
    }
    return fabs(g_[j + 1]);
}

void HessenbergQR::solve(size_t n, double* y) const {
    if (n == 0) return;
    const int nn = (int)n;
    y[nn - 1] = g_[nn - 1] / H(nn - 1, nn - 1);
    for (int i = nn - 2; i >= 0; i--) {
        double sum = g_[i];
        for (int j = i + 1; j < nn; j++) sum -= H(i, j) * y[j];
        y[i] = sum / H(i, i);
    }
}

// ========================================================================================
// Gmres (linear_algebra/gmres.cpp:80-182)
// ========================================================================================
static void compute_r0_(ibis_b200_ctx* ctx_, LinearSystem& system, int x0, int r0, int w) {
    system.matrix_vector_product(x0, w);                                 // gmres.cpp:13
    ABI(ibis_b200_vec_sub(ctx_, system.rhs(), w, r0));                   // "Gmres::b-Ax0" :18
}

Gmres::Gmres(ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, size_t max_iters, double tol)
    : ctx_(ctx), system_(system), num_vars_(system->num_vars()),
      max_iters_(std::min(system->num_vars(), max_iters)), tol_(tol), qr_(max_iters_) {
    ABI(ibis_b200_basis_create(ctx_, (int64_t)num_vars_, (int)max_iters_ + 1, &krylov_vectors_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &r0_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &w_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &v_));
    h_.assign(max_iters_ + 2, 0.0);
    ym_.assign(max_iters_ + 1, 0.0);
}

Gmres::~Gmres() {
    for (int h : {krylov_vectors_, r0_, w_, v_})
        if (h) ibis_b200_free(ctx_, h);
}

LinearSolveResult Gmres::solve(int x0) {
    qr_.reset();
    compute_r0_(ctx_, *system_, x0, r0_, w_);
    double beta = 0.0;
    ABI(ibis_b200_vec_norm2(ctx_, r0_, &beta));
    LinearSolveResult result;
    result.tol = tol_;
    result.residual = beta;
    if (beta < tol_) {                                                   // gmres.cpp:132
        result.success = true;
        return result;
    }
    qr_.set_g0(beta);
    ABI(ibis_b200_vec_scale(ctx_, r0_, v_, 1.0 / beta));
    ABI(ibis_b200_basis_set_col(ctx_, krylov_vectors_, 0, v_));

    for (size_t j = 0; j < max_iters_; j++) {
        system_->matrix_vector_product(v_, w_);
        // modified Gram-Schmidt + normalisation + V_{j+1} = v  (gmres.cpp:144-150)
        ABI(ibis_b200_arnoldi_mgs(ctx_, krylov_vectors_, (int)j, w_, v_, h_.data()));
        const double residual = qr_.add_column(j, h_.data());
        result.residual = residual / beta;
        result.n_iters = j + 1;
        if (residual < tol_ * beta) {                                     // gmres.cpp:163
            result.success = true;
            break;
        }
    }
    const size_t n_vectors = result.n_iters;
    qr_.solve(n_vectors, ym_.data());
    ABI(ibis_b200_basis_gemv(ctx_, krylov_vectors_, (int)n_vectors, ym_.data(), w_));
    ABI(ibis_b200_vec_axpy(ctx_, x0, w_, 1.0));
    return result;
}

// ========================================================================================
// FGmres (linear_algebra/gmres.cpp:184-303)
// ========================================================================================
FGmres::FGmres(ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, size_t max_iters, double tol,
               std::shared_ptr<LinearSystem> precondition_system, size_t max_precondition_iters,
               double precondition_tol)
    : ctx_(ctx), system_(system), precondition_system_(precondition_system),
      num_vars_(system->num_vars()), max_iters_(std::min(system->num_vars(), max_iters)), tol_(tol),
      qr_(max_iters_), preconditioner_(ctx, precondition_system, max_precondition_iters, precondition_tol) {
    ABI(ibis_b200_basis_create(ctx_, (int64_t)num_vars_, (int)max_iters_ + 1, &krylov_vectors_));
    ABI(ibis_b200_basis_create(ctx_, (int64_t)num_vars_, (int)max_iters_ + 1, &preconditioned_krylov_vectors_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &r0_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &w_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &v_));
    ABI(ibis_b200_vec_create(ctx_, (int64_t)num_vars_, &z_));
    h_.assign(max_iters_ + 2, 0.0);
    ym_.assign(max_iters_ + 1, 0.0);
}

FGmres::~FGmres() {
    for (int h : {krylov_vectors_, preconditioned_krylov_vectors_, r0_, w_, v_, z_})
        if (h) ibis_b200_free(ctx_, h);
}

LinearSolveResult FGmres::solve(int x) {
    qr_.reset();
    compute_r0_(ctx_, *system_, x, r0_, w_);
    double beta = 0.0;
    ABI(ibis_b200_vec_norm2(ctx_, r0_, &beta));
    LinearSolveResult result;
    result.tol = tol_;
    result.residual = beta;
    if (beta < tol_) {
        result.success = true;
        return result;
    }
    qr_.set_g0(beta);
    ABI(ibis_b200_vec_scale(ctx_, r0_, v_, 1.0 / beta));
    ABI(ibis_b200_basis_set_col(ctx_, krylov_vectors_, 0, v_));

    for (size_t j = 0; j < max_iters_; j++) {
        precondition_system_->set_rhs(v_);                                // gmres.cpp:259
        ABI(ibis_b200_zero(ctx_, z_));
        preconditioner_.solve(z_);
        ABI(ibis_b200_basis_set_col(ctx_, preconditioned_krylov_vectors_, (int)j, z_));

        system_->matrix_vector_product(z_, w_);
        ABI(ibis_b200_arnoldi_mgs(ctx_, krylov_vectors_, (int)j, w_, v_, h_.data()));
        const double residual = qr_.add_column(j, h_.data());
        result.residual = residual / beta;
        result.n_iters = j + 1;
        if (residual < tol_ * beta) {
            result.success = true;
            break;
        }
    }
    const size_t n_vectors = result.n_iters;
    qr_.solve(n_vectors, ym_.data());
    ABI(ibis_b200_basis_gemv(ctx_, preconditioned_krylov_vectors_, (int)n_vectors, ym_.data(), w_));
    ABI(ibis_b200_vec_axpy(ctx_, x, w_, 1.0));
    return result;
}

std::unique_ptr<IterativeLinearSolver> make_linear_solver(
    ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, std::shared_ptr<LinearSystem> preconditioner,
    const ibis_b200_linear_solver_config& config) {
    if (config.type == IBIS_B200_LINEAR_GMRES)
        return std::unique_ptr<IterativeLinearSolver>(new Gmres(ctx, system, (size_t)config.max_iters, config.tol));
    if (config.type == IBIS_B200_LINEAR_FGMRES)
        return std::unique_ptr<IterativeLinearSolver>(
            new FGmres(ctx, system, (size_t)config.max_iters, config.tol, preconditioner,
                       (size_t)config.max_preconditioner_iters, config.preconditioner_tolerance));
    throw std::runtime_error("Unknown linear solver");
}

// ========================================================================================
// SteadyStateLinearisation (solvers/steady_state.cpp:18-170)
// ========================================================================================
SteadyStateLinearisation::SteadyStateLinearisation(ibis_b200_ctx* ctx, SimSizes sizes, int residuals, int cq,
                                                   int fs, bool allow_reconstruction)
    : ctx_(ctx), sizes_(sizes), n_cells_(sizes.n_cells), n_cons_(sizes.dim + 2), residuals_(residuals),
      cq_(cq), fs_(fs), allow_reconstruction_(allow_reconstruction) {
    ABI(ibis_b200_vec_create(ctx_, n_cells_ * n_cons_, &own_rhs_));
    rhs_ = own_rhs_;
    ABI(ibis_b200_cq_create(ctx_, 1, &cq_tmp_));
    ABI(ibis_b200_fs_create(ctx_, 1, &fs_tmp_));
}

SteadyStateLinearisation::~SteadyStateLinearisation() {
    for (int h : {own_rhs_, cq_tmp_, fs_tmp_})
        if (h) ibis_b200_free(ctx_, h);
}

std::shared_ptr<PseudoTransientLinearSystem> SteadyStateLinearisation::preconditioner() {
    // the same system with allow_reconstruction = false (steady_state.cpp:47-50)
    return std::shared_ptr<PseudoTransientLinearSystem>(
        new SteadyStateLinearisation(ctx_, sizes_, residuals_, cq_, fs_, false));
}

void SteadyStateLinearisation::matrix_vector_product(int vec, int result) {
    ABI(ibis_b200_jfnk_set_dual(ctx_, cq_, vec, cq_tmp_));                                  // :60-68
    ABI(ibis_b200_cons_to_prim(ctx_, cq_tmp_, fs_tmp_));                                    // :91
    ABI(ibis_b200_compute_dudt(ctx_, fs_tmp_, residuals_, allow_reconstruction_ ? 1 : 0));  // :94-98
    ABI(ibis_b200_jfnk_set_vector(ctx_, residuals_, vec, dt_star_, result));                // :102-110
}

void SteadyStateLinearisation::eval_rhs() {
    ABI(ibis_b200_compute_dudt(ctx_, fs_, residuals_, allow_reconstruction_ ? 1 : 0));      // :132-136
    ABI(ibis_b200_jfnk_eval_rhs(ctx_, residuals_, rhs_));                                   // :142-149
}

// ========================================================================================
// Jfnk (solvers/jfnk.cpp)
// ========================================================================================
Jfnk::Jfnk(ibis_b200_ctx* ctx, SimSizes sizes, std::shared_ptr<PseudoTransientLinearSystem> system,
           std::unique_ptr<CflSchedule>&& cfl, int residuals, const ibis_b200_steady_config& config)
    : ctx_(ctx) {
    (void)sizes;
    max_steps_ = (size_t)config.max_steps;
    tolerance_ = config.tolerance;
    system_ = system;
    preconditioner_ = system_->preconditioner();
    gmres_ = make_linear_solver(ctx, system_, preconditioner_, config.linear_solver);
    cfl_ = std::move(cfl);
    residual_based_cfl_ = cfl_->residual_based();
    ABI(ibis_b200_vec_create(ctx_, (int64_t)system_->num_vars(), &dU_));
    residuals_ = residuals;
}

Jfnk::~Jfnk() {
    gmres_.reset();
    if (dU_) ibis_b200_free(ctx_, dU_);
}

int Jfnk::initialise() {
    system_->eval_rhs();
    ABI(ibis_b200_l2_norms(ctx_, residuals_, residual_norms_.v));
    initial_residual_norms_ = residual_norms_;
    return 0;
}

double Jfnk::calculate_cfl(size_t step) const {                       // jfnk.h:41-49
    if (residual_based_cfl_) return cfl_->eval(relative_residual_norms().global());
    return cfl_->eval((double)step);
}

Norms6 Jfnk::relative_residual_norms() const {                        // conserved_quantities.h operator/
    Norms6 r;
    for (int i = 0; i < 6; i++) r.v[i] = residual_norms_.v[i] / initial_residual_norms_.v[i];
    return r;
}

void Jfnk::set_pseudo_time_step_size(double dt_star) {
    system_->set_pseudo_time_step(dt_star);
    if (preconditioner_) preconditioner_->set_pseudo_time_step(dt_star);
}

LinearSolveResult Jfnk::step(int cq, int fs, size_t step) {
    ABI(ibis_b200_zero(ctx_, dU_));                                     // jfnk.cpp:44
    ABI(ibis_b200_estimate_dt(ctx_, fs, &stable_dt_));                  // :47
    const double cfl = calculate_cfl(step);
    set_pseudo_time_step_size(cfl * stable_dt_);
    last_gmres_result_ = gmres_->solve(dU_);                            // :52
    apply_update_(cq, fs);                                              // :57
    system_->eval_rhs();
    ABI(ibis_b200_l2_norms(ctx_, residuals_, residual_norms_.v));
    return last_gmres_result_;
}

void Jfnk::apply_update_(int cq, int fs) {
    ABI(ibis_b200_jfnk_apply_update(ctx_, cq, dU_));                    // jfnk.cpp:69-76
    ABI(ibis_b200_cons_to_prim(ctx_, cq, fs));                          // :92
}

// ========================================================================================
// Solver::solve (solvers/solver.cpp:18-64)
// ========================================================================================
int Solver::solve() {
    memset(&report, 0, sizeof report);
    int success = initialise();
    if (success != 0) {
        report.status = success;
        return success;
    }
    const auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    int rc = 0;
    for (size_t step = 0; step < max_step(); step++) {
        int result = take_step(step);
        report.steps = (int64_t)step + 1;
        if (residuals_this_step((unsigned)step)) write_residuals((unsigned)step, elapsed());
        if (result != 0) {
            plot_solution((unsigned)step);
            rc = 1;
            break;
        }
        int bad_cells = count_bad_cells();
        if (bad_cells > 0) {
            plot_solution((unsigned)step);
            rc = 1;
            break;
        }
        if (stop_now((unsigned)step)) {
            report.stop_reason = stop_reason((unsigned)step);
            plot_solution((unsigned)step);
            break;
        }
        if (print_this_step((unsigned)step)) print_progress((unsigned)step, elapsed());
        if (plot_this_step((unsigned)step)) plot_solution((unsigned)step);
    }
    const int64_t steps = report.steps;
    const int reason = report.stop_reason;
    fill_report();
    report.steps = steps;
    report.stop_reason = reason;
    report.status = rc;
    report.wall_seconds = elapsed();
    finalise();
    return rc;
}

// ========================================================================================
// RungeKutta (solvers/runge_kutta.cpp)
// ========================================================================================
ButcherTableau::ButcherTableau(int n_stages, const double* a, const double* b) {
    b_.assign(b, b + n_stages);
    int row = 0;
    for (int i = 1; i < n_stages; i++) {
        a_.push_back(std::vector<double>(a + row, a + row + i));
        row += i;
    }
    a_flat_.assign(a, a + row);
    if (a_flat_.empty()) a_flat_.push_back(0.0);
}

RungeKutta::RungeKutta(ibis_b200_ctx* ctx, const ibis_b200_rk_config& c, const double* prim0, double* prim_out,
                       const ibis_b200_observers* obs)
    : ctx_(ctx), prim0_(prim0), prim_out_(prim_out) {
    memset(&obs_, 0, sizeof obs_);
    if (obs) obs_ = *obs;
    if (c.n_stages < 1 || !c.b) throw std::runtime_error("RungeKutta: bad butcher tableau");
    tableau_ = ButcherTableau(c.n_stages, c.a, c.b);
    cfl_ = make_cfl_schedule(c.cfl);
    moving_grid_ = c.moving_grid != 0;                                                 // :58
    fused_ = c.fused_step != 0 && !moving_grid_;
    max_time_ = c.max_time;
    max_step_ = (size_t)c.max_step;
    dt_init_ = c.dt_init;
    print_frequency_ = (unsigned)std::max(c.print_frequency, 1);
    plot_every_n_steps_ = c.plot_every_n_steps;
    plot_frequency_ = c.plot_frequency;
    residuals_every_n_steps_ = c.residuals_every_n_steps;
    residual_frequency_ = c.residual_frequency;
    // memory (runge_kutta.cpp:45-52): flow_, conserved_quantities_, k_[stages], k_tmp_, flow_tmp_
    ABI(ibis_b200_fs_create(ctx_, 0, &flow_));
    ABI(ibis_b200_cq_create(ctx_, 0, &conserved_quantities_));
    ABI(ibis_b200_rk_distinct_stages(ctx_, c.distinct_stages));
    if (!fused_) {
        // runge_kutta.cpp:49-50: std::vector<ConservedQuantities>(num_stages, ConservedQuantities(n, dim))
        // copies ONE view handle into every slot -- the stages alias (kept unless distinct_stages)
        for (size_t i = 0; i < tableau_.num_stages(); i++) {
            int h = 0;
            if (i == 0 || c.distinct_stages) {
                ABI(ibis_b200_cq_create(ctx_, 0, &h));
                k_own_.push_back(h);
            } else {
                h = k_[0];
            }
            k_.push_back(h);
        }
        ABI(ibis_b200_cq_create(ctx_, 0, &k_tmp_));
        ABI(ibis_b200_fs_create(ctx_, 0, &flow_tmp_));
    }
}

RungeKutta::~RungeKutta() {
    for (int h : k_own_) ibis_b200_free(ctx_, h);
    for (int h : {flow_, conserved_quantities_, k_tmp_, flow_tmp_})
        if (h) ibis_b200_free(ctx_, h);
}

int RungeKutta::initialise() {
    ABI(ibis_b200_fs_upload(ctx_, flow_, prim0_, nullptr));                            // io_.read :87
    ABI(ibis_b200_prim_to_cons(ctx_, conserved_quantities_, flow_));                   // :89
    dt_ = (dt_init_ > 0) ? dt_init_ : std::numeric_limits<double>::max();              // :90
    if (!fused_) function_eval_(flow_, conserved_quantities_, 0);                      // :93
    if (residuals_every_n_steps_ > 0 || residual_frequency_ > 0) {
        if (fused_) {
            // the fused step owns its k_[0]; evaluate the initial residual into a scratch buffer
            int k0 = 0;
            double norms[6];
            ABI(ibis_b200_cq_create(ctx_, 0, &k0));
            ABI(ibis_b200_compute_dudt(ctx_, flow_, k0, 1));
            ABI(ibis_b200_l2_norms(ctx_, k0, norms));
            ABI(ibis_b200_free(ctx_, k0));
            if (obs_.residuals) obs_.residuals(obs_.user, 0, t_, 0.0, norms);
            time_since_last_residual_ = 0;
        } else {
            write_residuals(0, 0.0);
        }
    }
    return 0;
}

int RungeKutta::finalise() { return 0; }

void RungeKutta::estimate_dt() {                                                       // :108-119
    ABI(ibis_b200_estimate_dt(ctx_, flow_, &stable_dt_));
    double dt_startup = std::min(cfl_->eval(t_) * stable_dt_, 1.5 * dt_);
    dt_ = std::min(dt_startup, max_time_ - t_);
    if (plot_frequency_ > 0.0 && time_since_last_plot_ < plot_frequency_)
        dt_ = std::min(dt_, plot_frequency_ - time_since_last_plot_);
}

void RungeKutta::function_eval_(int fs, int cq, size_t index) {                        // :121-129
    if (moving_grid_) ABI(ibis_b200_compute_dudt_moving(ctx_, fs, cq, k_[index], 1));
    else ABI(ibis_b200_compute_dudt(ctx_, fs, k_[index], 1));
}

int RungeKutta::take_step(size_t step) {
    (void)step;
    if (fused_) {
        // the whole of :131-203 in one device-resident call; the extra dt bound of the plot
        // frequency enters through t_remaining
        double t_remaining = max_time_ - t_;
        if (plot_frequency_ > 0.0 && time_since_last_plot_ < plot_frequency_)
            t_remaining = std::min(t_remaining, plot_frequency_ - time_since_last_plot_);
        double dt = 0.0, dts = 0.0;
        ABI(ibis_b200_rk_step(ctx_, flow_, conserved_quantities_, (int)tableau_.num_stages(),
                              tableau_.a_flat().data(), tableau_.b_vec().data(), cfl_->eval(t_), dt_, t_remaining,
                              &dt, &dts));
        dt_ = dt;
        stable_dt_ = dts;
    } else {
        // Moving grid: `init_vertex_pos_ = grid_.vertices().positions()` (:136) copies a view HANDLE, so
        // init_vertex_pos_ IS the grid's position array, and vertex_vel_ (:64-65) is one array in every stage
        // slot: each add_scaled_vector below (:158,172,189) therefore displaces the grid's own vertices in
        // place, cumulatively.  Reproduced as such: one move per reference call.
        function_eval_(flow_, conserved_quantities_, 0);                               // :143
        estimate_dt();                                                                 // :147
        for (size_t i = 1; i < tableau_.num_stages(); i++) {
            ABI(ibis_b200_apply_time_derivative(ctx_, conserved_quantities_, k_tmp_, k_[0],
                                                tableau_.a(i, 0) * dt_));              // :155
            if (moving_grid_) ABI(ibis_b200_move_vertices(ctx_, tableau_.a(i, 0) * dt_, 0));   // :157-160
            for (size_t j = 1; j < i; j++) {
                if (tableau_.a(i, j) < 1e-14) continue;                                // :166
                ABI(ibis_b200_apply_time_derivative(ctx_, k_tmp_, k_tmp_, k_[j], tableau_.a(i, j) * dt_));
                if (moving_grid_) ABI(ibis_b200_move_vertices(ctx_, tableau_.a(i, j) * dt_, 0));   // :170-173
            }
            ABI(ibis_b200_cons_to_prim(ctx_, k_tmp_, flow_tmp_));                      // :177
            if (moving_grid_) ABI(ibis_b200_move_vertices(ctx_, 0.0, 1));              // compute_geometric_data :178-180
            function_eval_(flow_tmp_, k_tmp_, i);                                      // :181
        }
        for (size_t i = 0; i < tableau_.num_stages(); i++) {                           // :185-191
            ABI(ibis_b200_apply_time_derivative(ctx_, conserved_quantities_, conserved_quantities_, k_[i],
                                                tableau_.b(i) * dt_));
            if (moving_grid_) ABI(ibis_b200_move_vertices(ctx_, tableau_.b(i) * dt_, 0));
        }
        ABI(ibis_b200_cons_to_prim(ctx_, conserved_quantities_, flow_));               // :193
        if (moving_grid_) ABI(ibis_b200_move_vertices(ctx_, 0.0, 1));                  // set_vertex_positions :194-196
    }
    t_ += dt_;
    time_since_last_plot_ += dt_;
    time_since_last_residual_ += dt_;
    return 0;
}

bool RungeKutta::print_this_step(unsigned int step) { return step != 0 && step % print_frequency_ == 0; }

bool RungeKutta::residuals_this_step(unsigned int step) {                              // :210-218
    if ((residuals_every_n_steps_ > 0) && (step != 0 || (step % residuals_every_n_steps_ == 0))) return true;
    if (residual_frequency_ > 0 && (time_since_last_residual_ >= residual_frequency_ - 1e-15)) return true;
    return false;
}

bool RungeKutta::write_residuals(unsigned int step, double wc) {                       // :220-227
    double norms[6];
    if (fused_) ABI(ibis_b200_rk_residual_norms(ctx_, norms));
    else ABI(ibis_b200_l2_norms(ctx_, k_[0], norms));                                  // L2_norms() = k_[0] :264
    if (obs_.residuals) obs_.residuals(obs_.user, step, t_, wc, norms);
    time_since_last_residual_ = 0;
    return true;
}

bool RungeKutta::plot_this_step(unsigned int step) {                                   // :229-236
    if (plot_every_n_steps_ > 0 && step != 0 && step % plot_every_n_steps_ == 0) return true;
    if (plot_frequency_ > 0 && time_since_last_plot_ >= plot_frequency_ - 1e-15) return true;
    return false;
}

int RungeKutta::plot_solution(unsigned int step) {                                     // :238-243
    if (obs_.plot) obs_.plot(obs_.user, step, t_, flow_);
    time_since_last_plot_ = 0.0;
    return 0;
}

bool RungeKutta::stop_now(unsigned int step) {                                         // :258-262
    if (step >= max_step_ - 1) return true;
    if (t_ >= max_time_) return true;
    return false;
}

int RungeKutta::stop_reason(unsigned int step) {                                       // :251-257
    if (t_ >= max_time_ - 1e-15) return 2;
    if (step >= max_step_ - 1) return 1;
    return 0;
}

int RungeKutta::count_bad_cells() {
    uint64_t n = 0;
    ABI(ibis_b200_count_bad_cells(ctx_, flow_, &n));
    last_bad_ = (int64_t)n;
    return (int)std::min<uint64_t>(n, 0x7fffffff);
}

void RungeKutta::fill_report() {
    report.t = t_;
    report.dt = dt_;
    report.stable_dt = stable_dt_;
    report.bad_cells = last_bad_;
    if (fused_) {
        if (report.steps > 0) ABI(ibis_b200_rk_residual_norms(ctx_, report.residual_norms));
    } else {
        ABI(ibis_b200_l2_norms(ctx_, k_[0], report.residual_norms));
    }
    if (prim_out_) ABI(ibis_b200_fs_download(ctx_, flow_, prim_out_, nullptr));
}

// ========================================================================================
// SteadyState (solvers/steady_state.cpp:176-326)
// ========================================================================================
SteadyState::SteadyState(ibis_b200_ctx* ctx, SimSizes sizes, const ibis_b200_steady_config& c, const double* prim0,
                         double* prim_out, const ibis_b200_observers* obs)
    : ctx_(ctx), sizes_(sizes), prim0_(prim0), prim_out_(prim_out) {
    memset(&obs_, 0, sizeof obs_);
    if (obs) obs_ = *obs;
    ABI(ibis_b200_fs_create(ctx_, 1, &fs_));
    ABI(ibis_b200_cq_create(ctx_, 1, &cq_));
    ABI(ibis_b200_cq_create(ctx_, 1, &residuals_));
    auto cfl = make_cfl_schedule(c.cfl);
    std::shared_ptr<PseudoTransientLinearSystem> system(
        new SteadyStateLinearisation(ctx_, sizes_, residuals_, cq_, fs_));
    jfnk_.reset(new Jfnk(ctx_, sizes_, system, std::move(cfl), residuals_, c));
    print_frequency_ = (unsigned)std::max(c.print_frequency, 1);
    plot_frequency_ = (unsigned)std::max(c.plot_frequency, 1);
    diagnostics_frequency_ = (unsigned)std::max(c.diagnostics_frequency, 0);
}

SteadyState::~SteadyState() {
    jfnk_.reset();
    for (int h : {fs_, cq_, residuals_})
        if (h) ibis_b200_free(ctx_, h);
}

int SteadyState::initialise() {
    ABI(ibis_b200_fs_upload(ctx_, fs_, prim0_, nullptr));          // io_.read: dual parts zero
    ABI(ibis_b200_prim_to_cons(ctx_, cq_, fs_));
    int jfnk_init = jfnk_->initialise();
    if (diagnostics_frequency_ > 0) write_residuals(0, 0.0);
    return jfnk_init;
}

int SteadyState::take_step(size_t step) {
    LinearSolveResult r = jfnk_->step(cq_, fs_, step);
    linear_iters_ += (int64_t)r.n_iters;
    return 0;
}

bool SteadyState::print_this_step(unsigned int step) { return step != 0 && step % print_frequency_ == 0; }
bool SteadyState::residuals_this_step(unsigned int step) {
    return (diagnostics_frequency_ > 0) && (step != 0) && (step % diagnostics_frequency_ == 0);
}
bool SteadyState::plot_this_step(unsigned int step) { return step != 0 && step % plot_frequency_ == 0; }

int SteadyState::plot_solution(unsigned int step) {
    if (obs_.plot) obs_.plot(obs_.user, step, (double)step, fs_);
    return 0;
}

bool SteadyState::stop_now(unsigned int step) {
    if (step >= max_step() - 1) return true;
    if (jfnk_->relative_residual_norms().global() < jfnk_->target_residual()) return true;
    return false;
}

int SteadyState::stop_reason(unsigned int step) {
    if (step >= max_step() - 1) return 1;
    if (jfnk_->relative_residual_norms().global() < jfnk_->target_residual()) return 3;
    return 0;
}

bool SteadyState::write_residuals(unsigned int step, double wc) {
    Norms6 n = jfnk_->residual_norms();
    if (obs_.residuals) obs_.residuals(obs_.user, step, (double)step, wc, n.v);
    return true;
}

int SteadyState::count_bad_cells() {
    uint64_t n = 0;
    ABI(ibis_b200_count_bad_cells(ctx_, fs_, &n));
    last_bad_ = (int64_t)n;
    return (int)std::min<uint64_t>(n, 0x7fffffff);
}

void SteadyState::fill_report() {
    report.t = (double)report.steps;
    report.dt = jfnk_->pseudo_time_step_size();
    report.stable_dt = jfnk_->pseudo_time_step_size();
    report.bad_cells = last_bad_;
    Norms6 n = jfnk_->residual_norms();
    for (int i = 0; i < 6; i++) report.residual_norms[i] = n.v[i];
    report.relative_residual = jfnk_->relative_residual_norms().global();
    report.linear_iters = linear_iters_;
    const LinearSolveResult& r = jfnk_->last_gmres_result();
    report.last_linear_success = r.success ? 1 : 0;
    report.last_linear_iters = (int32_t)r.n_iters;
    report.last_linear_residual = r.residual;
    if (prim_out_) ABI(ibis_b200_fs_download(ctx_, fs_, prim_out_, nullptr));
}

}  // namespace ibis_host

// ========================================================================================
// extern "C"
// ========================================================================================
namespace {
thread_local std::string g_host_error;

// dense test system: A held as a device basis of n columns (column j = A(:, j)); the
// matrix-vector product is the basis gemv with the vector's entries as coefficients
// (the reference's TestLinearSystem, linear_algebra/gmres.cpp:321-385)
class DenseSystem : public ibis_host::LinearSystem {
public:
    DenseSystem(ibis_b200_ctx* ctx, const double* a, const double* b, int n) : ctx_(ctx), n_(n) {
        ibis_host::check(ctx, ibis_b200_basis_create(ctx, n, n, &A_), "basis_create");
        ibis_host::check(ctx, ibis_b200_vec_create(ctx, n, &b_), "vec_create");
        ibis_host::check(ctx, ibis_b200_vec_create(ctx, n, &col_), "vec_create");
        ibis_host::check(ctx, ibis_b200_vec_upload(ctx, b_, b), "vec_upload");
        std::vector<double> c(n);
        for (int j = 0; j < n; j++) {
            for (int i = 0; i < n; i++) c[i] = a[i * n + j];
            ibis_host::check(ctx, ibis_b200_vec_upload(ctx, col_, c.data()), "vec_upload");
            ibis_host::check(ctx, ibis_b200_basis_set_col(ctx, A_, j, col_), "basis_set_col");
        }
        rhs_ = b_;
        tmp_.resize(n);
    }
    ~DenseSystem() {
        for (int h : {A_, b_, col_}) ibis_b200_free(ctx_, h);
    }
    void matrix_vector_product(int vec, int result) override {
        ibis_host::check(ctx_, ibis_b200_vec_download(ctx_, vec, tmp_.data()), "vec_download");
        ibis_host::check(ctx_, ibis_b200_basis_gemv(ctx_, A_, n_, tmp_.data(), result), "basis_gemv");
    }
    void eval_rhs() override {}
    void set_rhs(int rhs) override { rhs_ = rhs; }
    int rhs() override { return rhs_; }
    size_t num_vars() const override { return (size_t)n_; }
private:
    ibis_b200_ctx* ctx_;
    int n_, A_ = 0, b_ = 0, col_ = 0, rhs_ = 0;
    std::vector<double> tmp_;
};

ibis_host::SimSizes sizes_of(ibis_b200_ctx* ctx) {
    ibis_host::SimSizes s;
    int64_t n_ghost = 0, n_faces = 0;
    int32_t dim = 0;
    ibis_host::check(ctx, ibis_b200_sizes(ctx, &dim, &s.n_cells, &n_ghost, &n_faces), "ibis_b200_sizes");
    s.dim = dim;
    s.n_total = s.n_cells + n_ghost;
    return s;
}
}  // namespace

extern "C" {

const char* ibis_b200_host_last_error(void) { return g_host_error.c_str(); }

#define HOST_GUARD(body)                                   \
    try {                                                  \
        body                                               \
    } catch (ibis_host::AbiError & e) {                    \
        g_host_error = e.what();                           \
        return e.code;                                     \
    } catch (std::exception & e) {                         \
        g_host_error = e.what();                           \
        return IBIS_B200_EINVAL;                           \
    }

int ibis_b200_host_run_rk(ibis_b200_ctx* ctx, const ibis_b200_rk_config* cfg, const double* prim0,
                          double* prim_out, ibis_b200_solve_report* report, const ibis_b200_observers* obs) {
    if (!ctx || !cfg || !prim0) { g_host_error = "null argument"; return IBIS_B200_EINVAL; }
    HOST_GUARD({
        ibis_host::RungeKutta solver(ctx, *cfg, prim0, prim_out, obs);
        int rc = solver.solve();
        if (report) *report = solver.report;
        return rc;
    })
}

int ibis_b200_host_run_steady(ibis_b200_ctx* ctx, const ibis_b200_steady_config* cfg, const double* prim0,
                              double* prim_out, ibis_b200_solve_report* report, const ibis_b200_observers* obs) {
    if (!ctx || !cfg || !prim0) { g_host_error = "null argument"; return IBIS_B200_EINVAL; }
    HOST_GUARD({
        ibis_host::SteadyState solver(ctx, sizes_of(ctx), *cfg, prim0, prim_out, obs);
        int rc = solver.solve();
        if (report) *report = solver.report;
        return rc;
    })
}

int ibis_b200_host_gmres_dense(ibis_b200_ctx* ctx, const ibis_b200_linear_solver_config* cfg, const double* a,
                               const double* b, int n, double* x, int* n_iters, double* residual) {
    if (!ctx || !cfg || !a || !b || !x || n < 1) { g_host_error = "null argument"; return IBIS_B200_EINVAL; }
    HOST_GUARD({
        std::shared_ptr<ibis_host::LinearSystem> sys(new DenseSystem(ctx, a, b, n));
        std::shared_ptr<ibis_host::LinearSystem> pre;
        if (cfg->type == IBIS_B200_LINEAR_FGMRES) pre.reset(new DenseSystem(ctx, a, b, n));
        int xh = 0;
        ibis_host::check(ctx, ibis_b200_vec_create(ctx, n, &xh), "vec_create");
        ibis_host::check(ctx, ibis_b200_vec_upload(ctx, xh, x), "vec_upload");
        ibis_host::LinearSolveResult r;
        {
            auto solver = ibis_host::make_linear_solver(ctx, sys, pre, *cfg);
            r = solver->solve(xh);
        }
        ibis_host::check(ctx, ibis_b200_vec_download(ctx, xh, x), "vec_download");
        ibis_b200_free(ctx, xh);
        if (n_iters) *n_iters = (int)r.n_iters;
        if (residual) *residual = r.residual;
        return r.success ? 1 : 0;
    })
}

}  // extern "C"
