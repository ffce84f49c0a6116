// solvers.hpp -- the reference's host-side solver classes, re-expressed above the C ABI of
// libibis_b200.so.  Same class names, same control flow and the same floating-point
// operations on the host as the reference (paths relative to /root/reference/src); every
// Kokkos kernel they launch there is one ABI call here.  Device data are the context's
// integer buffer handles.
#pragma once

#include <stdint.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ibis_b200.h"
#include "ibis_b200_host.h"

namespace ibis_host {

struct AbiError : std::runtime_error {
    int code;
    AbiError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// throws AbiError with ibis_b200_last_error() when an ABI call fails
void check(ibis_b200_ctx* ctx, int rc, const char* what);

// ---- solvers/cfl.h, cfl.cpp:11-84 ---------------------------------------------------
class CflSchedule {
public:
    virtual ~CflSchedule() {}
    virtual double eval(double t) = 0;
    virtual bool residual_based() const { return false; }
};
class ConstantSchedule : public CflSchedule {
public:
    explicit ConstantSchedule(double cfl) : cfl_(cfl) {}
    double eval(double) override { return cfl_; }
private:
    double cfl_;
};
class LinearSchedule : public CflSchedule {
public:
    LinearSchedule(std::vector<double> times, std::vector<double> cfls) : times_(times), cfls_(cfls) {}
    double eval(double t) override;
private:
    std::vector<double> times_, cfls_;
};
class ResidualBasedCfl : public CflSchedule {
public:
    ResidualBasedCfl(double threshold, double power, double start_cfl, double max_cfl)
        : threshold_(threshold), power_(power), start_cfl_(start_cfl), max_cfl_(max_cfl),
          previous_cfl_(start_cfl), previous_residual_(1.0) {}
    double eval(double t) override;
    bool residual_based() const override { return true; }
private:
    double threshold_, power_, start_cfl_, max_cfl_, previous_cfl_, previous_residual_;
};
std::unique_ptr<CflSchedule> make_cfl_schedule(const ibis_b200_cfl_config& c);

// ---- linear_algebra/linear_system.h:8-32 ---------------------------------------------
// Vectors are device-vector handles of the context.
class LinearSystem {
public:
    virtual ~LinearSystem() {}
    virtual void matrix_vector_product(int vec, int result) = 0;
    virtual void eval_rhs() = 0;
    virtual void set_rhs(int rhs) = 0;
    virtual int rhs() = 0;
    virtual size_t num_vars() const = 0;
};
class PseudoTransientLinearSystem : public LinearSystem {
public:
    virtual void set_pseudo_time_step(double dt_star) = 0;
    virtual std::shared_ptr<PseudoTransientLinearSystem> preconditioner() = 0;
};

// ---- linear_algebra/gmres.h, gmres.cpp -------------------------------------------------
struct LinearSolveResult {
    bool success = false;
    size_t n_iters = 0;
    double tol = -1.0;
    double residual = -1.0;
};
class IterativeLinearSolver {
public:
    virtual ~IterativeLinearSolver() {}
    virtual LinearSolveResult solve(int x) = 0;
};

// The host-side least-squares problem of GMRES (gmres.cpp:22-74): Hessenberg matrix rotated
// progressively into upper-triangular form with plane rotations.
class HessenbergQR {
public:
    explicit HessenbergQR(size_t max_iters);
    void reset();
    void set_g0(double beta) { g_[0] = beta; }
    // column j of the Hessenberg matrix, h[0..j+1]; returns abs(g(j+1))
    double add_column(size_t j, const double* h);
    // y = H(0:n,0:n)^-1 g(0:n)   (upper_triangular_solve, dense_linear_algebra.h:280-297)
    void solve(size_t n, double* y) const;
private:
    size_t m_;                      // leading dimension (max_iters + 2)
    std::vector<double> H_, Q_, g_, hr_;
    double& H(size_t r, size_t c) { return H_[r * m_ + c]; }
    double H(size_t r, size_t c) const { return H_[r * m_ + c]; }
    double& Q(size_t r, size_t c) { return Q_[r * m_ + c]; }
};

class Gmres : public IterativeLinearSolver {
public:
    Gmres(ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, size_t max_iters, double tol);
    ~Gmres();
    LinearSolveResult solve(int x0) override;
private:
    ibis_b200_ctx* ctx_;
    std::shared_ptr<LinearSystem> system_;
    size_t num_vars_, max_iters_;
    double tol_;
    HessenbergQR qr_;
    int krylov_vectors_ = 0, r0_ = 0, w_ = 0, v_ = 0;
    std::vector<double> h_, ym_;
};

class FGmres : public IterativeLinearSolver {
public:
    FGmres(ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, size_t max_iters, double tol,
           std::shared_ptr<LinearSystem> precondition_system, size_t max_precondition_iters,
           double precondition_tol);
    ~FGmres();
    LinearSolveResult solve(int x) override;
private:
    ibis_b200_ctx* ctx_;
    std::shared_ptr<LinearSystem> system_, precondition_system_;
    size_t num_vars_, max_iters_;
    double tol_;
    HessenbergQR qr_;
    Gmres preconditioner_;
    int krylov_vectors_ = 0, preconditioned_krylov_vectors_ = 0, r0_ = 0, w_ = 0, v_ = 0, z_ = 0;
    std::vector<double> h_, ym_;
};

std::unique_ptr<IterativeLinearSolver> make_linear_solver(
    ibis_b200_ctx* ctx, std::shared_ptr<LinearSystem> system, std::shared_ptr<LinearSystem> preconditioner,
    const ibis_b200_linear_solver_config& config);

// ---- solvers/steady_state.h, steady_state.cpp:18-170 -----------------------------------
// Sim<dual> (simulation/simulation.h) is the ibis_b200_ctx; residuals / cq / fs are dual
// ConservedQuantities / FlowStates handles shared with SteadyState.
struct SimSizes {
    int dim;
    int64_t n_cells, n_total;
};
class SteadyStateLinearisation : public PseudoTransientLinearSystem {
public:
    SteadyStateLinearisation(ibis_b200_ctx* ctx, SimSizes sizes, int residuals, int cq, int fs,
                             bool allow_reconstruction = true);
    ~SteadyStateLinearisation();
    void matrix_vector_product(int vec, int result) override;
    void eval_rhs() override;
    void set_rhs(int rhs) override { rhs_ = rhs; }
    int rhs() override { return rhs_; }
    size_t num_vars() const override { return (size_t)(n_cells_ * n_cons_); }
    void set_pseudo_time_step(double dt_star) override { dt_star_ = dt_star; }
    std::shared_ptr<PseudoTransientLinearSystem> preconditioner() override;
private:
    ibis_b200_ctx* ctx_;
    SimSizes sizes_;
    int64_t n_cells_, n_cons_;
    int residuals_, cq_, fs_;
    bool allow_reconstruction_;
    double dt_star_ = 0.0;
    int rhs_ = 0, own_rhs_ = 0, cq_tmp_ = 0, fs_tmp_ = 0;
};

// ---- solvers/jfnk.h, jfnk.cpp ----------------------------------------------------------
struct Norms6 {
    double v[6] = {0, 0, 0, 0, 0, 0};
    double global() const { return v[0]; }
};
class Jfnk {
public:
    Jfnk(ibis_b200_ctx* ctx, SimSizes sizes, std::shared_ptr<PseudoTransientLinearSystem> system,
         std::unique_ptr<CflSchedule>&& cfl, int residuals, const ibis_b200_steady_config& config);
    ~Jfnk();
    int initialise();
    LinearSolveResult step(int cq, int fs, size_t step);
    size_t max_steps() const { return max_steps_; }
    double pseudo_time_step_size() const { return stable_dt_; }
    double calculate_cfl(size_t step) const;
    Norms6 residual_norms() const { return residual_norms_; }
    Norms6 relative_residual_norms() const;
    double target_residual() const { return tolerance_; }
    const LinearSolveResult& last_gmres_result() const { return last_gmres_result_; }
private:
    ibis_b200_ctx* ctx_;
    std::shared_ptr<PseudoTransientLinearSystem> system_, preconditioner_;
    std::unique_ptr<CflSchedule> cfl_;
    std::unique_ptr<IterativeLinearSolver> gmres_;
    int dU_ = 0;
    size_t max_steps_;
    double tolerance_, stable_dt_ = 0.0;
    int residuals_;
    Norms6 residual_norms_, initial_residual_norms_;
    LinearSolveResult last_gmres_result_;
    bool residual_based_cfl_;
    void set_pseudo_time_step_size(double dt_star);
    void apply_update_(int cq, int fs);
};

// ---- solvers/solver.h, solver.cpp:18-64 ------------------------------------------------
class Solver {
public:
    virtual ~Solver() {}
    int solve();
    ibis_b200_solve_report report;
protected:
    virtual int initialise() = 0;
    virtual int finalise() = 0;
    virtual int take_step(size_t step) = 0;
    virtual bool print_this_step(unsigned int step) = 0;
    virtual bool plot_this_step(unsigned int step) = 0;
    virtual bool residuals_this_step(unsigned int step) = 0;
    virtual bool write_residuals(unsigned int step, double wc) = 0;
    virtual int plot_solution(unsigned int step) = 0;
    virtual void print_progress(unsigned int step, double wc) = 0;
    virtual bool stop_now(unsigned int step) = 0;
    virtual int stop_reason(unsigned int step) = 0;
    virtual size_t max_step() const = 0;
    virtual int count_bad_cells() = 0;
    virtual void fill_report() = 0;
};

// ---- solvers/runge_kutta.h, runge_kutta.cpp ----------------------------------------------
class ButcherTableau {
public:
    ButcherTableau() {}
    ButcherTableau(int n_stages, const double* a, const double* b);
    size_t num_stages() const { return b_.size(); }
    double a(size_t i, size_t j) const { return a_[i - 1][j]; }   // runge_kutta.h: a(i, j), i >= 1
    double b(size_t i) const { return b_[i]; }
    const std::vector<double>& a_flat() const { return a_flat_; }
    const std::vector<double>& b_vec() const { return b_; }
private:
    std::vector<std::vector<double>> a_;
    std::vector<double> a_flat_, b_;
};

class RungeKutta : public Solver {
public:
    RungeKutta(ibis_b200_ctx* ctx, const ibis_b200_rk_config& config, const double* prim0, double* prim_out,
               const ibis_b200_observers* obs);
    ~RungeKutta();
protected:
    int initialise() override;
    int finalise() override;
    int take_step(size_t step) override;
    bool print_this_step(unsigned int step) override;
    bool plot_this_step(unsigned int step) override;
    bool residuals_this_step(unsigned int step) override;
    bool write_residuals(unsigned int step, double wc) override;
    int plot_solution(unsigned int step) override;
    void print_progress(unsigned int, double) override {}
    bool stop_now(unsigned int step) override;
    int stop_reason(unsigned int step) override;
    size_t max_step() const override { return max_step_; }
    int count_bad_cells() override;
    void fill_report() override;
private:
    ibis_b200_ctx* ctx_;
    const double* prim0_;
    double* prim_out_;
    ibis_b200_observers obs_;
    int flow_ = 0, conserved_quantities_ = 0, k_tmp_ = 0, flow_tmp_ = 0;
    std::vector<int> k_, k_own_;
    ButcherTableau tableau_;
    std::unique_ptr<CflSchedule> cfl_;
    bool fused_;
    bool moving_grid_ = false;
    double t_ = 0.0, dt_ = 0.0, stable_dt_ = 0.0, max_time_, dt_init_;
    size_t max_step_;
    unsigned int print_frequency_;
    int plot_every_n_steps_, residuals_every_n_steps_;
    double plot_frequency_, residual_frequency_;
    double time_since_last_plot_ = 0.0, time_since_last_residual_ = 0.0;
    int64_t last_bad_ = 0;
    void estimate_dt();
    void function_eval_(int fs, int cq, size_t index);
};

// ---- solvers/steady_state.h: SteadyState ------------------------------------------------
class SteadyState : public Solver {
public:
    SteadyState(ibis_b200_ctx* ctx, SimSizes sizes, const ibis_b200_steady_config& config, const double* prim0,
                double* prim_out, const ibis_b200_observers* obs);
    ~SteadyState();
protected:
    int initialise() override;
    int finalise() override { return 0; }
    int take_step(size_t step) override;
    bool print_this_step(unsigned int step) override;
    bool plot_this_step(unsigned int step) override;
    bool residuals_this_step(unsigned int step) override;
    bool write_residuals(unsigned int step, double wc) override;
    int plot_solution(unsigned int step) override;
    void print_progress(unsigned int, double) override {}
    bool stop_now(unsigned int step) override;
    int stop_reason(unsigned int step) override;
    size_t max_step() const override { return jfnk_->max_steps(); }
    int count_bad_cells() override;
    void fill_report() override;
private:
    ibis_b200_ctx* ctx_;
    SimSizes sizes_;
    const double* prim0_;
    double* prim_out_;
    ibis_b200_observers obs_;
    int fs_ = 0, cq_ = 0, residuals_ = 0;
    std::unique_ptr<Jfnk> jfnk_;
    unsigned int print_frequency_, plot_frequency_, diagnostics_frequency_;
    int64_t linear_iters_ = 0, last_bad_ = 0;
};

}  // namespace ibis_host
