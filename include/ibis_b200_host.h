/* ibis_b200_host.h -- C entry points of libibis_b200_host.so: the host side ABOVE the C ABI of
 * include/ibis_b200.h.  The library is C++ (ibis_b200/host/), because the reference's host code
 * is C++: it mirrors the reference's solver classes -- Solver::solve, RungeKutta, SteadyState,
 * SteadyStateLinearisation, Jfnk, Gmres, FGmres, CflSchedule -- with every Kokkos kernel replaced
 * by a call into libibis_b200.so.  These extern "C" functions only construct those classes from
 * POD configuration (config/config.json "solver" object flattened) and run
 * Solver::solve(); they exist so that tests and bench can drive the C++ host code without a JSON
 * dependency.  Paths cited are relative to /root/reference/src.
 */
#ifndef IBIS_B200_HOST_H
#define IBIS_B200_HOST_H

#include <stdint.h>

#include "ibis_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* solvers/cfl.cpp:72-84 (make_cfl_schedule) */
enum { IBIS_B200_CFL_CONSTANT = 0, IBIS_B200_CFL_LINEAR_INTERPOLATE = 1, IBIS_B200_CFL_RESIDUAL_BASED = 2 };
typedef struct {
    int32_t type;
    int32_t n;                 /* linear_interpolate: number of (time, cfl) points */
    double value;              /* constant */
    const double* times;       /* linear_interpolate */
    const double* cfls;
    double growth_threshold, power, start_cfl, max_cfl;   /* residual_based (cfl.cpp:57-70) */
} ibis_b200_cfl_config;

/* solvers/runge_kutta.cpp:21-37 */
typedef struct {
    int32_t n_stages;
    int32_t fused_step;        /* 1: take_step = one ibis_b200_rk_step call (device-resident dt);
                                  0: the reference's call sequence, one ABI call per kernel */
    const double* a;           /* butcher_tableau.a rows 1..stages-1 concatenated (row i has i entries) */
    const double* b;
    ibis_b200_cfl_config cfl;
    double max_time;
    int64_t max_step;
    double dt_init;
    int32_t print_frequency;
    int32_t plot_every_n_steps;
    double plot_frequency;
    int32_t residuals_every_n_steps;
    int32_t distinct_stages;   /* 0 (default): k_[i] alias one array like the reference
                                  (runge_kutta.cpp:49-50, see ibis_b200_rk_distinct_stages) */
    double residual_frequency;
    int32_t moving_grid;       /* grid.motion.enabled (runge_kutta.cpp:58-70): the context must hold face / cell
                                  vertices, vertex positions and the motion driver's vertex velocities (rigid body
                                  translation: one constant vector); forces the unfused call sequence */
    int32_t reserved;
} ibis_b200_rk_config;

/* linear_algebra/gmres.cpp:305-318 (make_linear_solver) */
enum { IBIS_B200_LINEAR_GMRES = 0, IBIS_B200_LINEAR_FGMRES = 1 };
typedef struct {
    int32_t type;
    int32_t max_iters;
    double tol;                            /* gmres: "tol"; fgmres: "tolerance" */
    int32_t max_preconditioner_iters;      /* fgmres */
    int32_t reserved;
    double preconditioner_tolerance;
} ibis_b200_linear_solver_config;

/* solvers/steady_state.cpp:176-221, solvers/jfnk.cpp:7-23 */
typedef struct {
    ibis_b200_cfl_config cfl;
    ibis_b200_linear_solver_config linear_solver;
    int64_t max_steps;
    double tolerance;
    int32_t print_frequency;
    int32_t plot_frequency;
    int32_t diagnostics_frequency;
    int32_t reserved;
} ibis_b200_steady_config;

/* What Solver::solve() (solvers/solver.cpp:18-64) leaves behind. */
typedef struct {
    int32_t status;            /* 0 ok; 1 = failed step / bad cells (solver.cpp:31-43) */
    int32_t stop_reason;       /* 0 loop ran out, 1 max step, 2 max time, 3 target residual */
    int64_t steps;             /* number of take_step calls made */
    int64_t bad_cells;
    double t;                  /* RungeKutta::t_ (steady: last step index) */
    double dt;                 /* last dt_ (steady: last stable dt) */
    double stable_dt;
    double residual_norms[6];  /* last L2 norms: global, mass, mom x y z, energy */
    double relative_residual;  /* steady: global relative residual */
    int64_t linear_iters;      /* steady: total GMRES iterations */
    int32_t last_linear_success;
    int32_t last_linear_iters;
    double last_linear_residual;
    double wall_seconds;
} ibis_b200_solve_report;

/* Optional observers, called from Solver::solve at the reference's plot / residual points
 * (plot_solution, write_residuals).  fs is the FlowStates handle to download from. */
typedef struct {
    void* user;
    void (*plot)(void* user, int64_t step, double t, int fs);
    void (*residuals)(void* user, int64_t step, double t, double wall, const double norms[6]);
} ibis_b200_observers;

/* make_solver(... "runge_kutta") + Solver::solve (solver.cpp:18-84).  prim0 / prim_out are
 * FlowStates rows (n_total x 7, reference numbering); prim0 stands for FVIO::read of snapshot 0,
 * prim_out for the final plot_solution. */
int ibis_b200_host_run_rk(ibis_b200_ctx* ctx, const ibis_b200_rk_config* cfg, const double* prim0,
                          double* prim_out, ibis_b200_solve_report* report,
                          const ibis_b200_observers* obs);
/* make_solver(... "steady_state") + Solver::solve */
int ibis_b200_host_run_steady(ibis_b200_ctx* ctx, const ibis_b200_steady_config* cfg,
                              const double* prim0, double* prim_out, ibis_b200_solve_report* report,
                              const ibis_b200_observers* obs);
/* Gmres / FGmres on a dense n x n system held on the device (the reference's own
 * TestLinearSystem known-answer tests, linear_algebra/gmres.cpp:320-500).  a is row-major.
 * For fgmres the preconditioner system is the same matrix. Returns 1 converged, 0 not, <0 error. */
int ibis_b200_host_gmres_dense(ibis_b200_ctx* ctx, const ibis_b200_linear_solver_config* cfg,
                               const double* a, const double* b, int n, double* x, int* n_iters,
                               double* residual);
const char* ibis_b200_host_last_error(void);

/* ---- synthetic box grid (BASELINE.json configs[4]) -------------------------------------------
 * The GridBlock arrays (grid/grid.h:89-167) of the nx x ny x nz hex box "in unstructured form" of
 * SURVEY.md appendix B.4, built in closed form (ibis_b200/host/box_grid.cpp) instead of through
 * the reference's SU2 reader, bit-identical to it.  [k_lo, k_hi) selects the z-slab owned by this
 * rank (0, nz = the whole box); a slab comes with its two-ring halo cells (numbered after the
 * owned cells, ring 1 before ring 2) and the matching send / receive lists for
 * ibis_b200_halo_setup.  Boundary tags are, in order: bottom east north south top west. */
typedef struct ibis_b200_box ibis_b200_box;
int ibis_b200_host_box_create(ibis_b200_box** out, int nx, int ny, int nz, double lx, double ly,
                              double lz, int k_lo, int k_hi);
const ibis_b200_grid_desc* ibis_b200_host_box_desc(const ibis_b200_box* box);
/* peer_dir[p] = -1 (the rank below) / +1 (the rank above) */
int ibis_b200_host_box_halo(const ibis_b200_box* box, int* n_peers, int32_t peer_dir[2],
                            const int64_t** send_offsets, const int64_t** send_cells,
                            const int64_t** recv_offsets, const int64_t** recv_cells);
void ibis_b200_host_box_free(ibis_b200_box* box);
const char* ibis_b200_host_box_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* IBIS_B200_HOST_H */
