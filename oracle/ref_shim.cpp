// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A thin extern "C" window onto the UNMODIFIED reference (wattrg/ibis) compiled
// from /root/reference by oracle/ref_build/Makefile into
// oracle/_ref/libibis_ref.so.  Everything numerical below is a call into the
// reference's own classes (GridBlock, FiniteVolume, RungeKutta, SteadyState,
// Gmres ...); this file only moves arrays in and out so that
//   * tests/ can compare the CUDA path with the real reference,
//   * tests/golden/ fixtures can be generated (tests/golden/make_golden.py),
//   * bench.py --impl reference can time the reference's Kokkos-OpenMP path.
// Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
// legs may load this library.
//
// State layout used at this boundary: prim[n_total][7] = rho,p,T,e,vx,vy,vz
// (reference FlowStates: gas/gas_state.h:39-45 + vel), cons[n][dim+2].
// For T = dual, "re" and "du" arrays are passed separately.

#include <finite_volume/conserved_quantities.h>
#include <finite_volume/finite_volume.h>
#include <finite_volume/grid_motion_driver.h>
#include <finite_volume/primative_conserved_conversion.h>
#include <gas/flow_state.h>
#include <gas/transport_properties.h>
#include <grid/gradient.h>
#include <grid/grid.h>
#include <linear_algebra/gmres.h>
#include <simulation/simulation.h>
#include <solvers/solver.h>
#include <solvers/steady_state.h>
#include <spdlog/spdlog.h>
#include <unistd.h>

#include <Kokkos_Core.hpp>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <nlohmann/json.hpp>
#include <string>

using json = nlohmann::json;
using Ibis::dual;
using Ibis::real;

namespace {

inline double re(const double& x) { return x; }
inline double re(const dual& x) { return x.real(); }
inline double du(const double&) { return 0.0; }
inline double du(const dual& x) { return x.dual(); }
inline void set(double& x, double r, double) { x = r; }
inline void set(dual& x, double r, double d) { x = dual(r, d); }

struct CaseBase {
    virtual ~CaseBase() {}
    bool is_dual = false;
};

template <typename T>
struct Case : CaseBase {
    using scalar = T;
    bool motion_set = false;
    json config;
    GridBlock<T> grid;
    FiniteVolume<T> fv;
    IdealGas<T> gm;
    TransportProperties<T> tp;
    FlowStates<T> fs;
    ConservedQuantities<T> dudt;
    ConservedQuantities<T> cq;

    Case(const std::string& su2, const json& cfg) : config(cfg) {
        json grid_cfg = config.at("grid");
        grid = GridBlock<T>(su2, grid_cfg);
        fv = FiniteVolume<T>(grid, config);
        gm = IdealGas<T>(config.at("gas_model"));
        tp = TransportProperties<T>(config.at("transport_properties"));
        fs = FlowStates<T>(grid.num_total_cells());
        dudt = ConservedQuantities<T>(grid.num_total_cells(), grid.dim());
        cq = ConservedQuantities<T>(grid.num_total_cells(), grid.dim());
    }

    void load_prim(const double* pr, const double* pd) {
        size_t n = grid.num_total_cells();
        for (size_t i = 0; i < n; i++) {
            const double* r = pr + 7 * i;
            double z[7] = {0, 0, 0, 0, 0, 0, 0};
            const double* d = pd ? pd + 7 * i : z;
            set(fs.gas.rho(i), r[0], d[0]);
            set(fs.gas.pressure(i), r[1], d[1]);
            set(fs.gas.temp(i), r[2], d[2]);
            set(fs.gas.energy(i), r[3], d[3]);
            set(fs.vel.x(i), r[4], d[4]);
            set(fs.vel.y(i), r[5], d[5]);
            set(fs.vel.z(i), r[6], d[6]);
        }
    }

    void store_prim(double* pr, double* pd) {
        size_t n = grid.num_total_cells();
        for (size_t i = 0; i < n; i++) {
            T v[7] = {fs.gas.rho(i),   fs.gas.pressure(i), fs.gas.temp(i), fs.gas.energy(i),
                      fs.vel.x(i),     fs.vel.y(i),        fs.vel.z(i)};
            for (int k = 0; k < 7; k++) {
                if (pr) pr[7 * i + k] = re(v[k]);
                if (pd) pd[7 * i + k] = du(v[k]);
            }
        }
    }
};

template <typename F>
auto dispatch(void* h, F&& f) {
    CaseBase* b = static_cast<CaseBase*>(h);
    if (b->is_dual) return f(*static_cast<Case<dual>*>(b));
    return f(*static_cast<Case<real>*>(b));
}

template <typename T>
void copy_vec3(const Vector3s<T>& v, size_t n, double* out) {
    if (!out) return;
    for (size_t i = 0; i < n; i++) {
        out[3 * i + 0] = re(v.x(i));
        out[3 * i + 1] = re(v.y(i));
        out[3 * i + 2] = re(v.z(i));
    }
}

bool g_initialised = false;

}  // namespace

extern "C" {

int ref_init(int num_threads) {
    if (g_initialised) return 0;
    Kokkos::InitializationSettings s;
    if (num_threads > 0) s.set_num_threads(num_threads);
    s.set_disable_warnings(true);
    Kokkos::initialize(s);
    spdlog::set_level(spdlog::level::warn);
    g_initialised = true;
    return 0;
}

int ref_num_threads() { return Kokkos::DefaultExecutionSpace().concurrency(); }

void ref_finalize() {
    if (g_initialised) Kokkos::finalize();
    g_initialised = false;
}

void* ref_case_create(const char* su2_path, const char* config_json, int is_dual) {
    try {
        json cfg = json::parse(config_json);
        CaseBase* c;
        if (is_dual)
            c = new Case<dual>(su2_path, cfg);
        else
            c = new Case<real>(su2_path, cfg);
        c->is_dual = is_dual != 0;
        return c;
    } catch (std::exception& e) {
        fprintf(stderr, "ref_case_create: %s\n", e.what());
        return nullptr;
    } catch (std::runtime_error* e) {
        fprintf(stderr, "ref_case_create: %s\n", e->what());
        return nullptr;
    }
}

void ref_case_destroy(void* h) { delete static_cast<CaseBase*>(h); }

// out = {dim, n_cells, n_ghost, n_faces, n_vertices, n_cell_face_entries, n_tags, n_cons}
void ref_case_sizes(void* h, int64_t* out) {
    dispatch(h, [&](auto& c) {
        out[0] = c.grid.dim();
        out[1] = c.grid.num_cells();
        out[2] = c.grid.num_ghost_cells();
        out[3] = c.grid.num_interfaces();
        out[4] = c.grid.num_vertices();
        out[5] = c.grid.cells().faces().num_face_ids();
        out[6] = c.grid.boundary_tags().size();
        out[7] = c.grid.dim() + 2;
        return 0;
    });
}

void ref_grid_faces(void* h, int64_t* left, int64_t* right, double* area, double* norm,
                    double* tan1, double* tan2, double* centre) {
    dispatch(h, [&](auto& c) {
        auto f = c.grid.interfaces();
        size_t nf = c.grid.num_interfaces();
        for (size_t i = 0; i < nf; i++) {
            left[i] = (int64_t)f.left_cell(i);
            right[i] = (int64_t)f.right_cell(i);
            area[i] = re(f.area(i));
        }
        copy_vec3(f.norm(), nf, norm);
        copy_vec3(f.tan1(), nf, tan1);
        copy_vec3(f.tan2(), nf, tan2);
        copy_vec3(f.centre(), nf, centre);
        return 0;
    });
}

void ref_grid_cells(void* h, int64_t* offsets, int64_t* face_ids, int32_t* outsigns,
                    int64_t* neighbours, double* volume, double* centroid) {
    dispatch(h, [&](auto& c) {
        auto cells = c.grid.cells();
        auto cf = cells.faces();
        size_t n = c.grid.num_cells();
        size_t k = 0;
        offsets[0] = 0;
        for (size_t i = 0; i < n; i++) {
            auto ids = cf.face_ids(i);
            auto os = cf.outsigns(i);
            for (size_t j = 0; j < ids.size(); j++, k++) {
                face_ids[k] = (int64_t)ids(j);
                outsigns[k] = os(j);
                neighbours[k] = (int64_t)cells.neighbour_cells(i, j);
            }
            offsets[i + 1] = (int64_t)k;
            volume[i] = re(cells.volume(i));
        }
        copy_vec3(cells.centroids(), c.grid.num_total_cells(), centroid);
        return 0;
    });
}

// boundary tag i (alphabetical order = the order FiniteVolume applies BCs in)
int64_t ref_grid_tag(void* h, int i, char* name, int name_cap) {
    return dispatch(h, [&](auto& c) {
        std::string tag = c.grid.boundary_tags()[i];
        snprintf(name, name_cap, "%s", tag.c_str());
        return (int64_t)c.grid.boundary_faces(tag).size();
    });
}

void ref_grid_tag_faces(void* h, int i, int64_t* faces, int64_t* ghosts) {
    dispatch(h, [&](auto& c) {
        std::string tag = c.grid.boundary_tags()[i];
        auto bf = c.grid.boundary_faces(tag);
        auto gc = c.grid.ghost_cells(tag);
        for (size_t k = 0; k < bf.size(); k++) faces[k] = (int64_t)bf(k);
        for (size_t k = 0; k < gc.size(); k++) ghosts[k] = (int64_t)gc(k);
        return 0;
    });
}

// WLS r-factors (n_cells x 3 or 6); returns number of columns, 0 if not allocated
int ref_grid_wls_r(void* h, double* out) {
    return dispatch(h, [&](auto& c) {
        if (!c.grid.grad_calc_) return 0;
        auto r = c.grid.grad_calc().r_;
        int nc = (int)r.extent(1);
        for (size_t i = 0; i < r.extent(0); i++)
            for (int j = 0; j < nc; j++) out[i * nc + j] = re(r(i, j));
        return nc;
    });
}

// One residual evaluation: FiniteVolume<T>::compute_dudt (finite_volume.cpp:57-96).
// prim is in/out (ghost rows are rewritten by the boundary conditions).
int ref_compute_dudt(void* h, double* prim_re, double* prim_du, double* dudt_re,
                     double* dudt_du, int allow_reconstruction) {
    return dispatch(h, [&](auto& c) {
        c.load_prim(prim_re, prim_du);
        int rc = (int)c.fv.compute_dudt(c.fs, c.grid, c.dudt, c.gm, c.tp,
                                        allow_reconstruction != 0);
        c.store_prim(prim_re, prim_du);
        size_t n = c.grid.num_cells();
        size_t nc = c.dudt.n_conserved();
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < nc; k++) {
                if (dudt_re) dudt_re[i * nc + k] = re(c.dudt(i, k));
                if (dudt_du) dudt_du[i * nc + k] = du(c.dudt(i, k));
            }
        return rc;
    });
}

// The moving-grid overload FiniteVolume<T>::compute_dudt(fs, vertex_vel, cq, grid, dudt, ...)
// (finite_volume.cpp:57-85) with the motion driver of config.grid.motion, as RungeKutta wires it
// (runge_kutta.cpp:58-66,121-129); cq = primatives_to_conserved(fs).  face_vel (n_faces x 3) receives
// GridBlock::face_vel() as the evaluation leaves it.
int ref_compute_dudt_moving(void* h, double* prim_re, double* prim_du, double* dudt_re,
                            double* dudt_du, int allow_reconstruction, double* face_vel) {
    return dispatch(h, [&](auto& c) {
        using T = typename std::decay_t<decltype(c)>::scalar;
        if (!c.grid.moving()) return -1;
        if (!c.motion_set) {
            auto driver = build_grid_motion_driver<T>(c.grid, c.config.at("grid").at("motion"));
            c.grid.set_motion_driver(driver);
            c.motion_set = true;
        }
        c.load_prim(prim_re, prim_du);
        primatives_to_conserved(c.cq, c.fs, c.gm);
        Vector3s<T> vertex_vel(c.grid.num_vertices());
        int rc = (int)c.fv.compute_dudt(c.fs, vertex_vel, c.cq, c.grid, c.dudt, c.gm, c.tp,
                                        allow_reconstruction != 0);
        c.store_prim(prim_re, prim_du);
        size_t n = c.grid.num_cells();
        size_t nc = c.dudt.n_conserved();
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < nc; k++) {
                if (dudt_re) dudt_re[i * nc + k] = re(c.dudt(i, k));
                if (dudt_du) dudt_du[i * nc + k] = du(c.dudt(i, k));
            }
        copy_vec3(c.grid.face_vel(), c.grid.num_interfaces(), face_vel);
        return rc;
    });
}

// time `nrep` back-to-back residual evaluations on the state last loaded; seconds
double ref_time_dudt(void* h, double* prim_re, int nrep, int allow_reconstruction) {
    return dispatch(h, [&](auto& c) {
        c.load_prim(prim_re, nullptr);
        c.fv.compute_dudt(c.fs, c.grid, c.dudt, c.gm, c.tp, allow_reconstruction != 0);
        Kokkos::fence();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < nrep; i++)
            c.fv.compute_dudt(c.fs, c.grid, c.dudt, c.gm, c.tp, allow_reconstruction != 0);
        Kokkos::fence();
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    });
}

// intermediates of the last evaluation, real parts (any pointer may be NULL):
// grad_* [n_cells x 3], lim_* [n_cells], flux [n_faces x n_cons]
void ref_intermediates(void* h, double* grad_rho, double* grad_u, double* grad_p,
                       double* grad_T, double* grad_vx, double* grad_vy, double* grad_vz,
                       double* lim_rho, double* lim_u, double* lim_p, double* lim_T,
                       double* lim_vx, double* lim_vy, double* lim_vz, double* flux) {
    dispatch(h, [&](auto& c) {
        size_t n = c.grid.num_cells();
        auto& g = c.fv.cell_grad_;
        auto cp3 = [&](auto& v, double* out) {
            if (out && v.view_.extent(0) == n) copy_vec3(v, n, out);
        };
        cp3(g.rho, grad_rho);
        cp3(g.u, grad_u);
        cp3(g.p, grad_p);
        cp3(g.temp, grad_T);
        cp3(g.vx, grad_vx);
        cp3(g.vy, grad_vy);
        cp3(g.vz, grad_vz);
        auto& l = c.fv.convective_flux_.limiters_;
        auto cp1 = [&](auto& v, double* out) {
            if (out && v.size() == n)
                for (size_t i = 0; i < n; i++) out[i] = re(v(i));
        };
        cp1(l.rho, lim_rho);
        cp1(l.u, lim_u);
        cp1(l.p, lim_p);
        cp1(l.temp, lim_T);
        cp1(l.vx, lim_vx);
        cp1(l.vy, lim_vy);
        cp1(l.vz, lim_vz);
        if (flux) {
            size_t nf = c.grid.num_interfaces();
            size_t nc = c.fv.flux_.n_conserved();
            for (size_t i = 0; i < nf; i++)
                for (size_t k = 0; k < nc; k++) flux[i * nc + k] = re(c.fv.flux_(i, k));
        }
        return 0;
    });
}

// FiniteVolume<T>::estimate_dt (finite_volume.cpp:131-184) on the given state
double ref_estimate_dt(void* h, double* prim_re) {
    return dispatch(h, [&](auto& c) {
        c.load_prim(prim_re, nullptr);
        return (double)c.fv.estimate_dt(c.fs, c.grid, c.gm, c.tp);
    });
}

int64_t ref_count_bad_cells(void* h, double* prim_re) {
    return dispatch(h, [&](auto& c) {
        c.load_prim(prim_re, nullptr);
        return (int64_t)c.fv.count_bad_cells(c.fs, c.grid.num_cells());
    });
}

// primatives_to_conserved / conserved_to_primatives
// (primative_conserved_conversion.cpp:5-57) over all n_total rows
void ref_prim_to_cons(void* h, double* prim_re, double* prim_du, double* cons_re,
                      double* cons_du) {
    dispatch(h, [&](auto& c) {
        c.load_prim(prim_re, prim_du);
        primatives_to_conserved(c.cq, c.fs, c.gm);
        size_t n = c.grid.num_total_cells(), nc = c.cq.n_conserved();
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < nc; k++) {
                if (cons_re) cons_re[i * nc + k] = re(c.cq(i, k));
                if (cons_du) cons_du[i * nc + k] = du(c.cq(i, k));
            }
        return 0;
    });
}

void ref_cons_to_prim(void* h, double* cons_re, double* cons_du, double* prim_re,
                      double* prim_du) {
    dispatch(h, [&](auto& c) {
        size_t n = c.grid.num_total_cells(), nc = c.cq.n_conserved();
        for (size_t i = 0; i < n; i++)
            for (size_t k = 0; k < nc; k++)
                set(c.cq(i, k), cons_re[i * nc + k], cons_du ? cons_du[i * nc + k] : 0.0);
        conserved_to_primatives(c.cq, c.fs, c.gm);
        c.store_prim(prim_re, prim_du);
        return 0;
    });
}

// ConservedQuantities<T>::L2_norms (conserved_quantities.cpp:41-69) of the last dudt:
// out = {global, mass, mom_x, mom_y, mom_z, energy}
void ref_l2_norms(void* h, double* out) {
    dispatch(h, [&](auto& c) {
        auto n = c.dudt.L2_norms();
        out[0] = re(n.global());
        out[1] = re(n.mass());
        out[2] = re(n.momentum_x());
        out[3] = re(n.momentum_y());
        out[4] = re(n.momentum_z());
        out[5] = re(n.energy());
        return 0;
    });
}

// The whole `ibis run` (commands/run/run.cpp:36-64 minus CLI): cwd must hold
// config/config.json, io/grid, io/flow, log/.
int ref_solve(const char* case_dir) {
    char old[4096];
    if (!getcwd(old, sizeof old)) return -1;
    if (chdir(case_dir) != 0) return -2;
    int rc = 0;
    try {
        std::ifstream f("config/config.json");
        json config = json::parse(f);
        auto solver = make_solver(config, "io/grid", "io/flow");
        rc = solver->solve();
    } catch (std::exception& e) {
        fprintf(stderr, "ref_solve: %s\n", e.what());
        rc = -3;
    } catch (std::runtime_error* e) {
        fprintf(stderr, "ref_solve: %s\n", e->what());
        rc = -3;
    }
    if (chdir(old) != 0) return -4;
    return rc;
}

// ---- linear algebra known-answer access: Gmres on a dense host matrix ----
// (mirrors the reference's own TestLinearSystem, linear_algebra/gmres.cpp:321-385)
struct DenseSystem : LinearSystem {
    Ibis::Matrix<real> A;
    Ibis::Vector<real> b;
    size_t n;
    DenseSystem(const double* a, const double* rhs, size_t n_) : n(n_) {
        A = Ibis::Matrix<real>("A", n, n);
        b = Ibis::Vector<real>("b", n);
        for (size_t i = 0; i < n; i++) {
            b(i) = rhs[i];
            for (size_t j = 0; j < n; j++) A(i, j) = a[i * n + j];
        }
    }
    void matrix_vector_product(Ibis::Vector<real>& vec, Ibis::Vector<real>& result) {
        Ibis::gemv(A, vec, result);
    }
    std::unique_ptr<LinearSystem> preconditioner() { return nullptr; }
    void eval_rhs() {}
    void set_rhs(Ibis::Vector<real>& rhs) { b = rhs; }
    real& rhs(const size_t i) const { return b(i); }
    real& rhs(const size_t cell_i, const size_t cons_i) const { return b(cell_i + cons_i); }
    Ibis::Vector<real>& rhs() { return b; }
    size_t num_vars() const { return n; }
};

int ref_gmres_dense(const double* a, const double* rhs, int n, int max_iters, double tol,
                    double* x, int* n_iters, double* residual) {
    auto sys = std::shared_ptr<LinearSystem>(new DenseSystem(a, rhs, n));
    Gmres solver(sys, (size_t)max_iters, tol);
    Ibis::Vector<real> x0("x", n);
    for (int i = 0; i < n; i++) x0(i) = x[i];
    auto result = solver.solve(x0);
    for (int i = 0; i < n; i++) x[i] = x0(i);
    *n_iters = (int)result.n_iters;
    *residual = result.residual;
    return result.success ? 1 : 0;
}

}  // extern "C"
