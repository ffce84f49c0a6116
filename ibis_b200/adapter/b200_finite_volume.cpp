// b200_finite_volume.cpp -- the drop-in: wattrg/ibis's own FiniteVolume<T>, bound to libibis_b200.so.
//
// This translation unit is compiled against the reference's UNMODIFIED headers and linked INSTEAD of the
// reference's src/finite_volume/finite_volume/finite_volume.cpp.  It defines the same class
// (finite_volume/finite_volume.h:29-113) with every method its callers use forwarding to the C ABI of
// include/ibis_b200.h, so that the reference's own solver loops -- RungeKutta::take_step
// (solvers/runge_kutta.cpp:108-129,131-203), SteadyStateLinearisation (solvers/steady_state.cpp:52-136),
// Jfnk (solvers/jfnk.cpp:47) -- its JSON configuration, IO and CFL schedules run unchanged while the residual is
// evaluated on the B200:
//
//   FiniteVolume<T>::compute_dudt    -> ibis_b200_fs_upload, ibis_b200_compute_dudt, ibis_b200_cq_download
//   FiniteVolume<T>::estimate_dt     -> ibis_b200_estimate_dt
//   FiniteVolume<T>::count_bad_cells -> ibis_b200_count_bad_cells
//
// The flow state lives in the reference's Kokkos Views here (host memory with the OpenMP backend) and travels to the
// device and back on every call: this adapter demonstrates the boundary and its parity, not the speed of the fused
// path (ibis_b200_rk_step keeps everything resident).  Build: oracle/ref_build/Makefile
// (the drop-in library next to the compiled reference); test: tests/test_gpu_dropin.py.
#include <finite_volume/finite_volume.h>
#include <ibis_b200.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using Ibis::dual;
using Ibis::real;

namespace {

inline double re(const double& x) { return x; }
inline double re(const dual& x) { return x.real(); }
inline double du(const double&) { return 0.0; }
inline double du(const dual& x) { return x.dual(); }
inline void set(double& x, double r, double) { x = r; }
inline void set(dual& x, double r, double d) { x = dual(r, d); }
template <typename T> constexpr int is_dual_v = 0;
template <> constexpr int is_dual_v<dual> = 1;

[[noreturn]] void fail(ibis_b200_ctx* ctx, const char* what) {
    throw std::runtime_error(std::string("ibis_b200: ") + what + ": " + ibis_b200_last_error(ctx));
}
void ck(ibis_b200_ctx* ctx, int rc, const char* what) {
    if (rc != 0) fail(ctx, what);
}

// config.json -> ibis_b200_config (the same flattening as ibis_b200/lib.py::flatten_config)
struct FlatConfig {
    ibis_b200_config c{};
    std::vector<ibis_b200_bc_action> actions;
    std::vector<double> tables;
};

int lookup(const std::map<std::string, int>& m, const std::string& k, const char* err) {
    auto it = m.find(k);
    if (it == m.end()) throw std::runtime_error(err);
    return it->second;
}

void flow_state_params(const json& fs, double* p) {          // rho p T e vx vy vz (gas/flow_state.h: FlowState(json))
    static const char* KEYS[7] = {"rho", "p", "T", "energy", "vx", "vy", "vz"};
    for (int k = 0; k < 7; k++) p[k] = fs.at(KEYS[k]);
}

std::unique_ptr<FlatConfig> flatten(const json& config, const std::vector<std::string>& tags) {
    static const std::map<std::string, int> FLUX = {{"hanel", IBIS_B200_FLUX_HANEL}, {"ausmdv", IBIS_B200_FLUX_AUSMDV},
                                                    {"ldfss", IBIS_B200_FLUX_LDFSS2}, {"ldfss2", IBIS_B200_FLUX_LDFSS2},
                                                    {"rusanov", IBIS_B200_FLUX_RUSANOV}};
    static const std::map<std::string, int> LIM = {{"unlimited", IBIS_B200_LIMITER_UNLIMITED},
                                                   {"barth_jespersen", IBIS_B200_LIMITER_BARTH_JESPERSEN}};
    static const std::map<std::string, int> THERMO = {{"rho_p", IBIS_B200_THERMO_RHO_P}, {"rho_T", IBIS_B200_THERMO_RHO_T},
                                                      {"rho_u", IBIS_B200_THERMO_RHO_U}, {"p_T", IBIS_B200_THERMO_P_T}};
    static const std::map<std::string, int> BC = {
        {"flow_state_copy", IBIS_B200_BC_FLOW_STATE_COPY}, {"internal_copy", IBIS_B200_BC_INTERNAL_COPY},
        {"internal_copy_reflect_normal", IBIS_B200_BC_INTERNAL_COPY_REFLECT_NORMAL},
        {"internal_vel_copy_reflect", IBIS_B200_BC_INTERNAL_VEL_COPY_REFLECT}, {"fix_temperature", IBIS_B200_BC_FIX_TEMPERATURE},
        {"subsonic_inflow", IBIS_B200_BC_SUBSONIC_INFLOW}, {"subsonic_outflow", IBIS_B200_BC_SUBSONIC_OUTFLOW},
        {"boundary_layer_profile", IBIS_B200_BC_BOUNDARY_LAYER_PROFILE}};
    auto out = std::make_unique<FlatConfig>();
    ibis_b200_config& c = out->c;
    const json& cf = config.at("convective_flux");
    const json& fc = cf.at("flux_calculator");
    c.flux_calculator = lookup(FLUX, fc.at("type"), "Unknown flux calculator");
    c.ldfss_delta = fc.value("delta", 2.0);
    c.reconstruction_order = cf.at("reconstruction_order");
    c.limiter = IBIS_B200_LIMITER_BARTH_JESPERSEN;
    c.limiter_epsilon = 1e-25;
    c.thermo_interpolator = IBIS_B200_THERMO_RHO_U;
    if (c.reconstruction_order > 1) {
        const json& lim = cf.at("limiter");
        c.limiter = lookup(LIM, lim.at("type"), "Unknown limiter");
        c.limiter_epsilon = lim.value("epsilon", 1e-25);
        c.thermo_interpolator = lookup(THERMO, cf.at("thermo_interpolator"), "Unknown thermo interpolator");
    }
    const json& vf = config.at("viscous_flux");
    c.viscous = vf.at("enabled").get<bool>() ? 1 : 0;
    c.viscous_signal_factor = vf.at("signal_factor");
    const json& gm = config.at("gas_model");
    c.R = gm.at("R"); c.Cv = gm.at("Cv"); c.Cp = gm.at("Cp"); c.gamma = gm.at("gamma");
    const json& tp = config.at("transport_properties");
    if (tp.at("viscosity").at("type") != "sutherland") throw std::runtime_error("Unkown viscosity model");
    if (tp.at("thermal_conductivity").at("type") != "constant_prandtl_number")
        throw std::runtime_error("Unkown thermal conductivity model");
    c.mu0 = tp.at("viscosity").at("mu_0"); c.T0 = tp.at("viscosity").at("T_0"); c.Ts = tp.at("viscosity").at("T_s");
    c.Pr = tp.at("thermal_conductivity").at("Pr");
    const json& bnds = config.at("grid").at("boundaries");
    static const char* PHASES[3] = {"pre_reconstruction", "post_convective_flux", "pre_viscous_grad"};
    static const int PHASE_ID[3] = {IBIS_B200_PHASE_PRE_RECONSTRUCTION, IBIS_B200_PHASE_POST_CONVECTIVE_FLUX,
                                    IBIS_B200_PHASE_PRE_VISCOUS_GRAD};
    for (size_t t = 0; t < tags.size(); t++) {
        const json& b = bnds.at(tags[t]);
        for (int ph = 0; ph < 3; ph++) {
            if (!b.contains(PHASES[ph])) continue;
            for (const json& act : b.at(PHASES[ph])) {
                ibis_b200_bc_action a{};
                a.tag = (int32_t)t;
                a.phase = PHASE_ID[ph];
                const std::string type = act.at("type");
                if (ph == 1) {
                    if (type != "constant_flux") throw std::runtime_error("Unknown flux action");
                    a.type = IBIS_B200_BC_CONSTANT_FLUX;
                } else {
                    a.type = lookup(BC, type, "Unknown boundary action");
                }
                if (type == "flow_state_copy" || type == "subsonic_inflow" || type == "constant_flux") {
                    flow_state_params(act.at("flow_state"), a.params);
                } else if (type == "fix_temperature") {
                    a.params[0] = act.at("temperature");
                } else if (type == "subsonic_outflow") {
                    a.params[0] = act.at("pressure");
                } else if (type == "boundary_layer_profile") {          // boundary.cpp:51-58
                    const json& prof = act.at("profile");
                    const size_t n = prof.at("height").size();
                    if (prof.at("v").size() != n || prof.at("T").size() != n)
                        throw std::runtime_error("Cubic spline has different number of x and y points");
                    a.params[0] = prof.at("p");
                    a.params[1] = (double)n;
                    a.params[2] = (double)out->tables.size();
                    for (const char* k : {"height", "v", "T"})
                        for (const json& v : prof.at(k)) out->tables.push_back(v.get<double>());
                }
                out->actions.push_back(a);
            }
        }
    }
    c.n_bc_actions = (int32_t)out->actions.size();
    c.bc_actions = out->actions.data();
    c.n_bc_tables = (int64_t)out->tables.size();
    c.bc_tables = out->tables.data();
    return out;
}

// one B200 context per grid (the reference builds one GridBlock per number type: Ibis::real for the explicit
// solver, Ibis::dual for the steady one; finite_volume.h has no room for extra members, so the context hangs off
// the grid's address)
struct Backend {
    ibis_b200_ctx* ctx = nullptr;
    int fs = 0, dudt = 0;
    int64_t n_cells = 0, n_total = 0;
    int n_cons = 0, is_dual = 0;
    std::vector<double> prim_re, prim_du, cons_re, cons_du;
    ~Backend() {
        if (ctx) ibis_b200_destroy(ctx);
    }
};
std::map<const void*, std::unique_ptr<Backend>>& backends() {
    static std::map<const void*, std::unique_ptr<Backend>> m;
    return m;
}
// config of the FiniteVolume under construction / last constructed, per number type (simulation.h builds one each)
template <typename T>
json& stored_config() {
    static json cfg;
    return cfg;
}

template <typename T>
void copy_vec3(const Vector3s<T>& v, size_t n, std::vector<double>& out) {
    out.resize(3 * n);
    for (size_t i = 0; i < n; i++) {
        out[3 * i + 0] = re(v.x(i));
        out[3 * i + 1] = re(v.y(i));
        out[3 * i + 2] = re(v.z(i));
    }
}

// GridBlock<T> -> ibis_b200_grid_desc -> ibis_b200_create (grid/grid.h:609-625, interface.h:329-344, cell.h:112-114,325-333)
template <typename T>
Backend& backend_for(GridBlock<T>& grid) {
    auto& slot = backends()[static_cast<const void*>(&grid)];
    if (slot) return *slot;
    auto b = std::make_unique<Backend>();
    const size_t nf = grid.num_interfaces(), nc = grid.num_cells(), nt = grid.num_total_cells();
    auto f = grid.interfaces();
    auto cells = grid.cells();
    auto cf = cells.faces();
    std::vector<int64_t> left(nf), right(nf), off(nc + 1, 0), ids, nbr, tag_off(1, 0), tag_faces;
    std::vector<int32_t> outsign;
    std::vector<double> area(nf), norm, t1, t2, centre, vol(nc), centroid;
    for (size_t i = 0; i < nf; i++) {
        left[i] = (int64_t)f.left_cell(i);
        right[i] = (int64_t)f.right_cell(i);
        area[i] = re(f.area(i));
    }
    copy_vec3(f.norm(), nf, norm);
    copy_vec3(f.tan1(), nf, t1);
    copy_vec3(f.tan2(), nf, t2);
    copy_vec3(f.centre(), nf, centre);
    for (size_t i = 0; i < nc; i++) {
        auto fid = cf.face_ids(i);
        auto os = cf.outsigns(i);
        for (size_t j = 0; j < fid.size(); j++) {
            ids.push_back((int64_t)fid(j));
            outsign.push_back(os(j));
            nbr.push_back((int64_t)cells.neighbour_cells(i, j));
        }
        off[i + 1] = (int64_t)ids.size();
        vol[i] = re(cells.volume(i));
    }
    copy_vec3(cells.centroids(), nt, centroid);
    const std::vector<std::string> tags = grid.boundary_tags();       // alphabetical: the order finite_volume.cpp:187-218 applies
    for (const std::string& tag : tags) {
        auto bf = grid.boundary_faces(tag);
        for (size_t k = 0; k < bf.size(); k++) tag_faces.push_back((int64_t)bf(k));
        tag_off.push_back((int64_t)tag_faces.size());
    }
    ibis_b200_grid_desc d{};
    d.dim = (int32_t)grid.dim();
    d.n_tags = (int32_t)tags.size();
    d.n_cells = (int64_t)nc;
    d.n_ghost = (int64_t)grid.num_ghost_cells();
    d.n_faces = (int64_t)nf;
    d.face_left = left.data(); d.face_right = right.data(); d.face_area = area.data();
    d.face_norm = norm.data(); d.face_tan1 = t1.data(); d.face_tan2 = t2.data(); d.face_centre = centre.data();
    d.cell_face_offsets = off.data(); d.cell_face_ids = ids.data(); d.cell_outsigns = outsign.data();
    d.cell_neighbours = nbr.data(); d.cell_volume = vol.data(); d.cell_centroid = centroid.data();
    d.tag_offsets = tag_off.data(); d.tag_faces = tag_faces.data();
    d.n_owned = (int64_t)nc;
    auto flat = flatten(stored_config<T>(), tags);
    int rc = ibis_b200_create(&b->ctx, &d, &flat->c, /*device=*/0);
    if (rc != 0) throw std::runtime_error(std::string("ibis_b200_create: ") + ibis_b200_last_error(nullptr));
    b->n_cells = (int64_t)nc;
    b->n_total = (int64_t)nt;
    b->n_cons = (int)grid.dim() + 2;
    b->is_dual = is_dual_v<T>;
    ck(b->ctx, ibis_b200_fs_create(b->ctx, is_dual_v<T>, &b->fs), "fs_create");
    ck(b->ctx, ibis_b200_cq_create(b->ctx, is_dual_v<T>, &b->dudt), "cq_create");
    b->prim_re.resize(7 * nt);
    b->prim_du.resize(7 * nt);
    b->cons_re.resize((size_t)b->n_cons * nt);
    b->cons_du.resize((size_t)b->n_cons * nt);
    slot = std::move(b);
    fprintf(stderr, "[ibis_b200 adapter] FiniteVolume<%s> bound to libibis_b200.so: %zu cells, %zu interfaces\n",
            is_dual_v<T> ? "dual" : "real", nc, nf);
    return *slot;
}

// FlowStates<T> -> FlowStates rows rho p T e vx vy vz (include/ibis_b200.h conventions) -> device
template <typename T>
void upload_flow(Backend& b, const FlowStates<T>& fs) {
    // a caller may hold fewer rows than the grid has cells + ghosts (count_bad_cells looks at valid cells only)
    const int64_t n = std::min<int64_t>(b.n_total, (int64_t)fs.number_flow_states());
    for (int64_t i = n; i < b.n_total; i++)
        for (int k = 0; k < 7; k++) b.prim_re[7 * i + k] = b.prim_du[7 * i + k] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const T v[7] = {fs.gas.rho(i), fs.gas.pressure(i), fs.gas.temp(i), fs.gas.energy(i), fs.vel.x(i), fs.vel.y(i), fs.vel.z(i)};
        for (int k = 0; k < 7; k++) {
            b.prim_re[7 * i + k] = re(v[k]);
            b.prim_du[7 * i + k] = du(v[k]);
        }
    }
    ck(b.ctx, ibis_b200_fs_upload(b.ctx, b.fs, b.prim_re.data(), is_dual_v<T> ? b.prim_du.data() : nullptr), "fs_upload");
}

}  // namespace

template <typename T>
FiniteVolume<T>::FiniteVolume(GridBlock<T>& grid, json config) {
    dim_ = grid.dim();
    stored_config<T>() = config;
    // the B200 context is created from the grid at the first evaluation (backend_for): the solvers copy-assign
    // FiniteVolume objects, the grid they pass to every call is the stable key
}

template <typename T>
size_t FiniteVolume<T>::compute_dudt(FlowStates<T>& flow_state, GridBlock<T>& grid, ConservedQuantities<T>& dudt,
                                     IdealGas<T>& gas_model, TransportProperties<T>& trans_prop, bool allow_reconstruction) {
    (void)gas_model;
    (void)trans_prop;
    Backend& b = backend_for(grid);
    static const bool trace = getenv("IBIS_B200_ADAPTER_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[ibis_b200 adapter] compute_dudt<%s>(allow_reconstruction = %d)\n", is_dual_v<T> ? "dual" : "real", (int)allow_reconstruction);
    upload_flow(b, flow_state);
    ck(b.ctx, ibis_b200_compute_dudt(b.ctx, b.fs, b.dudt, allow_reconstruction ? 1 : 0), "compute_dudt");
    // the boundary conditions have rewritten the ghost cells (finite_volume.cpp:62, 187-199): they go back into the
    // caller's flow state as the reference leaves them
    ck(b.ctx, ibis_b200_fs_download(b.ctx, b.fs, b.prim_re.data(), is_dual_v<T> ? b.prim_du.data() : nullptr), "fs_download");
    for (int64_t i = b.n_cells; i < b.n_total; i++) {
        const double* r = &b.prim_re[7 * i];
        const double* d = &b.prim_du[7 * i];
        set(flow_state.gas.rho(i), r[0], d[0]);
        set(flow_state.gas.pressure(i), r[1], d[1]);
        set(flow_state.gas.temp(i), r[2], d[2]);
        set(flow_state.gas.energy(i), r[3], d[3]);
        set(flow_state.vel.x(i), r[4], d[4]);
        set(flow_state.vel.y(i), r[5], d[5]);
        set(flow_state.vel.z(i), r[6], d[6]);
    }
    ck(b.ctx, ibis_b200_cq_download(b.ctx, b.dudt, b.cons_re.data(), is_dual_v<T> ? b.cons_du.data() : nullptr), "cq_download");
    const int nc = b.n_cons;
    for (int64_t i = 0; i < b.n_cells; i++)
        for (int k = 0; k < nc; k++) set(dudt(i, k), b.cons_re[(size_t)i * nc + k], b.cons_du[(size_t)i * nc + k]);
    return 0;
}

template <typename T>
size_t FiniteVolume<T>::compute_dudt(FlowStates<T>& flow_state, Vector3s<T> vertex_vel, const ConservedQuantities<T>& cq,
                                     GridBlock<T>& grid, ConservedQuantities<T>& dudt, IdealGas<T>& gas_model,
                                     TransportProperties<T>& trans_prop, bool allow_reconstruction) {
    (void)vertex_vel;
    (void)cq;
    if (grid.moving())
        throw std::runtime_error("ibis_b200 adapter: moving grids go through ibis_b200_compute_dudt_moving (INTEGRATION.md A.5)");
    return compute_dudt(flow_state, grid, dudt, gas_model, trans_prop, allow_reconstruction);
}

template <typename T>
Ibis::real FiniteVolume<T>::estimate_dt(const FlowStates<T>& flow_state, GridBlock<T>& grid, IdealGas<T>& gas_model,
                                        TransportProperties<T>& trans_prop) {
    (void)gas_model;
    (void)trans_prop;
    Backend& b = backend_for(grid);
    upload_flow(b, flow_state);
    double dt = 0.0;
    ck(b.ctx, ibis_b200_estimate_dt(b.ctx, b.fs, &dt), "estimate_dt");
    return dt;
}

template <typename T>
size_t FiniteVolume<T>::count_bad_cells(const FlowStates<T>& fs, const size_t num_cells) {
    (void)num_cells;
    // the solvers call this right after a step on the grid they evaluate on: there is one context per number type
    for (auto& kv : backends()) {
        Backend& b = *kv.second;
        if ((size_t)b.n_cells != num_cells || b.is_dual != is_dual_v<T>) continue;
        upload_flow(b, fs);
        uint64_t n_bad = 0;
        ck(b.ctx, ibis_b200_count_bad_cells(b.ctx, b.fs, &n_bad), "count_bad_cells");
        return (size_t)n_bad;
    }
    throw std::runtime_error("ibis_b200 adapter: count_bad_cells before the first residual evaluation");
}

// ---- the rest of the class is the inside of compute_dudt in the reference (finite_volume.cpp:187-330) or serves the
// gradient post-processing of io/accessor.cpp: not provided by this backend ----
#define IBIS_B200_NOT_PROVIDED(name) throw std::runtime_error("ibis_b200 adapter: FiniteVolume::" name " is internal to the B200 residual")
template <typename T>
void FiniteVolume<T>::apply_pre_reconstruction_bc(FlowStates<T>&, const GridBlock<T>&, const IdealGas<T>&, const TransportProperties<T>&) {
    IBIS_B200_NOT_PROVIDED("apply_pre_reconstruction_bc");
}
template <typename T>
void FiniteVolume<T>::apply_post_convective_flux_bc(const FlowStates<T>&, const GridBlock<T>&, const IdealGas<T>&, const TransportProperties<T>&) {
    IBIS_B200_NOT_PROVIDED("apply_post_convective_flux_bc");
}
template <typename T>
void FiniteVolume<T>::apply_pre_viscous_grad_bc(FlowStates<T>&, const GridBlock<T>&, const IdealGas<T>&, const TransportProperties<T>&) {
    IBIS_B200_NOT_PROVIDED("apply_pre_viscous_grad_bc");
}
template <typename T>
void FiniteVolume<T>::apply_geometric_conservation_law(const ConservedQuantities<T>&, const GridBlock<T>&, ConservedQuantities<T>&) {
    IBIS_B200_NOT_PROVIDED("apply_geometric_conservation_law");
}
template <typename T>
void FiniteVolume<T>::flux_surface_integral(const GridBlock<T>&, ConservedQuantities<T>&) {
    IBIS_B200_NOT_PROVIDED("flux_surface_integral");
}
template <typename T>
void FiniteVolume<T>::compute_viscous_gradient(FlowStates<T>&, const GridBlock<T>&, const IdealGas<T>&, const TransportProperties<T>&) {
    IBIS_B200_NOT_PROVIDED("compute_viscous_gradient");
}
template <typename T>
void FiniteVolume<T>::compute_convective_gradient(FlowStates<T>&, const GridBlock<T>&, const IdealGas<T>&, const TransportProperties<T>&) {
    IBIS_B200_NOT_PROVIDED("compute_convective_gradient");
}

template class FiniteVolume<Ibis::real>;
template class FiniteVolume<Ibis::dual>;
