/* ibis_b200.h -- C ABI of libibis_b200.so: the fp64 finite-volume residual hot path of
 * wattrg/ibis as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI seam: its solvers call C++ templates on Kokkos Views.  This
 * header is the seam a maintainer would bind: every entry point
 * names the reference interface it replaces (paths relative to /root/reference/src).
 *
 * Conventions
 *   - extern "C", POD arguments only; every call returns 0 or a negative IBIS_B200_E*
 *     code and never throws; ibis_b200_last_error() gives the message.
 *   - the context owns ALL device memory; the host passes/receives plain host arrays in
 *     the REFERENCE numbering and layout (cells: valid then ghost, ghost ids in
 *     alphabetical-tag order; FlowStates rows = rho,p,T,e,vx,vy,vz; conserved rows =
 *     mass, mom_x, mom_y[, mom_z], energy).  Any renumbering is internal.
 *   - buffers are opaque integer handles (> 0) owned by the context.
 *   - T = Ibis::real  <->  is_dual = 0 ;  T = Ibis::dual (util/dual.h) <-> is_dual = 1,
 *     real and dual parts travel as two separate host arrays.
 *   - one context per process and GPU; calls are asynchronous on the context's stream
 *     unless they return a value to the host.
 */
#ifndef IBIS_B200_H
#define IBIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ibis_b200_ctx ibis_b200_ctx;

enum {
    IBIS_B200_OK = 0,
    IBIS_B200_EINVAL = -1,   /* bad argument / unsupported configuration */
    IBIS_B200_ECUDA = -2,    /* CUDA runtime error                        */
    IBIS_B200_ENOMEM = -3,
    IBIS_B200_EHANDLE = -4,  /* unknown or wrong-kind buffer handle       */
    IBIS_B200_ECOMM = -5     /* halo exchange / collective failure        */
};

/* finite_volume/flux_calc.cpp:9-23 (make_flux_calculator) */
enum { IBIS_B200_FLUX_HANEL = 0, IBIS_B200_FLUX_AUSMDV = 1, IBIS_B200_FLUX_LDFSS2 = 2,
       IBIS_B200_FLUX_RUSANOV = 3 };
/* finite_volume/limiter.cpp:8-19 (make_limiter) */
enum { IBIS_B200_LIMITER_UNLIMITED = 0, IBIS_B200_LIMITER_BARTH_JESPERSEN = 1 };
/* finite_volume/convective_flux.h:15-20 (ThermoReconstructionVars); p_T is broken
 * upstream (convective_flux.cpp:53-55) and rejected with EINVAL */
enum { IBIS_B200_THERMO_RHO_P = 0, IBIS_B200_THERMO_RHO_T = 1, IBIS_B200_THERMO_RHO_U = 2,
       IBIS_B200_THERMO_P_T = 3 };
/* finite_volume/boundaries/boundary.cpp:449-479 (build_boundary_action) and :482-492
 * (build_flux_action: constant_flux, the one FluxAction) */
enum { IBIS_B200_BC_FLOW_STATE_COPY = 0, IBIS_B200_BC_INTERNAL_COPY = 1,
       IBIS_B200_BC_INTERNAL_COPY_REFLECT_NORMAL = 2, IBIS_B200_BC_INTERNAL_VEL_COPY_REFLECT = 3,
       IBIS_B200_BC_FIX_TEMPERATURE = 4, IBIS_B200_BC_SUBSONIC_INFLOW = 5,
       IBIS_B200_BC_SUBSONIC_OUTFLOW = 6, IBIS_B200_BC_BOUNDARY_LAYER_PROFILE = 7,
       IBIS_B200_BC_CONSTANT_FLUX = 8 };
/* which list of BoundaryCondition the action belongs to (boundary.cpp:495-515) */
enum { IBIS_B200_PHASE_PRE_RECONSTRUCTION = 0, IBIS_B200_PHASE_PRE_VISCOUS_GRAD = 1,
       IBIS_B200_PHASE_POST_CONVECTIVE_FLUX = 2 };

/* One GhostCellAction / FluxAction of one boundary tag, in the order the reference applies them
 * (tags alphabetical: finite_volume.cpp:187-218; actions in list order: boundary.cpp:517-543).
 * params: flow_state_copy / subsonic_inflow / constant_flux = {rho,p,T,e,vx,vy,vz};
 *         fix_temperature = {Twall}; subsonic_outflow = {pressure};
 *         boundary_layer_profile = {p, n_pts, offset}: the profile's "height", "v" and "T" lists
 *         (boundary.cpp:51-58), n_pts doubles each, back to back at config->bc_tables + offset.
 * constant_flux is only valid in phase POST_CONVECTIVE_FLUX and the ghost-cell actions only in the
 * other two, like the reference's two factories. */
typedef struct {
    int32_t tag;
    int32_t phase;
    int32_t type;
    int32_t reserved;
    double params[8];
} ibis_b200_bc_action;

/* The reference GridBlock<T> (grid/grid.h:609-625) as plain arrays, reference numbering.
 * Pointers are borrowed during ibis_b200_create only. Vector arrays are n x 3, row-major. */
typedef struct {
    int32_t dim;                      /* grid.h:617 */
    int32_t n_tags;
    int64_t n_cells;                  /* num_valid_cells_ (grid.h:618) */
    int64_t n_ghost;                  /* num_ghost_cells_ (grid.h:619) */
    int64_t n_faces;
    const int64_t* face_left;         /* Interfaces::left_cells_  (grid/interface.h:333) */
    const int64_t* face_right;        /* Interfaces::right_cells_ (interface.h:334)      */
    const double* face_area;          /* interface.h:337 */
    const double* face_norm;          /* interface.h:339 */
    const double* face_tan1;          /* interface.h:340 */
    const double* face_tan2;          /* interface.h:341 */
    const double* face_centre;        /* interface.h:342 */
    const int64_t* cell_face_offsets; /* CellFaces::offsets_  (grid/cell.h:112), n_cells+1 */
    const int64_t* cell_face_ids;     /* CellFaces::face_ids_ (cell.h:113)  */
    const int32_t* cell_outsigns;     /* CellFaces::outsigns_ (cell.h:114)  */
    const int64_t* cell_neighbours;   /* Cells::neighbour_cells_ (cell.h:327), same offsets */
    const double* cell_volume;        /* Cells::volume_   (cell.h:329), n_cells  */
    const double* cell_centroid;      /* Cells::centroid_ (cell.h:330), n_cells + n_ghost */
    const int64_t* tag_offsets;       /* n_tags + 1, into tag_faces; tags alphabetical */
    const int64_t* tag_faces;         /* GridBlock::boundary_faces_ (grid.h:624), concatenated */
    /* domain decomposition (no reference counterpart: the reference is single block).
     * Cells [0, n_owned) are owned by this rank, cells [n_owned, n_cells) are halo copies
     * of cells owned by other ranks; residuals are produced for owned cells only.
     * n_owned == n_cells on one GPU. */
    int64_t n_owned;
} ibis_b200_grid_desc;

/* config/config.json flattened. */
typedef struct {
    int32_t flux_calculator;          /* convective_flux.flux_calculator.type */
    int32_t reconstruction_order;     /* convective_flux.reconstruction_order (1|2) */
    int32_t limiter;                  /* convective_flux.limiter.type */
    int32_t thermo_interpolator;      /* convective_flux.thermo_interpolator */
    double limiter_epsilon;           /* barth_jespersen epsilon (limiter.cpp:11) */
    double ldfss_delta;               /* ldfss2 delta (flux_calc.h:62) */
    int32_t viscous;                  /* viscous_flux.enabled */
    int32_t n_bc_actions;
    double viscous_signal_factor;     /* viscous_flux.signal_factor */
    double R, Cv, Cp, gamma;          /* gas_model (gas/gas_model.h:58-62) */
    double mu0, T0, Ts, Pr;           /* transport_properties (gas/transport_properties.h:22-26,57-60) */
    const ibis_b200_bc_action* bc_actions;
    int64_t n_bc_tables;              /* doubles at bc_tables (0 without boundary_layer_profile) */
    const double* bc_tables;          /* borrowed during ibis_b200_create only */
} ibis_b200_config;

/* ---- grid build ---------------------------------------------
 * replaces: GridIO::read_su2_grid's element tables + InterfaceLookup (string-hashed face
 * de-duplication, grid/interface.cpp:38-80) + GridBlock<T>::init_grid_block and the geometry
 * kernels it calls (grid/grid.h:89-167,186-213,364-433; interface.h:187-250; cell.h:236-303;
 * geom.h:30-83).  Input: an SU2 mesh as plain arrays.  Output: the arrays of the reference
 * GridBlock in the reference's numbering (faces in order of first discovery, ghosts appended per
 * tag in alphabetical order) -- an ibis_b200_grid_desc that ibis_b200_create takes as is.
 * Connectivity is one lexicographic sort on the host; every geometric quantity is computed by
 * CUDA kernels on `device` with the reference's formulas (bit-identical arrays). */
typedef struct {
    int32_t dim;
    int32_t n_tags;                   /* boundary markers, alphabetical by tag                   */
    int64_t n_vertices;
    int64_t n_cells;
    const double* vertices;           /* n_vertices x 3 (z = 0 in 2-D)                           */
    const int32_t* cell_types;        /* VTK code: 5 tri, 9 quad, 10 tetra, 12 hex, 13 wedge, 14 pyramid */
    const int64_t* cell_vertices;     /* n_cells x 8, padded with -1, VTK vertex order           */
    const int64_t* marker_offsets;    /* n_tags + 1, into the marker elements                    */
    const int64_t* marker_vertices;   /* n_marker_elements x 4, padded with -1                   */
} ibis_b200_mesh_desc;
/* replaces GridIO::GridIO(file_name) + GridIO::read_su2_grid (grid/grid_io.cpp:100-268): an SU2 text
 * file as a mesh description; tags alphabetical (GridIO keeps its markers in a std::map), a tag that
 * appears twice has its elements appended, unknown lines ignored.  Host only. */
typedef struct ibis_b200_mesh ibis_b200_mesh;
int ibis_b200_mesh_read_su2(ibis_b200_mesh** mesh, const char* path);
const ibis_b200_mesh_desc* ibis_b200_mesh_desc_of(const ibis_b200_mesh* mesh);
const char* ibis_b200_mesh_tag(const ibis_b200_mesh* mesh, int i);        /* NULL past the last tag */
int ibis_b200_mesh_destroy(ibis_b200_mesh* mesh);

typedef struct ibis_b200_grid ibis_b200_grid;
int ibis_b200_grid_build(ibis_b200_grid** grid, const ibis_b200_mesh_desc* mesh, int device);
/* the built arrays; valid until ibis_b200_grid_destroy */
const ibis_b200_grid_desc* ibis_b200_grid_desc_of(const ibis_b200_grid* grid);
/* Interfaces::vertex_ids (grid/interface.h:331): n_faces x 4 padded with -1, orientation of the
 * first discoverer; and the ghost cell of every boundary face, in tag_faces order */
const int64_t* ibis_b200_grid_face_vertices(const ibis_b200_grid* grid);
const int64_t* ibis_b200_grid_tag_ghosts(const ibis_b200_grid* grid);
int ibis_b200_grid_destroy(ibis_b200_grid* grid);

/* ---- context ------------------------------------------------------------------------- */
/* replaces: GridBlock<T>(su2, config) + FiniteVolume<T>(grid, config) + IdealGas/
 * TransportProperties constructors (solvers/runge_kutta.cpp:21-80, simulation/simulation.h).
 * device < 0 selects the current CUDA device. */
int ibis_b200_create(ibis_b200_ctx** ctx, const ibis_b200_grid_desc* grid,
                     const ibis_b200_config* config, int device);
int ibis_b200_destroy(ibis_b200_ctx* ctx);
const char* ibis_b200_last_error(const ibis_b200_ctx* ctx);   /* ctx may be NULL */
int ibis_b200_sync(ibis_b200_ctx* ctx);
/* GridBlock::dim / num_cells / num_ghost_cells / num_interfaces (grid/grid.h:617-621) */
int ibis_b200_sizes(const ibis_b200_ctx* ctx, int32_t* dim, int64_t* n_cells, int64_t* n_ghost,
                    int64_t* n_faces);
/* number of kernels launched by this context since creation (bench.py: gpu_launches) */
int64_t ibis_b200_launch_count(const ibis_b200_ctx* ctx);
/* the CUDA stream all kernels are launched on, as a cudaStream_t cast to void* */
void* ibis_b200_stream(ibis_b200_ctx* ctx);

/* Per-kernel device timing for bench.py's roofline: when enabled, pass G (gradients + limiters)
 * and pass F (reconstruction + flux + surface integral) launches and the halo exchange are
 * bracketed by CUDA events on the context's stream.  profile_read synchronises and returns the
 * accumulated milliseconds and counts since the last read: [0] = pass G, [1] = pass F,
 * [2] = halo exchange (pack, ncclSend/ncclRecv, unpack), [3] = reserved. */
int ibis_b200_profile_enable(ibis_b200_ctx* ctx, int on);
int ibis_b200_profile_read(ibis_b200_ctx* ctx, double ms[4], int64_t count[4]);

/* ---- buffers ------------------------------------------------------------------------- */
/* FlowStates<T>(num_total_cells) (gas/flow_state.h:62) */
int ibis_b200_fs_create(ibis_b200_ctx*, int is_dual, int* handle);
/* ConservedQuantities<T>(num_total_cells, dim) (finite_volume/conserved_quantities.cpp:16-23) */
int ibis_b200_cq_create(ibis_b200_ctx*, int is_dual, int* handle);
/* Ibis::Vector<real>(n) (linear_algebra/dense_linear_algebra.h:17-90) */
int ibis_b200_vec_create(ibis_b200_ctx*, int64_t n, int* handle);
int ibis_b200_free(ibis_b200_ctx*, int handle);
/* host <-> device in reference numbering; the *_du pointers may be NULL (dual part 0 /
 * not wanted).  replaces FVIO::read/write host mirrors (io/io.cpp:107-160). */
int ibis_b200_fs_upload(ibis_b200_ctx*, int fs, const double* prim_re, const double* prim_du);
int ibis_b200_fs_download(ibis_b200_ctx*, int fs, double* prim_re, double* prim_du);
int ibis_b200_cq_upload(ibis_b200_ctx*, int cq, const double* cons_re, const double* cons_du);
int ibis_b200_cq_download(ibis_b200_ctx*, int cq, double* cons_re, double* cons_du);
int ibis_b200_vec_upload(ibis_b200_ctx*, int vec, const double* x);
int ibis_b200_vec_download(ibis_b200_ctx*, int vec, double* x);
/* The reference's native_binary flow snapshot (io/native.cpp:200-339): one raw f64 array per variable,
 * T, p, vx, vy[, vz], over the num_valid_cells in reference numbering (no ghost rows).
 * flow_read_native = NativeBinaryInput::read followed by its update_thermo_from_pT loop
 * (native.cpp:280-336: rho = p / (R T), e = Cv T; ghost rows untouched); flow_write_native =
 * NativeBinaryOutput::write (native.cpp:200-262).  40 bytes per cell cross the bus instead of the 56
 * bytes per row (ghosts included) of fs_upload / fs_download.  vz may be NULL in 2-D. */
int ibis_b200_flow_read_native(ibis_b200_ctx*, int fs, const double* T, const double* p, const double* vx,
                               const double* vy, const double* vz);
int ibis_b200_flow_write_native(ibis_b200_ctx*, int fs, double* T, double* p, double* vx, double* vy, double* vz);
/* Snapshot ring (the native flow IO fast path).  The same native_binary arrays, moved
 * WHILE the device computes: two upload and two download slots of device staging, one copy stream per direction
 * beside the compute stream, events between them.  Host arrays must be page-locked (ibis_b200_pinned_alloc) and
 * stay untouched between a begin and its end.
 *   upload_begin(slot, arrays)        the H2D copy of a snapshot starts on the upload stream; returns at once
 *   upload_end(slot, fs)              the compute stream waits for it and unpacks into fs (NativeBinaryInput::read +
 *                                     update_thermo_from_pT, native.cpp:280-336); returns at once
 *   download_begin(slot, fs, arrays)  the compute stream packs fs into the slot, the download stream copies it out
 *                                     (NativeBinaryOutput::write's arrays, native.cpp:200-262); returns at once
 *   download_end(slot)                the host waits until that snapshot has landed in the host arrays
 * With two slots, snapshot k + 1 arrives and snapshot k - 1 leaves while step k runs (bench.py's e2e loop;
 * ibis_b200/native_io.py writes the reference's io/flow/NNNN files from the landed arrays). */
int ibis_b200_pinned_alloc(ibis_b200_ctx*, int64_t bytes, void** host_ptr);
int ibis_b200_pinned_free(ibis_b200_ctx*, void* host_ptr);
int ibis_b200_snapshot_upload_begin(ibis_b200_ctx*, int slot, const double* T, const double* p, const double* vx,
                                    const double* vy, const double* vz);
int ibis_b200_snapshot_upload_end(ibis_b200_ctx*, int slot, int fs);
int ibis_b200_snapshot_download_begin(ibis_b200_ctx*, int slot, int fs, double* T, double* p, double* vx, double* vy,
                                      double* vz);
int ibis_b200_snapshot_download_end(ibis_b200_ctx*, int slot);
int ibis_b200_copy(ibis_b200_ctx*, int dst, int src);          /* deep_copy, same kind */
int ibis_b200_zero(ibis_b200_ctx*, int buf);

/* ---- the residual ------------------------------------------------------------------ */
/* FiniteVolume<T>::compute_dudt (finite_volume/finite_volume.cpp:57-96): mutates the
 * ghost rows of fs, writes rows [0, n_cells) of dudt.  fs and dudt must both be real or
 * both dual. */
int ibis_b200_compute_dudt(ibis_b200_ctx*, int fs, int dudt, int allow_reconstruction);
/* FiniteVolume<T>::estimate_dt (finite_volume.cpp:131-184) */
int ibis_b200_estimate_dt(ibis_b200_ctx*, int fs, double* dt);
/* FiniteVolume<T>::count_bad_cells (finite_volume.cpp:258-270) */
int ibis_b200_count_bad_cells(ibis_b200_ctx*, int fs, uint64_t* n_bad);
/* conserved_to_primatives / primatives_to_conserved
 * (finite_volume/primative_conserved_conversion.cpp:5-57) */
int ibis_b200_cons_to_prim(ibis_b200_ctx*, int cq, int fs);
int ibis_b200_prim_to_cons(ibis_b200_ctx*, int cq, int fs);
/* apply_time_derivative(U0, U1, dUdt, dt) (conserved_quantities.cpp:80-94); U1 == U0
 * gives ConservedQuantities::apply_time_derivative (conserved_quantities.cpp:26-38) */
int ibis_b200_apply_time_derivative(ibis_b200_ctx*, int u0, int u1, int dudt, double dt);
/* ConservedQuantities<T>::L2_norms (conserved_quantities.cpp:41-69):
 * out = {global, mass, momentum_x, momentum_y, momentum_z, energy} */
int ibis_b200_l2_norms(ibis_b200_ctx*, int cq, double out[6]);
/* ---- moving grid: the residual terms (first part) -----------------
 * compute_dudt(flow_state, vertex_vel, cq, grid, dudt, ...) of a grid with motion enabled
 * (finite_volume.cpp:57-85): interface velocities = average of the vertex velocities
 * (GridBlock::compute_face_vel, grid/grid.h:571-592), fluxes relative to the moving interface with the
 * correction back to the inertial frame (convective_flux.cpp:105-140,166-176), constant_flux relative
 * to the interface (boundary.cpp:420-426), and the geometric-conservation-law term -U dV/dt / V
 * (finite_volume.cpp:99-128).  The vertex velocities are what the reference's GridMotionDriver
 * produces (rigid_body_translation.cpp:14-27: one constant vector; shock fitting is not built); moving
 * the vertices and recomputing the geometry between stages: ibis_b200_move_vertices below.
 * face_vertices: Interfaces::vertex_ids (grid/interface.h:331) = ibis_b200_grid_face_vertices, n_faces x 4
 * padded with -1, set once; vertex_vel: n_vertices x 3, copied at the call. */
int ibis_b200_set_face_vertices(ibis_b200_ctx*, const int64_t* face_vertices, int64_t n_vertices);
int ibis_b200_set_vertex_velocities(ibis_b200_ctx*, const double* vertex_vel);
int ibis_b200_compute_dudt_moving(ibis_b200_ctx*, int fs, int cq, int dudt, int allow_reconstruction);
/* Moving the vertices (solvers/runge_kutta.cpp:136,157-195): cell_types / cell_vertices as in
 * ibis_b200_mesh_desc, set once after the face vertices; set_vertex_positions hands the positions over (the
 * first call: the positions the grid was built from; later calls recompute the geometry); move_vertices does
 * positions += scale * vertex_vel on the device (add_scaled_vector, util/vector3.cpp:39-47) and then, when
 * recompute_geometry is set (several moves may precede one recomputation, runge_kutta.cpp:157-180), what
 * GridBlock::compute_geometric_data does (grid/grid.h:169-178): interface centres and areas, cell centroids and
 * volumes, ghost-cell centres, WLS weights -- not the interface orientations, as in the reference.
 * Once the vertices have moved the context serves ibis_b200_compute_dudt_moving only (the fused static
 * kernels keep tables of the original geometry). */
int ibis_b200_set_cell_vertices(ibis_b200_ctx*, const int32_t* cell_types, const int64_t* cell_vertices);
int ibis_b200_set_vertex_positions(ibis_b200_ctx*, const double* positions);
int ibis_b200_move_vertices(ibis_b200_ctx*, double scale, int recompute_geometry);
int ibis_b200_get_vertex_positions(ibis_b200_ctx*, double* positions);
/* GridBlock::face_vel() after the last moving evaluation (n_faces x 3, reference numbering) */
int ibis_b200_debug_face_velocities(ibis_b200_ctx*, double* face_vel);

/* intermediates of the last real evaluation, for bisecting parity (reference numbering):
 * grad [n_cells x 3] and limiter [n_cells] of variable var (0 rho, 1 thermo-2 (u|p|T),
 * 2 vx, 3 vy, 4 vz); either pointer may be NULL. */
int ibis_b200_debug_gradients(ibis_b200_ctx*, int var, double* grad, double* limiter);
/* the dual-number operators of the device code (Ibis::Dual, util/dual.h:75-314) on one pair of operands, so
 * that the reference's own Dual tests (util/dual.cpp:4-222) can be replayed on the device:
 * in = {x.re, x.du, y.re, y.du, r}; out = 22 (re, du) pairs: x+y, x+=y, x+r, x-y, x-=y, x-r, r-x, x*y, x*=y,
 * x*r, x*=r, r*x, x/y, x/=y, x/r, x/=r, r/x, sqrt(x), abs(x), abs(x*x), x*x+2x, x*x-sqrt(x) */
int ibis_b200_debug_dual_ops(ibis_b200_ctx*, const double* in5, double* out44);
/* Self-test of the paired / shared fp64 divisions the flux calculators and the limiter use on the device
 * (ibis_b200/csrc/scalar.h::div_pair, div_shared, div_z): n pseudo-random operand pairs per launch thread spanning
 * the whole exponent range, zeros, infinities and NaNs included, each quotient compared bit for bit with the plain
 * IEEE division.  *mismatches receives the number of differing quotients (0 expected). */
int ibis_b200_debug_division_check(ibis_b200_ctx*, int64_t n, uint64_t seed, uint64_t* mismatches);

/* ---- fused explicit step (solvers/runge_kutta.cpp:131-203) --------------------------- */
/* One RungeKutta::take_step with the tableau a (row i has i entries, i = 1..stages-1,
 * concatenated), b (stages): stage evaluations, dt estimate
 * (dt = min(cfl*dt_stable, 1.5*dt_prev, t_max - t), runge_kutta.cpp:108-119), stage updates
 * and final update, entirely on the device stream.  fs/cq are the solver's flow_ /
 * conserved_quantities_.  Returns dt and dt_stable.
 * Asynchronous form: with dt == NULL and dt_stable == NULL nothing is read back and the call returns as soon as
 * the step is enqueued; with dt_prev = NaN the "1.5 * dt_prev" of runge_kutta.cpp:113 takes the dt of the previous
 * ibis_b200_rk_step from the device (DBL_MAX before the first).  A solver that plots / prints every k steps runs k
 * such steps back to back and then calls ibis_b200_rk_last_dt (which synchronises). */
int ibis_b200_rk_step(ibis_b200_ctx*, int fs, int cq, int n_stages, const double* a,
                      const double* b, double cfl, double dt_prev, double t_remaining,
                      double* dt, double* dt_stable);
/* Reference quirk (runge_kutta.cpp:49-50): `k_` is a std::vector filled with copies of ONE
 * ConservedQuantities, i.e. of one reference-counted Kokkos View, so all Runge-Kutta stages k_[i]
 * alias the same array and every function evaluation overwrites "all" of them.  Schemes with up
 * to two stages are unaffected; for three or more the reference therefore does not integrate the
 * tableau it was given.  The default reproduces the reference bit for bit; distinct = 1 gives
 * each stage its own array (the scheme the tableau describes). */
int ibis_b200_rk_last_dt(ibis_b200_ctx*, double* dt, double* dt_stable);
int ibis_b200_rk_distinct_stages(ibis_b200_ctx*, int distinct);
/* RungeKutta::L2_norms = k_[0].L2_norms() of the last rk_step (runge_kutta.cpp:264) */
int ibis_b200_rk_residual_norms(ibis_b200_ctx*, double out[6]);

/* ---- JFNK pack / unpack (solvers/steady_state.cpp, solvers/jfnk.cpp) ------------------ */
/* set_dual: cq_tmp.re = cq.re, cq_tmp.du = vec[cell*n_cons + c] (steady_state.cpp:60-68) */
int ibis_b200_jfnk_set_dual(ibis_b200_ctx*, int cq, int vec, int cq_tmp);
/* set_vector: result = vec/dt_star - dual_part(residual) (steady_state.cpp:102-110) */
int ibis_b200_jfnk_set_vector(ibis_b200_ctx*, int residual, int vec, double dt_star, int result);
/* eval_rhs: rhs = real_part(residual) (steady_state.cpp:142-149) */
int ibis_b200_jfnk_eval_rhs(ibis_b200_ctx*, int residual, int rhs);
/* apply_update: cq.re += dU, cq.du = 0 (jfnk.cpp:69-76) */
int ibis_b200_jfnk_apply_update(ibis_b200_ctx*, int cq, int du);

/* ---- vectors (linear_algebra/dense_linear_algebra.h:185-309) ------------------------ */
int ibis_b200_vec_dot(ibis_b200_ctx*, int a, int b, double* out);          /* dot   :299 */
int ibis_b200_vec_norm2(ibis_b200_ctx*, int a, double* out);               /* norm2 :185 */
int ibis_b200_vec_axpy(ibis_b200_ctx*, int y, int x, double s);            /* y += s x :217 */
int ibis_b200_vec_scale(ibis_b200_ctx*, int x, int result, double s);      /* r = s x  :208 */
int ibis_b200_vec_sub(ibis_b200_ctx*, int a, int b, int result);           /* r = a - b (gmres.cpp:18) */
/* Krylov basis: a device matrix of n_cols contiguous columns of length n */
int ibis_b200_basis_create(ibis_b200_ctx*, int64_t n, int n_cols, int* handle);
int ibis_b200_basis_set_col(ibis_b200_ctx*, int basis, int col, int vec);  /* column(j).deep_copy_layout(v) */
/* one modified Gram-Schmidt sweep of gmres.cpp:144-150 on the device: for i <= j:
 * h[i] = dot(w, V_i); w -= h[i] V_i; then h[j+1] = norm2(w); v = w / h[j+1]; V_{j+1} = v.
 * h (host, j + 2 doubles) is returned after ONE synchronisation. */
int ibis_b200_arnoldi_mgs(ibis_b200_ctx*, int basis, int j, int w, int v, double* h);
/* gemv: out = V(:, 0:m) y (dense_linear_algebra.h:244) with y on the host */
int ibis_b200_basis_gemv(ibis_b200_ctx*, int basis, int m, const double* y, int out);

/* ---- domain decomposition (no reference counterpart) ----------- */
/* Per neighbour rank: the local ids (reference numbering of this rank's grid) of owned
 * cells to send and of halo cells to receive, in matching order on both sides. */
int ibis_b200_halo_setup(ibis_b200_ctx*, int n_peers, const int32_t* peer_ranks,
                         const int64_t* send_offsets, const int64_t* send_cells,
                         const int64_t* recv_offsets, const int64_t* recv_cells);
/* NCCL communicator from an ncclUniqueId (128 bytes) distributed by the host plumbing
 * (torch.distributed); used for the halo exchange inside compute_dudt / cons_to_prim and
 * for the MIN / SUM reductions of estimate_dt, l2_norms, dot, norm2, count_bad_cells. */
int ibis_b200_comm_init(ibis_b200_ctx*, const void* nccl_unique_id, int rank, int n_ranks);
int ibis_b200_nccl_unique_id(void* out_128_bytes);
/* Peer-memory halo exchange (one process per GPU, NVLink / NVSwitch): instead of ncclSend/ncclRecv the
 * boundary cells are stored by a kernel of the sending rank straight into the receiving rank's
 * memory, with device-side release / acquire flags -- no NCCL call and no host synchronisation on
 * the data path.  After halo_setup every rank exports a 64-byte CUDA IPC handle of its receive
 * allocation (and its total receive count); the host plumbing (torch.distributed / MPI) hands each
 * rank, per peer p of its halo_setup list: that peer's handle and receive count, the offset (in
 * cells) of this rank's segment in the peer's receive order, and this rank's index in the peer's
 * peer list.  halo_p2p_enable(ctx, 1) then switches the exchange over (0: back to NCCL). */
int ibis_b200_halo_p2p_export(ibis_b200_ctx*, void* out_handle_64_bytes, int64_t* n_recv_cells);
int ibis_b200_halo_p2p_import(ibis_b200_ctx*, int peer_index, const void* handle_64_bytes,
                              int64_t peer_n_recv_cells, int64_t my_offset_in_peer, int my_slot_in_peer);
int ibis_b200_halo_p2p_enable(ibis_b200_ctx*, int on);

#ifdef __cplusplus
}
#endif
#endif /* IBIS_B200_H */
