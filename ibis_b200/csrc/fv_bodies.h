// fv_bodies.h -- per-index bodies of the finite-volume kernels (one boundary face, one
// cell or one interface per call).  The __global__ wrappers in fv_kernels.cu map one
// thread to one index.  The functions are __host__ __device__ only so that the very same
// source can be exercised on the host by tests/emu (the development container has no
// GPU); the shipped library never runs them on the CPU.
#pragma once

#include <stdint.h>

#include "grid_bodies.h"
#include "physics.h"

namespace ib {

constexpr int MAXF = 6;       // faces per cell: hex 6, wedge/pyramid 5, tet/quad 4, tri 3
// The side states of the sides path: one record (rho p T vx vy vz) per interface side, written by pass G' every
// evaluation into per-tile arrays; the static frame of an interface is t1 t2 (6 doubles).  Bytes of one interface
// slot in shared memory: 2 side records + 6 doubles.
template <typename S>
struct frec {
    static constexpr int bytes = 12 * (int)sizeof(S) + 48;
};
constexpr int TILE_INFO = 64;  // ints per tile descriptor (GridDev::tile_info): 12 header + 10 inline entries, 256 B
constexpr int TILE_INLINE_ENT = 10;
constexpr int NVARS = 6;      // gradient slots: rho, thermo-2, vx, vy, vz, T(viscous)
enum { V_RHO = 0, V_TH2 = 1, V_VX = 2, V_VY = 3, V_VZ = 4, V_TVISC = 5 };
enum { F_RHO = 0, F_P = 1, F_T = 2, F_E = 3, F_VX = 4, F_VY = 5, F_VZ = 6 };
enum { BC_FLOW_STATE_COPY = 0, BC_INTERNAL_COPY = 1, BC_REFLECT_NORMAL = 2, BC_VEL_REFLECT = 3,
       BC_FIX_TEMPERATURE = 4, BC_SUBSONIC_INFLOW = 5, BC_SUBSONIC_OUTFLOW = 6,
       BC_BOUNDARY_LAYER_PROFILE = 7, BC_CONSTANT_FLUX = 8 };
enum { PHASE_PRE_RECONSTRUCTION = 0, PHASE_PRE_VISCOUS_GRAD = 1, PHASE_POST_CONVECTIVE_FLUX = 2, N_PHASES = 3 };

struct BcAction {
    int type;
    int pad;
    double params[7];
};

// Device-resident grid: int32 ids, SoA geometry.  Cell ids: [0, n_owned) owned,
// [n_owned, n_cells) halo copies, [n_cells, n_total) boundary ghosts.
struct GridDev {
    int dim, n_owned, n_grad, n_cells, n_total, n_faces, n_bfaces, max_faces;
    long cstride;               // padded n_total: stride of every per-cell SoA component
    long fstride;               // padded n_faces
    // per cell, ELL [slot][cell] with stride cstride (slot < max_faces)
    const int* cf_face;         // face id, -1 past the cell's last face
    const int* cf_nbr;          // neighbour across that face (cells.neighbour_cells_)
    const signed char* cf_out;  // outsign
    const double* volume;       // n_cells
    const double* cx;           // centroids, n_total each
    const double* cy;
    const double* cz;
    const double* wls_r;        // [6][cstride] r11 r12 r22 r13 r23 r33 (grid/gradient.h:153-160)
    // static per cell-face data, [(slot*3 + xyz)*cstride + cell], computed once at setup by
    // cell_static_body with the reference's formulas (the grid does not move):
    const double* cf_w;         // WLS weights w_1 w_2 w_3 of the neighbour across the face (gradient.h:96-108)
    const double* cf_d;         // face centre - cell centroid (limiter.cpp:44-47, convective_flux.cpp:286-290)
    // per face, SoA with stride fstride
    const int* f_left;
    const int* f_right;
    const double* f_area;
    const double* f_n;          // [3][fstride]
    const double* f_t1;
    const double* f_t2;
    const double* f_c;          // centres
    const double* f_geom;       // sides path: [face*10 + k] nx ny nz t1x t1y t1z t2x t2y t2z area
    // boundary faces in (tag alphabetical, marker order)
    const int* bf_face;
    const int* bf_tag;
    // actions: for phase p and tag t the actions are act[act_off[p*(n_tags+1)+t] ...)
    const BcAction* act;
    const int* act_off;
    int n_tags;
    // spline tables of the boundary_layer_profile actions: for an action with params {p, n, off}
    // bc_tab[off ..] = x[n] | v[n] | T[n] | v''[n] | T''[n]  (util/cubic_spline.cpp:9-89)
    const double* bc_tab;
    // tiles of `tile` consecutive owned cells for the tiled residual kernel (pass F): the
    // interfaces first listed by a cell of tile b ("home" interfaces) are the contiguous range
    // [home0[b], home0[b+1]); the other interfaces of its cells ("foreign", homed in an earlier
    // tile) are foreign[foreign_off[b] .. foreign_off[b+1]).  cf_slot is the ELL table of the
    // tile-local flux slot of every cell face: home slot = face - home0[b], foreign slot =
    // n_home + position in the foreign list; 0xFFFF past the cell's last face.
    int tile, n_tiles, tile_cap;
    // tiles of the sides path (layout.h): n_tiles owned tiles, then the tiles of the inner halo ring;
    // tile_info[64 b ..] = {c0, nc, h0, nh, nR, n0, ent_off, n_ent, n_own_l, n_slots, -, -, then up to 10 entries
    // inline (one 256-byte read gives a block everything it needs to issue its transfers)}, tile_ent[5 e ..] = {f_start, count,
    // section, slot_start, stage_start}; cf_stage = staging index of the cell's own side of each
    // of its interfaces (pass G'); stage_cap = largest staging buffer of a tile (sides)
    int n_tiles_all, stage_cap;
    // tiles of pass G' that touch no halo cell and no boundary ghost ("early": they can run while the
    // halo exchange is in flight) and the rest ("late": after the exchange and the ghost fill)
    int n_early, n_late;
    const int* tiles_early;
    const int* tiles_late;
    const int* tile_info;
    const int* tile_ent;
    const unsigned short* cf_stage;
    const int* home0;
    const int* foreign_off;
    const int* foreign;
    const unsigned short* cf_slot;
    // sides of the tile's interfaces whose cell is NOT a cell of the tile (neighbouring tiles,
    // halo cells, boundary ghosts): entries [out_off[b], out_off[b+1]) of (slot | side << 15, cell)
    // and the static offset face centre - centroid of that cell, out_d[xyz * n_out + e]
    const int* out_off;
    const unsigned short* out_slot;
    const int* out_cell;
    const double* out_d;
    long n_out;
    // pipelined sides path (fv_pipe.cu): per-tile packed records, each fetched by ONE bulk copy, and per-tile
    // side-state arrays that pass G' fills and pass F' fetches with ONE bulk copy.
    //   t_wd     [tile][2 * pipe_nf * 3][tile]  WLS weights, then face offsets, of the tile's cells
    //   t_grec   [tile], grec_stride bytes: u16 nbr[pipe_nf][tile] (tile-local index of the neighbour: < tile an own
    //            cell, >= tile entry (index - tile) of the halo list; bit 14: this cell is the LEFT cell of the
    //            interface, bit 15: the neighbour is a boundary ghost; an absent slot points at the cell itself) |
    //            u16 stage[pipe_nf][tile] (cf_stage) |
    //            int hdr[8] = {n_halo, nR, n_own_l, n_stage, has_ghost, home, -, -} (n_stage: records of the tile's
    //            staging image, an even number; home: first record of the tile's side-state array) |
    //            int halo[halo_cap] (cell ids of the neighbours outside the tile, no duplicates)
    //   side-state array of tile b (halo-ring tiles included): records [t_pinfo[2b], + t_pinfo[2b+1]) of
    //            FvArgs::sidesP, SIDE_REC S each: the tile's staging image (its cells' sides in the staging order of
    //            layout.h, the ghost sides of its boundary interfaces included).  Pass G' writes it with ONE bulk
    //            store; every side is stored once.
    //   t_pull   [owned tile], pull_stride ints: {home, n_home, n_slots, n_pull, n_image, -, -, -} |
    //            int pull[pull_cap][3] = {first record in FvArgs::sidesP, count, first record in the tile's image}:
    //            the image of the tile's interface sides pass F' assembles in shared memory = its own array, then
    //            the runs of the neighbouring tiles' arrays that hold the far sides of its surface interfaces
    //            (first records and counts are even: 16-byte granularity of the bulk copies at 40 bytes a record;
    //            a run may therefore carry one record more at either end than the tile needs)
    //   t_freca  [owned tile], freca_stride bytes: u16 lidx[tile_cap] | u16 ridx[tile_cap] (record, in the tile's
    //            image, of the left / right state of the interface in flux slot s) |
    //            double frame[tile_cap][6] (t1 t2 of that interface: the normal is cross(t1, t2) as in the reference,
    //            grid/interface.h:250, util/vector3.h:153-160, recomputed bit for bit in the kernels)
    //   t_frec   [owned tile], frec_stride bytes: double vol[tile] | double sarea[6][tile] (-area * outsign, 0 for an
    //            absent slot) | u16 slot[6][tile] (cf_slot)
    const double* t_wd;
    const unsigned char* t_grec;
    const unsigned char* t_freca;
    const unsigned char* t_frec;
    const int* t_pinfo;
    const int* t_pull;
    int grec_stride, freca_stride, frec_stride, halo_cap, pull_cap, pull_stride, p_cap, pipe_nf;
    // NOTE cf_slot, cf_out and cf_d are always MAXF slots wide (absent slots: 0xFFFF / 0 / 0) so
    // that the tiled kernel can load all slots unconditionally; cf_face, cf_nbr, cf_w are
    // max_faces wide.
};

// the five reconstructed variables (rho, thermo-2, vx, vy, vz) of pass F, pointers resolved on
// the host so that the unrolled kernels index them at compile time
template <typename S>
struct Recon {
    const S* val[5];
    const S* g[5][3];
    const S* lim[5];
};

// one gradient pass (pass G): NV variables with their sources and destinations
template <typename S, int NV>
struct GradVars {
    const S* src[NV];
    S* g[NV][3];
    S* lim[NV];
};

template <typename S>
struct FvArgs {
    GridDev g;
    S* fs[7];          // FlowStates: rho p T e vx vy vz
    S* dudt[5];        // mass mx my mz energy
    S* grad;           // [(var*3 + comp)*cstride + cell]
    S* lim;            // [var*cstride + cell]
    S* flux;           // general path: [comp*fstride + face]
    S* sidesP;         // sides path: the per-tile side-state arrays (GridDev::t_pinfo), SIDE_REC S per record
    // sides path: when not null, pass F' also takes the minimum over the owned cells of the local
    // time step (estimate_dt fused into the first evaluation of a Runge-Kutta step): the bit
    // pattern of a positive double orders like the double, so the reduction is an integer atomicMin
    unsigned long long* dt_out;
    // pipelined kernels: word 0 != 0 aborts every wait loop (set by the first warp whose wait on a stage barrier
    // exceeds the time limit, with a record of which barrier in words 1..7); read by the host after a synchronise
    unsigned int* pipe_dbg;
    double* face_T;    // face_fs of the viscous path (finite_volume.h:128), real parts
    // moving grid (GridBlock::face_vel, grid.h:599,638): [3][fstride], nullptr on a static grid.  Written by
    // face_vel_body, replaced by its local-frame round trip in the flux kernel as in the reference
    // (convective_flux.cpp:114-116,166-176), then read by constant_flux and the GCL term
    double* face_vel;
    double* face_rho;
    Recon<S> rc;
    Gas<S> gas;
    S mu0, T0, Ts, Pr;
    double eps, delta, signal_factor;
    int thermo;        // THERMO_*
    int th2;           // index into fs[] of the second thermo variable (p, T or e)
    int limiter_on;
    int order2;        // reconstruction order 2 in effect for this evaluation
    int viscous;
};

// CubicSpline::eval (util/cubic_spline.h:14-45): x, y, y'' are n-point tables.  The reference's
// constructor sizes y'' for the n - 2 interior knots but indexes it by knot (cubic_spline.cpp:46,
// 80-83 and cubic_spline.h:33-40), so that its evaluations in the LAST TWO intervals read past the
// end of the allocation; the tables here hold the n values those formulas address (natural
// spline: y''_0 = y''_{n-1} = 0), which is what the reference computes wherever it is defined.
IBIS_HD double spline_eval(const double* x_, const double* y_, const double* ydd_, int n, double x) {
    if (x <= x_[0]) return y_[0];
    if (x >= x_[n - 1]) return y_[n - 1];
    int idx = 0;
    for (int i = 0; i < n; i++)
        if (x_[i] >= x) { idx = i - 1; break; }
    const double dx_plus = x_[idx + 1] - x;
    const double dx_minus = x - x_[idx];
    const double delta_xi = x_[idx + 1] - x_[idx];
    const double a = (ydd_[idx] * dx_plus * dx_plus * dx_plus + ydd_[idx + 1] * dx_minus * dx_minus * dx_minus) /
                     (6 * delta_xi);
    const double b = (y_[idx] / delta_xi - ydd_[idx] * delta_xi / 6) * dx_plus;
    const double c = (y_[idx + 1] / delta_xi - ydd_[idx + 1] * delta_xi / 6) * dx_minus;
    return a + b + c;
}

// ---------------------------------------------------------------------------------------
// boundary conditions: one thread per boundary face applies its tag's action list in order
// (finite_volume.cpp:187-218, boundaries/boundary.cpp:517-543)
// ---------------------------------------------------------------------------------------
template <typename S>
IBIS_HD void bc_body(const FvArgs<S>& A, int k, int phase) {
    const GridDev& g = A.g;
    const int f = g.bf_face[k];
    const int tag = g.bf_tag[k];
    const int l = g.f_left[f], r = g.f_right[f];
    int ghost, valid;
    if (l < g.n_cells) { ghost = r; valid = l; } else { ghost = l; valid = r; }
    const int a0 = g.act_off[phase * (g.n_tags + 1) + tag];
    const int a1 = g.act_off[phase * (g.n_tags + 1) + tag + 1];
    S* const* fs = A.fs;
    typedef mem<S> M;
    for (int a = a0; a < a1; a++) {
        const BcAction& act = g.act[a];
        switch (act.type) {
            case BC_FLOW_STATE_COPY: {                       // boundary.cpp:26-45
                M::store(fs[F_T], ghost, S(act.params[2]));
                M::store(fs[F_P], ghost, S(act.params[1]));
                M::store(fs[F_RHO], ghost, S(act.params[0]));
                M::store(fs[F_E], ghost, S(act.params[3]));
                M::store(fs[F_VX], ghost, S(act.params[4]));
                M::store(fs[F_VY], ghost, S(act.params[5]));
                M::store(fs[F_VZ], ghost, S(act.params[6]));
                break;
            }
            case BC_INTERNAL_COPY: {                         // boundary.cpp:110-136
                for (int q = 0; q < 7; q++) M::store(fs[q], ghost, M::load(fs[q], valid));
                break;
            }
            case BC_REFLECT_NORMAL: {                        // boundary.cpp:152-205
                for (int q = 0; q < 4; q++) M::store(fs[q], ghost, M::load(fs[q], valid));
                S x = M::load(fs[F_VX], valid), y = M::load(fs[F_VY], valid),
                  z = M::load(fs[F_VZ], valid);
                S nx = S(g.f_n[f]), ny = S(g.f_n[g.fstride + f]), nz = S(g.f_n[2 * g.fstride + f]);
                S t1x = S(g.f_t1[f]), t1y = S(g.f_t1[g.fstride + f]),
                  t1z = S(g.f_t1[2 * g.fstride + f]);
                S t2x = S(g.f_t2[f]), t2y = S(g.f_t2[g.fstride + f]),
                  t2z = S(g.f_t2[2 * g.fstride + f]);
                S xs = -(x * nx + y * ny + z * nz);
                S ys = x * t1x + y * t1y + z * t1z;
                S zs = x * t2x + y * t2y + z * t2z;
                M::store(fs[F_VX], ghost, xs * nx + ys * t1x + zs * t2x);
                M::store(fs[F_VY], ghost, xs * ny + ys * t1y + zs * t2y);
                M::store(fs[F_VZ], ghost, xs * nz + ys * t1z + zs * t2z);
                break;
            }
            case BC_VEL_REFLECT: {                           // boundary.cpp:219-240
                M::store(fs[F_VX], ghost, -M::load(fs[F_VX], valid));
                M::store(fs[F_VY], ghost, -M::load(fs[F_VY], valid));
                M::store(fs[F_VZ], ghost, -M::load(fs[F_VZ], valid));
                break;
            }
            case BC_FIX_TEMPERATURE: {                       // boundary.cpp:255-276
                S Tw = S(act.params[0]);
                M::store(fs[F_T], ghost, 2.0 * Tw - M::load(fs[F_T], valid));
                break;
            }
            case BC_SUBSONIC_INFLOW: {                       // boundary.cpp:296-342
                S in_rho = S(act.params[0]), in_p = S(act.params[1]);
                S in_vx = S(act.params[4]), in_vy = S(act.params[5]), in_vz = S(act.params[6]);
                S nx = S(g.f_n[f]), ny = S(g.f_n[g.fstride + f]), nz = S(g.f_n[2 * g.fstride + f]);
                S vvx = M::load(fs[F_VX], valid), vvy = M::load(fs[F_VY], valid),
                  vvz = M::load(fs[F_VZ], valid);
                S vel_correct = nx * (in_vx - vvx) + ny * (in_vy - vvy) + nz * (in_vz - vvz);
                S a = speed_of_sound(M::load(fs[F_T], valid), A.gas.R, A.gas.gamma);
                S rho_ref = M::load(fs[F_RHO], valid);
                S p_v = M::load(fs[F_P], valid);
                S p_face = 0.5 * (in_p + p_v - rho_ref * a * vel_correct);
                S rho_face = in_rho + (p_face - in_p) / a / a;
                S vx_face = in_vx - nx * (in_p - p_face) / (rho_ref * a);
                S vy_face = in_vy - ny * (in_p - p_face) / (rho_ref * a);
                S vz_face = in_vz - nz * (in_p - p_face) / (rho_ref * a);
                S pg = 2.0 * p_face - p_v;
                S rg = 2.0 * rho_face - rho_ref;
                S Tg = T_from_rhop(rg, pg, A.gas.R);
                M::store(fs[F_P], ghost, pg);
                M::store(fs[F_RHO], ghost, rg);
                M::store(fs[F_T], ghost, Tg);
                M::store(fs[F_E], ghost, A.gas.Cv * Tg);
                M::store(fs[F_VX], ghost, 2.0 * vx_face - vvx);
                M::store(fs[F_VY], ghost, 2.0 * vy_face - vvy);
                M::store(fs[F_VZ], ghost, 2.0 * vz_face - vvz);
                break;
            }
            case BC_SUBSONIC_OUTFLOW: {                      // boundary.cpp:359-399
                S nx = S(g.f_n[f]), ny = S(g.f_n[g.fstride + f]), nz = S(g.f_n[2 * g.fstride + f]);
                S p_face = S(act.params[0]);
                S rho_ref = M::load(fs[F_RHO], valid);
                S p_ref = M::load(fs[F_P], valid);
                S a = speed_of_sound(M::load(fs[F_T], valid), A.gas.R, A.gas.gamma);
                S vvx = M::load(fs[F_VX], valid), vvy = M::load(fs[F_VY], valid),
                  vvz = M::load(fs[F_VZ], valid);
                S rho_face = rho_ref + (p_face - p_ref) / a / a;
                S vx_face = vvx + nx * (p_ref - p_face) / (rho_ref * a);
                S vy_face = vvy + ny * (p_ref - p_face) / (rho_ref * a);
                S vz_face = vvz + nz * (p_ref - p_face) / (rho_ref * a);
                S pg = 2.0 * p_face - p_ref;
                S rg = 2.0 * rho_face - rho_ref;
                S Tg = T_from_rhop(rg, pg, A.gas.R);
                M::store(fs[F_P], ghost, pg);
                M::store(fs[F_RHO], ghost, rg);
                M::store(fs[F_T], ghost, Tg);
                M::store(fs[F_E], ghost, A.gas.Cv * Tg);
                M::store(fs[F_VX], ghost, 2.0 * vx_face - vvx);
                M::store(fs[F_VY], ghost, 2.0 * vy_face - vvy);
                M::store(fs[F_VZ], ghost, 2.0 * vz_face - vvz);
                break;
            }
            case BC_BOUNDARY_LAYER_PROFILE: {                // boundary.cpp:61-95
                const int n = (int)act.params[1];
                const double* tab = g.bc_tab + (long)act.params[2];
                const double pos = g.cy[ghost];              // real_part of the ghost centroid's y
                S p = S(act.params[0]);
                S T = S(spline_eval(tab, tab + 2 * n, tab + 4 * n, n, pos));
                M::store(fs[F_P], ghost, p);
                M::store(fs[F_T], ghost, T);
                M::store(fs[F_RHO], ghost, rho_from_pT(p, T, A.gas.R));       // update_thermo_from_pT
                M::store(fs[F_E], ghost, A.gas.Cv * T);
                M::store(fs[F_VX], ghost, S(spline_eval(tab, tab + n, tab + 3 * n, n, pos)));
                M::store(fs[F_VY], ghost, S(0.0));
                M::store(fs[F_VZ], ghost, S(0.0));
                break;
            }
            case BC_CONSTANT_FLUX: {                         // boundary.cpp:405-444 (a FluxAction: phase 2)
                const long fs_ = g.fstride;
                S nx = S(g.f_n[f]), ny = S(g.f_n[g.fstride + f]), nz = S(g.f_n[2 * g.fstride + f]);
                S rho = S(act.params[0]), p = S(act.params[1]), T = S(act.params[2]);
                S vx = S(act.params[4]), vy = S(act.params[5]), vz = S(act.params[6]);
                S u = A.gas.Cv * T;                          // gas_model.internal_energy (gas_model.h:164)
                const double* fv = A.face_vel;                                  // nullptr: static grid
                S rel_vx = vx - (fv ? fv[f] : 0.0), rel_vy = vy - (fv ? fv[fs_ + f] : 0.0),
                  rel_vz = vz - (fv ? fv[2 * fs_ + f] : 0.0);
                S mass_flux = rho * (rel_vx * nx + rel_vy * ny + rel_vz * nz);
                M::store(A.flux, 0 * fs_ + f, mass_flux);
                M::store(A.flux, 1 * fs_ + f, p * nx + vx * mass_flux);
                M::store(A.flux, 2 * fs_ + f, p * ny + vy * mass_flux);
                if (g.dim == 3) M::store(A.flux, 3 * fs_ + f, p * nz + vz * mass_flux);
                // `vy * nx` is the reference's (boundary.cpp:441), kept for parity
                M::store(A.flux, 4 * fs_ + f, mass_flux * (u + 0.5 * (vx * vx + vy * vy + vz * vz)) +
                                                  p * (vx * nx + vy * nx + vz * nz));
                break;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// WLS setup: r-factors of one cell (grid/gradient.h:119-162)
// ---------------------------------------------------------------------------------------
IBIS_HD void wls_r_body(const GridDev& g, double* r_out, int c) {
    double sxx = 0.0, sxy = 0.0, sxz = 0.0, syy = 0.0, syz = 0.0, szz = 0.0;
    const double xi = g.cx[c], yi = g.cy[c], zi = g.cz[c];
    for (int j = 0; j < g.max_faces; j++) {
        if (g.cf_face[j * g.cstride + c] < 0) break;
        const int nb = g.cf_nbr[j * g.cstride + c];
        const double dx = g.cx[nb] - xi, dy = g.cy[nb] - yi, dz = g.cz[nb] - zi;
        sxx += dx * dx;
        sxy += dx * dy;
        sxz += dx * dz;
        syy += dy * dy;
        syz += dy * dz;
        szz += dz * dz;
    }
    const double r11 = sqrt(sxx);
    const double r12 = 1.0 / r11 * sxy;
    const double r22 = sqrt(syy - r12 * r12);
    const double r13 = 1.0 / r11 * sxz;
    const double r23 = 1.0 / r22 * (syz - r12 / r11 * sxz);
    const double r33 = sqrt(szz - (r13 * r13 + r23 * r23));
    r_out[0 * g.cstride + c] = r11;
    r_out[1 * g.cstride + c] = r12;
    r_out[2 * g.cstride + c] = r22;
    if (g.dim == 3) {
        r_out[3 * g.cstride + c] = r13;
        r_out[4 * g.cstride + c] = r23;
        r_out[5 * g.cstride + c] = r33;
    }
}

// static per cell-face geometry: WLS weights of every face neighbour (gradient.h:84-108) and
// the face-centre offsets.  Written once at setup; the formulas and their association are the
// reference's, with the loop-invariant factors (1/r22^2, r12/r11, r23/r22, 1/r33^2) evaluated
// once per cell -- the same operations on the same operands, hence the same bits.
template <int DIM>
IBIS_HD void cell_static_body(const GridDev& g, double* w_out, double* d_out, int c) {
    const long cs = g.cstride, fs_ = g.fstride;
    const double r11 = g.wls_r[c], r12 = g.wls_r[cs + c], r22 = g.wls_r[2 * cs + c];
    double r13 = 0.0, r23 = 0.0, r33 = 0.0;
    if (DIM == 3) {
        r13 = g.wls_r[3 * cs + c];
        r23 = g.wls_r[4 * cs + c];
        r33 = g.wls_r[5 * cs + c];
    }
    const double beta = (r12 * r23 - r13 * r23) / (r11 * r22);     // gradient.h:89 (sic)
    const double r11sq = r11 * r11;
    const double inv_r22sq = 1.0 / (r22 * r22);
    const double r12_r11 = r12 / r11;
    const double r23_r22 = r23 / r22;
    const double inv_r33sq = (DIM == 3) ? 1.0 / (r33 * r33) : 0.0;
    const double xi = g.cx[c], yi = g.cy[c], zi = g.cz[c];
    for (int j = 0; j < g.max_faces; j++) {
        const int f = g.cf_face[j * cs + c];
        if (f < 0) break;
        const int nb = g.cf_nbr[j * cs + c];
        const double dx = g.cx[nb] - xi;
        const double dy = g.cy[nb] - yi;
        const double dz = g.cz[nb] - zi;
        const double alpha_1 = dx / r11sq;
        const double alpha_2 = inv_r22sq * (dy - r12_r11 * dx);
        double alpha_3 = 0.0;
        if (DIM == 3) alpha_3 = inv_r33sq * (dz - r23_r22 * dy + beta * dx);
        w_out[(j * 3 + 0) * cs + c] = alpha_1 - r12_r11 * alpha_2 + beta * alpha_3;
        w_out[(j * 3 + 1) * cs + c] = alpha_2 - r23_r22 * alpha_3;
        w_out[(j * 3 + 2) * cs + c] = alpha_3;
        d_out[(j * 3 + 0) * cs + c] = g.f_c[f] - xi;
        d_out[(j * 3 + 1) * cs + c] = g.f_c[fs_ + f] - yi;
        d_out[(j * 3 + 2) * cs + c] = g.f_c[2 * fs_ + f] - zi;
    }
}

// Barth-Jespersen limiters of one cell from its gradients and neighbour extrema
// (finite_volume/limiter.cpp:41-60).  Generic form: the reference's loop, one division per face.
template <int DIM, int NV, int NF, typename S>
IBIS_HD void bj_limiters(const GridDev& g, int c, double eps, const S* ui, const S* gx, const S* gy, const S* gz,
                         const S* umin, const S* umax, S* phi) {
    const long cs = g.cstride;
#pragma unroll
    for (int v = 0; v < NV; v++) phi[v] = S(1.0);
#pragma unroll
    for (int j = 0; j < NF; j++) {
        // absent slot: zero offsets -> delta_2 = 0 -> sgn = 0 -> no constraint
        const double dx = g.cf_d[(j * 3 + 0) * cs + c];
        const double dy = g.cf_d[(j * 3 + 1) * cs + c];
        const double dz = g.cf_d[(j * 3 + 2) * cs + c];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            S delta_2 = gx[v] * dx + gy[v] * dy + gz[v] * dz;
            const int sgn = (re(delta_2) > 0.0) - (re(delta_2) < 0.0);
            delta_2 = (double)sgn * (abs_(delta_2) + eps);
            if (sgn > 0) {
                phi[v] = min_(phi[v], (umax[v] - ui[v]) / delta_2);
            } else if (sgn < 0) {
                phi[v] = min_(phi[v], (umin[v] - ui[v]) / delta_2);
            }
        }
    }
}

// Real numbers: the same limiter with TWO divisions per variable instead of one per face.
// phi = min(1, min_{j: d_j > 0} (Umax - Ui) / (|d_j| + eps), min_{j: d_j < 0} (Umin - Ui) / -(|d_j| + eps)).
// IEEE division is correctly rounded, hence monotonic: for a fixed numerator a >= 0 the smallest
// quotient fl(a / m_j) is fl(a / max_j m_j), bit for bit, and fl((-a) / (-m)) = fl(a / m).  So the
// minimum over the faces of one sign is the single quotient by the largest |delta_2| + eps of that
// sign -- the same double the reference's loop ends with, at a third of the division count.
template <int DIM, int NV, int NF>
IBIS_HD void bj_limiters(const GridDev& g, int c, double eps, const double* ui, const double* gx, const double* gy,
                         const double* gz, const double* umin, const double* umax, double* phi) {
    const long cs = g.cstride;
    // x -> |x| + eps is increasing in |x|, so the largest |delta_2| + eps among the faces of one
    // sign belongs to the largest (most negative) delta_2 itself: track the signed extrema only.
    // An absent ELL slot has zero offsets, delta_2 = 0, and changes neither extremum's sign.
    double dmax[NV], dmin[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        dmax[v] = 0.0;
        dmin[v] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const double dx = g.cf_d[(j * 3 + 0) * cs + c];
        const double dy = g.cf_d[(j * 3 + 1) * cs + c];
        const double dz = g.cf_d[(j * 3 + 2) * cs + c];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const double delta_2 = gx[v] * dx + gy[v] * dy + gz[v] * dz;
            dmax[v] = dmax[v] < delta_2 ? delta_2 : dmax[v];
            dmin[v] = delta_2 < dmin[v] ? delta_2 : dmin[v];
        }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
        double p = 1.0;
        if (dmax[v] > 0.0) p = min_(p, div_z(umax[v] - ui[v], dmax[v] + eps));
        if (dmin[v] < 0.0) p = min_(p, div_z(umin[v] - ui[v], -(-dmin[v] + eps)));
        phi[v] = p;
    }
}

// ---------------------------------------------------------------------------------------
// pass G: WLS gradients (grid/gradient.h:66-116) of NV variables and their Barth-Jespersen
// limiters (finite_volume/limiter.cpp:24-63) for one cell, fused, from the static weights and
// offsets: per cell the only gathers left are the neighbours' NV values.
//
// Geometry is arithmetic on plain doubles also in the dual instantiation: the reference
// carries it as duals with a zero dual part, and x + 0*y = x, so values and derivatives are
// the same.  NF = ELL width (max faces per cell of the grid).
// ---------------------------------------------------------------------------------------
template <int DIM, int NV, int NF, typename S>
IBIS_HD void grad_limit_body(const FvArgs<S>& A, const GradVars<S, NV>& V, int c, bool do_limit) {
    const GridDev& g = A.g;
    typedef mem<S> M;
    const long cs = g.cstride;
    S ui[NV], gx[NV], gy[NV], gz[NV], umin[NV], umax[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        ui[v] = M::ldg(V.src[v], c);
        gx[v] = S(0.0);
        gy[v] = S(0.0);
        gz[v] = S(0.0);
        umin[v] = ui[v];
        umax[v] = ui[v];
    }
    // Branch-free over the ELL slots so that every load of the cell can be in flight at once:
    // an absent slot (cell with fewer than NF faces) has zero weights and zero offsets and is
    // pointed at the cell itself, so it adds w * (u_i - u_i) = 0 and leaves min / max alone.
    int nbr[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const int nb = g.cf_nbr[j * cs + c];
        nbr[j] = nb < 0 ? c : nb;
    }
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const int nb = nbr[j];
        const double w_1 = g.cf_w[(j * 3 + 0) * cs + c];
        const double w_2 = g.cf_w[(j * 3 + 1) * cs + c];
        const double w_3 = (DIM == 3) ? g.cf_w[(j * 3 + 2) * cs + c] : 0.0;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const S un = M::ldg(V.src[v], nb);
            const S diff_u = un - ui[v];
            gx[v] += w_1 * diff_u;
            gy[v] += w_2 * diff_u;
            gz[v] += w_3 * diff_u;
            umin[v] = min_(umin[v], un);                       // limiter.cpp:37-38
            umax[v] = max_(umax[v], un);
        }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
        M::store(V.g[v][0], c, gx[v]);
        M::store(V.g[v][1], c, gy[v]);
        M::store(V.g[v][2], c, gz[v]);
    }
    if (!do_limit) return;

    S phi[NV];
    bj_limiters<DIM, NV, NF>(g, c, A.eps, ui, gx, gy, gz, umin, umax, phi);
#pragma unroll
    for (int v = 0; v < NV; v++) M::store(V.lim[v], c, phi[v]);
}

// ---------------------------------------------------------------------------------------
// reconstruction of one side of one face (convective_flux.cpp:212-420) -> Side in the
// GLOBAL frame (u,v,w = vx,vy,vz); rotation is applied by the caller.  All inputs go through
// the read-only path: no thread of a flux kernel writes a FlowStates / gradient / limiter array.
// ---------------------------------------------------------------------------------------
template <int DIM, typename S>
IBIS_HD S interp(const FvArgs<S>& A, int v, int cell, bool valid, bool limit, double dx, double dy,
                 double dz) {
    // linear_interpolate (convective_flux.cpp:243-255); v = 0 rho, 1 thermo-2, 2 vx, 3 vy, 4 vz
    typedef mem<S> M;
    const S value = M::ldg(A.rc.val[v], cell);
    // 2-D: every z-gradient, and the whole vz gradient, is identically zero in the reference
    // (never computed, convective_flux.cpp:206-208) and its vz limiter is 1; a ghost cell has
    // zero gradient and unit limiter (convective_flux.cpp:281-300): value + 1 * 0 = value
    if (!valid || (DIM == 2 && v == 4)) return value;
    const S gx = M::ldg(A.rc.g[v][0], cell);
    const S gy = M::ldg(A.rc.g[v][1], cell);
    S r;
    if (DIM == 3) r = gx * dx + gy * dy + M::ldg(A.rc.g[v][2], cell) * dz;
    else r = gx * dx + gy * dy + S(0.0) * dz;
    if (limit) return value + M::ldg(A.rc.lim[v], cell) * r;
    return value + S(1.0) * r;
}

// one side from the cell's values / gradients / limiters and the offset (dx, dy, dz) = face
// centre - cell centroid
template <int DIM, bool O2, typename S>
IBIS_HD Side<S> reconstruct_at(const FvArgs<S>& A, int cell, double dx, double dy, double dz) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    Side<S> s;
    if (!O2) {                                                // copy_reconstruct :212-240
        s.T = M::ldg(A.fs[F_T], cell);
        s.p = M::ldg(A.fs[F_P], cell);
        s.rho = M::ldg(A.fs[F_RHO], cell);
        s.e = M::ldg(A.fs[F_E], cell);
        s.u = M::ldg(A.fs[F_VX], cell);
        s.v = M::ldg(A.fs[F_VY], cell);
        s.w = M::ldg(A.fs[F_VZ], cell);
        return s;
    }
    const bool valid = cell < g.n_cells;
    const bool limit = A.limiter_on != 0;
    // the two thermodynamic variables, then the closure (gas_model.h:91-117)
    s.p = S(0.0);
    s.T = S(0.0);
    s.e = S(0.0);
    s.rho = interp<DIM>(A, 0, cell, valid, limit, dx, dy, dz);
    const S th2 = interp<DIM>(A, 1, cell, valid, limit, dx, dy, dz);
    if (A.thermo == THERMO_RHO_P) s.p = th2;
    else if (A.thermo == THERMO_RHO_T) s.T = th2;
    else s.e = th2;
    thermo_close(A.thermo, A.gas, s.rho, s.p, s.T, s.e);
    s.u = interp<DIM>(A, 2, cell, valid, limit, dx, dy, dz);
    s.v = interp<DIM>(A, 3, cell, valid, limit, dx, dy, dz);
    s.w = interp<DIM>(A, 4, cell, valid, limit, dx, dy, dz);
    return s;
}

template <int DIM, bool O2, typename S>
IBIS_HD Side<S> reconstruct_side(const FvArgs<S>& A, int f, int cell) {
    const GridDev& g = A.g;
    double dx = 0.0, dy = 0.0, dz = 0.0;
    if (O2) {
        dx = g.f_c[f] - g.cx[cell];
        dy = g.f_c[g.fstride + f] - g.cy[cell];
        dz = g.f_c[2 * g.fstride + f] - g.cz[cell];
    }
    return reconstruct_at<DIM, O2>(A, cell, dx, dy, dz);
}

// convective flux of one interface in the GLOBAL frame: reconstruct, rotate
// (util/vector3.cpp:142-153), flux calculator, rotate back (convective_flux.cpp:148-164).
// The face frame (n, t1, t2) enters as plain doubles (zero dual part in the reference).
// MOVING (grid.moving(), convective_flux.cpp:105-140,166-176): the interface velocity is subtracted from
// both sides, rotated into the face frame, the flux is corrected from the moving frame back to the
// inertial one, and the velocity is rotated back (and stored: that round trip is what the GCL sees).
template <int DIM, int FLUX, bool O2, typename S, bool MOVING = false>
IBIS_HD Flux5<S> face_convective_flux(const FvArgs<S>& A, int f) {
    const GridDev& g = A.g;
    const long fs_ = g.fstride;
    Side<S> L = reconstruct_side<DIM, O2>(A, f, g.f_left[f]);
    Side<S> R_ = reconstruct_side<DIM, O2>(A, f, g.f_right[f]);
    const double nx = g.f_n[f], ny = g.f_n[fs_ + f], nz = g.f_n[2 * fs_ + f];
    const double t1x = g.f_t1[f], t1y = g.f_t1[fs_ + f], t1z = g.f_t1[2 * fs_ + f];
    const double t2x = g.f_t2[f], t2y = g.f_t2[fs_ + f], t2z = g.f_t2[2 * fs_ + f];
    double lx = 0.0, ly = 0.0, lz = 0.0;       // interface velocity in the face frame
    if (MOVING) {
        const double vx = A.face_vel[f], vy = A.face_vel[fs_ + f], vz = A.face_vel[2 * fs_ + f];
        L.u = L.u - vx; L.v = L.v - vy; L.w = L.w - vz;
        R_.u = R_.u - vx; R_.v = R_.v - vy; R_.w = R_.w - vz;
        lx = vx * nx + vy * ny + vz * nz;
        ly = vx * t1x + vy * t1y + vz * t1z;
        lz = vx * t2x + vy * t2y + vz * t2z;
    }
    {
        const S x = L.u * nx + L.v * ny + L.w * nz;
        const S y = L.u * t1x + L.v * t1y + L.w * t1z;
        const S z = L.u * t2x + L.v * t2y + L.w * t2z;
        L.u = x; L.v = y; L.w = z;
    }
    {
        const S x = R_.u * nx + R_.v * ny + R_.w * nz;
        const S y = R_.u * t1x + R_.v * t1y + R_.w * t1z;
        const S z = R_.u * t2x + R_.v * t2y + R_.w * t2z;
        R_.u = x; R_.v = y; R_.w = z;
    }
    Flux5<S> F = flux_calc<FLUX>(L, R_, A.gas, A.delta);
    if (MOVING) {
        const double v_sqr = lx * lx + ly * ly + lz * lz;
        F.energy += 0.5 * F.mass * v_sqr + F.mx * lx + F.my * ly;
        if (DIM == 3) F.energy += F.mz * lz;
        F.mx += lx * F.mass;
        F.my += ly * F.mass;
        if (DIM == 3) F.mz += lz * F.mass;
        A.face_vel[f] = lx * nx + ly * t1x + lz * t2x;
        A.face_vel[fs_ + f] = lx * ny + ly * t1y + lz * t2y;
        A.face_vel[2 * fs_ + f] = lx * nz + ly * t1z + lz * t2z;
    }
    const S px = F.mx, py = F.my;
    const S pz = (DIM == 3) ? F.mz : S(0.0);
    F.mx = px * nx + py * t1x + pz * t2x;
    F.my = px * ny + py * t1y + pz * t2y;
    F.mz = px * nz + py * t1z + pz * t2z;
    return F;
}

// ---------------------------------------------------------------------------------------
// viscous flux of one interface (finite_volume/viscous_flux.cpp:16-184), subtracted from F
// ---------------------------------------------------------------------------------------
template <typename S>
struct Vec3 {
    S x, y, z;
};

template <typename S>
IBIS_HD Vec3<S> hasselbacher(const FvArgs<S>& A, int var, const S& vl, const S& vr, int l, int r,
                             const Vec3<S>& ehat, const Vec3<S>& n, const S& len_e,
                             const S& ehat_dot_n) {
    typedef mem<S> M;
    const long cs = A.g.cstride;
    S ax = 0.5 * (M::load(A.grad, (var * 3 + 0) * cs + l) + M::load(A.grad, (var * 3 + 0) * cs + r));
    S ay = 0.5 * (M::load(A.grad, (var * 3 + 1) * cs + l) + M::load(A.grad, (var * 3 + 1) * cs + r));
    S az = 0.5 * (M::load(A.grad, (var * 3 + 2) * cs + l) + M::load(A.grad, (var * 3 + 2) * cs + r));
    S avg_dot_ehat = ax * ehat.x + ay * ehat.y + az * ehat.z;
    S correction = avg_dot_ehat - (vr - vl) / len_e;
    Vec3<S> o;
    o.x = ax - correction * n.x / ehat_dot_n;
    o.y = ay - correction * n.y / ehat_dot_n;
    o.z = az - correction * n.z / ehat_dot_n;
    return o;
}

template <int DIM, typename S>
IBIS_HD void face_viscous_flux(const FvArgs<S>& A, int f, Flux5<S>& F) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    const long cs = g.cstride, fs_ = g.fstride;
    const int l = g.f_left[f], r = g.f_right[f];
    const bool lv = l < g.n_cells, rv = r < g.n_cells;
    const Vec3<S> n = {S(g.f_n[f]), S(g.f_n[fs_ + f]), S(g.f_n[2 * fs_ + f])};
    Vec3<S> gT, gvx, gvy, gvz;
    if (!lv || !rv) {                                         // viscous_flux.cpp:94-96
        const int ic = lv ? l : r;
        gT = {M::load(A.grad, (V_TVISC * 3 + 0) * cs + ic), M::load(A.grad, (V_TVISC * 3 + 1) * cs + ic),
              M::load(A.grad, (V_TVISC * 3 + 2) * cs + ic)};
        gvx = {M::load(A.grad, (V_VX * 3 + 0) * cs + ic), M::load(A.grad, (V_VX * 3 + 1) * cs + ic),
               M::load(A.grad, (V_VX * 3 + 2) * cs + ic)};
        gvy = {M::load(A.grad, (V_VY * 3 + 0) * cs + ic), M::load(A.grad, (V_VY * 3 + 1) * cs + ic),
               M::load(A.grad, (V_VY * 3 + 2) * cs + ic)};
        gvz = {M::load(A.grad, (V_VZ * 3 + 0) * cs + ic), M::load(A.grad, (V_VZ * 3 + 1) * cs + ic),
               M::load(A.grad, (V_VZ * 3 + 2) * cs + ic)};
    } else {                                                  // viscous_flux.cpp:53-80
        const S ex = S(g.cx[r]) - S(g.cx[l]);
        const S ey = S(g.cy[r]) - S(g.cy[l]);
        const S ez = S(g.cz[r]) - S(g.cz[l]);
        const S len_e = sqrt_(ex * ex + ey * ey + ez * ez);
        const Vec3<S> ehat = {ex / len_e, ey / len_e, ez / len_e};
        const S ehat_dot_n = ehat.x * n.x + ehat.y * n.y + ehat.z * n.z;
        gvx = hasselbacher(A, V_VX, M::load(A.fs[F_VX], l), M::load(A.fs[F_VX], r), l, r, ehat, n,
                           len_e, ehat_dot_n);
        gvy = hasselbacher(A, V_VY, M::load(A.fs[F_VY], l), M::load(A.fs[F_VY], r), l, r, ehat, n,
                           len_e, ehat_dot_n);
        gvz = hasselbacher(A, V_VZ, M::load(A.fs[F_VZ], l), M::load(A.fs[F_VZ], r), l, r, ehat, n,
                           len_e, ehat_dot_n);
        gT = hasselbacher(A, V_TVISC, M::load(A.fs[F_T], l), M::load(A.fs[F_T], r), l, r, ehat, n,
                          len_e, ehat_dot_n);
    }
    // face flow state: average_flow_states_pT + update_thermo_from_pT
    // (gas/flow_state.h:89-91, gas/gas_state.h:108-113, viscous_flux.cpp:103-104)
    const S p = 0.5 * (M::load(A.fs[F_P], l) + M::load(A.fs[F_P], r));
    const S T = 0.5 * (M::load(A.fs[F_T], l) + M::load(A.fs[F_T], r));
    const S rho = rho_from_pT(p, T, A.gas.R);
    // Vector3s::average_vectors (util/vector3.h)
    const S vx = 0.5 * (M::load(A.fs[F_VX], l) + M::load(A.fs[F_VX], r));
    const S vy = 0.5 * (M::load(A.fs[F_VY], l) + M::load(A.fs[F_VY], r));
    const S vz = 0.5 * (M::load(A.fs[F_VZ], l) + M::load(A.fs[F_VZ], r));
    A.face_T[f] = re(T);
    A.face_rho[f] = re(rho);

    const S mu = viscosity(T, A.mu0, A.T0, A.Ts);
    const S k = conductivity(viscosity(T, A.mu0, A.T0, A.Ts), A.gas.Cp, A.Pr);
    const S lambda = -2.0 / 3.0 * mu;
    const S bulk = lambda * (gvx.x + gvy.y + gvz.z);
    const S tau_xx = 2.0 * mu * gvx.x + bulk;
    const S tau_yy = 2.0 * mu * gvy.y + bulk;
    const S tau_zz = 2.0 * mu * gvz.z + bulk;
    const S tau_xy = mu * (gvx.y + gvy.x);
    const S tau_xz = mu * (gvx.z + gvz.x);
    const S tau_yz = mu * (gvy.z + gvz.y);
    const S theta_x = vx * tau_xx + vy * tau_xy + vz * tau_xz + k * gT.x;
    const S theta_y = vx * tau_xy + vy * tau_yy + vz * tau_yz + k * gT.y;
    const S theta_z = vx * tau_xz + vy * tau_yz + vz * tau_zz + k * gT.z;
    F.mx -= n.x * tau_xx + n.y * tau_xy + n.z * tau_xz;
    F.my -= n.x * tau_xy + n.y * tau_yy + n.z * tau_yz;
    if (DIM == 3) F.mz -= n.x * tau_xz + n.y * tau_yz + n.z * tau_zz;
    F.energy -= n.x * theta_x + n.y * theta_y + n.z * theta_z;
}

// ---------------------------------------------------------------------------------------
// general path, interface kernels: write / update the face flux array
// ---------------------------------------------------------------------------------------
template <int DIM, int FLUX, bool O2, typename S, bool MOVING = false>
IBIS_HD void face_flux_body(const FvArgs<S>& A, int f) {
    typedef mem<S> M;
    const long fs_ = A.g.fstride;
    Flux5<S> F = face_convective_flux<DIM, FLUX, O2, S, MOVING>(A, f);
    M::store(A.flux, 0 * fs_ + f, F.mass);
    M::store(A.flux, 1 * fs_ + f, F.mx);
    M::store(A.flux, 2 * fs_ + f, F.my);
    if (DIM == 3) M::store(A.flux, 3 * fs_ + f, F.mz);
    M::store(A.flux, 4 * fs_ + f, F.energy);
}

template <int DIM, typename S>
IBIS_HD void face_viscous_body(const FvArgs<S>& A, int f) {
    typedef mem<S> M;
    const long fs_ = A.g.fstride;
    Flux5<S> F;
    F.mass = S(0.0);
    F.mx = M::load(A.flux, 1 * fs_ + f);
    F.my = M::load(A.flux, 2 * fs_ + f);
    F.mz = (DIM == 3) ? M::load(A.flux, 3 * fs_ + f) : S(0.0);
    F.energy = M::load(A.flux, 4 * fs_ + f);
    face_viscous_flux<DIM>(A, f, F);
    M::store(A.flux, 1 * fs_ + f, F.mx);
    M::store(A.flux, 2 * fs_ + f, F.my);
    if (DIM == 3) M::store(A.flux, 3 * fs_ + f, F.mz);
    M::store(A.flux, 4 * fs_ + f, F.energy);
}

// ---------------------------------------------------------------------------------------
// moving grid: GridBlock::compute_geometric_data (grid.h:169-178) on the device-resident layout after the
// vertices moved -- interface centres and areas, cell centroids and volumes, ghost-cell centres (the
// reference does not recompute the interface orientations there); the WLS tables follow through
// wls_r_body / cell_static_body.  Same formulas as grid_bodies.h (the initial build).
// ---------------------------------------------------------------------------------------
struct GeomUpdate {
    const int* fverts;     // [4][fstride] vertex ids of every interface, -1 padded
    const int* cverts;     // [8][cstride] vertex ids of every valid cell, VTK order
    const int* ctype;      // VTK code of every valid cell
    const double* pos;     // n_vertices x 3
    double* f_c;           // the GridDev arrays, writable
    double* f_area;
    double* f_geom;
    double* cxyz;          // cx | cy | cz
    double* volume;
};

IBIS_HD void geom_face_body(const GridDev& g, const GeomUpdate& U, int f) {
    long v[4];
    int nv = 0;
    for (int k = 0; k < 4; k++) {
        v[k] = U.fverts[k * g.fstride + f];
        if (v[k] >= 0) nv = k + 1;
    }
    for (int d = 0; d < 3; d++) {
        double acc = 0.0;
        for (int k = 0; k < nv; k++) acc += U.pos[3 * v[k] + d];
        U.f_c[d * g.fstride + f] = acc / (double)nv;
    }
    double area;
    if (nv == 2) {
        const double d0 = U.pos[3 * v[1]] - U.pos[3 * v[0]], d1 = U.pos[3 * v[1] + 1] - U.pos[3 * v[0] + 1],
                     d2 = U.pos[3 * v[1] + 2] - U.pos[3 * v[0] + 2];
        area = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    } else if (nv == 3) {
        area = gb_tri_area(U.pos, v[0], v[1], v[2]);
    } else {
        area = gb_tri_area(U.pos, v[0], v[1], v[2]) + gb_tri_area(U.pos, v[0], v[2], v[3]);
    }
    U.f_area[f] = area;
    U.f_geom[(long)f * 10 + 9] = area;
}

IBIS_HD void geom_centroid_body(const GridDev& g, const GeomUpdate& U, int c) {
    const int nv = gb_num_vertices(U.ctype[c]);
    for (int d = 0; d < 3; d++) {
        double acc = 0.0;
        for (int k = 0; k < nv; k++) acc += U.pos[3 * (long)U.cverts[k * g.cstride + c] + d];
        U.cxyz[d * g.cstride + c] = acc / (double)nv;
    }
}

IBIS_HD void geom_volume_body(const GridDev& g, const GeomUpdate& U, int c) {
    const long cs = g.cstride, fs_ = g.fstride;
    if (g.dim == 2) {
        const long a = U.cverts[c], b = U.cverts[cs + c], d = U.cverts[2 * cs + c];
        U.volume[c] = U.ctype[c] == VTK_TRI ? gb_tri_area(U.pos, a, b, d)
                                           : gb_tri_area(U.pos, a, b, d) + gb_tri_area(U.pos, a, d, (long)U.cverts[3 * cs + c]);
        return;
    }
    double acc = 0.0;
    for (int j = 0; j < g.max_faces; j++) {
        const int f = g.cf_face[j * cs + c];
        if (f < 0) break;
        const double out = (double)g.cf_out[j * cs + c];
        const double dotv = U.f_c[f] * out * g.f_n[f] + U.f_c[fs_ + f] * out * g.f_n[fs_ + f] +
                            U.f_c[2 * fs_ + f] * out * g.f_n[2 * fs_ + f];
        acc = acc + dotv * U.f_area[f];
    }
    U.volume[c] = acc / 3.0;
}

IBIS_HD void geom_ghost_body(const GridDev& g, const GeomUpdate& U, int k) {
    const int f = g.bf_face[k];
    const int l = g.f_left[f], r = g.f_right[f];
    int valid, ghost;
    if (l < g.n_cells) { valid = l; ghost = r; } else { valid = r; ghost = l; }
    for (int d = 0; d < 3; d++) {
        const double fc = U.f_c[d * g.fstride + f];
        const double dxyz = fc - U.cxyz[d * g.cstride + valid];
        U.cxyz[d * g.cstride + ghost] = fc + dxyz;
    }
}

// GridBlock::compute_face_vel (grid.h:571-592): average of the vertex velocities of the interface, summed in
// the interface's vertex order.  fverts = [4][fstride] vertex ids, -1 past the last; vvel = n_vertices x 3.
IBIS_HD void face_vel_body(const GridDev& g, const int* fverts, const double* vvel, double* face_vel, int f) {
    double vx = 0.0, vy = 0.0, vz = 0.0;
    int nv = 0;
    for (int k = 0; k < 4; k++) {
        const int v = fverts[k * g.fstride + f];
        if (v < 0) break;
        vx += vvel[3 * (long)v];
        vy += vvel[3 * (long)v + 1];
        vz += vvel[3 * (long)v + 2];
        nv++;
    }
    face_vel[f] = vx / (double)nv;
    face_vel[g.fstride + f] = vy / (double)nv;
    face_vel[2 * g.fstride + f] = vz / (double)nv;
}

// FiniteVolume::apply_geometric_conservation_law (finite_volume.cpp:99-128): dudt -= U dV/dt / V
template <int DIM, typename S>
IBIS_HD void gcl_body(const FvArgs<S>& A, S* const* cq, int c) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    const long cs = g.cstride, fs_ = g.fstride;
    S dVdt = S(0.0);
    for (int j = 0; j < MAXF; j++) {
        if (j >= g.max_faces) break;
        const int f = g.cf_face[j * cs + c];
        if (f < 0) break;
        const double area = g.f_area[f] * (double)g.cf_out[j * cs + c];
        const double vx = A.face_vel[f], vy = A.face_vel[fs_ + f], vz = A.face_vel[2 * fs_ + f];
        dVdt += (vx * g.f_n[f] + vy * g.f_n[fs_ + f] + vz * g.f_n[2 * fs_ + f]) * area;
    }
    const S V = S(g.volume[c]);
    for (int k = 0; k < 5; k++) {
        if (k == 3 && DIM != 3) continue;
        M::store(A.dudt[k], c, M::load(A.dudt[k], c) - M::load(cq[k], c) * dVdt / V);
    }
}

// flux_surface_integral from the stored face fluxes (finite_volume.cpp:221-256)
template <int DIM, typename S>
IBIS_HD void integral_body(const FvArgs<S>& A, int c) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    const long cs = g.cstride, fs_ = g.fstride;
    S d_mass = S(0.0), d_mx = S(0.0), d_my = S(0.0), d_mz = S(0.0), d_e = S(0.0);
    for (int j = 0; j < MAXF; j++) {
        if (j >= g.max_faces) break;
        const int f = g.cf_face[j * cs + c];
        if (f < 0) break;
        const S area = -S(g.f_area[f]) * (double)g.cf_out[j * cs + c];
        d_mass += M::load(A.flux, 0 * fs_ + f) * area;
        d_mx += M::load(A.flux, 1 * fs_ + f) * area;
        d_my += M::load(A.flux, 2 * fs_ + f) * area;
        if (DIM == 3) d_mz += M::load(A.flux, 3 * fs_ + f) * area;
        d_e += M::load(A.flux, 4 * fs_ + f) * area;
    }
    const S vol = S(g.volume[c]);
    M::store(A.dudt[0], c, d_mass / vol);
    M::store(A.dudt[1], c, d_mx / vol);
    M::store(A.dudt[2], c, d_my / vol);
    if (DIM == 3) M::store(A.dudt[3], c, d_mz / vol);
    M::store(A.dudt[4], c, d_e / vol);
}

// ---------------------------------------------------------------------------------------
// pass F, tiled (inviscid): one thread block owns a tile of consecutive cells and stages the
// face stencil in shared memory.  st = [2 sides][6][cap] of S.
//   phase 1a  one thread per CELL: its own values, gradients and limiters (coalesced loads, no
//             gather) are extrapolated to each of its faces (MUSCL, convective_flux.cpp:258-420)
//             and the state (rho p T u v w) is written to that face's left or right slot;
//   phase 1b  the sides that belong to cells outside the tile (neighbouring tiles, halo,
//             boundary ghosts) are reconstructed from gathered data;
//   phase 2   one thread per INTERFACE: rotate, flux calculator, rotate back -- each flux ONCE
//             instead of once per adjacent cell; the flux replaces the left state in place;
//   phase 3   one thread per cell: flux_surface_integral (finite_volume.cpp:221-256) from shared
//             memory in cell-face order: owner computes, no atomics, no face arrays in HBM.
// Same per-face arithmetic and the same in-cell summation order as the reference.
// ---------------------------------------------------------------------------------------
template <typename S>
IBIS_HD void tile_store_side(S* st, int cap, int side, int slot, const Side<S>& s) {
    S* p = st + (size_t)side * 6 * cap + slot;
    p[0 * cap] = s.rho;
    p[1 * cap] = s.p;
    p[2 * cap] = s.T;
    p[3 * cap] = s.u;
    p[4 * cap] = s.v;
    p[5 * cap] = s.w;
}

// phase 1a: cell c (a cell of the tile) -> one side of each of its faces.  Every load is
// unconditional (the slot tables are MAXF wide) so that all of them can be in flight together;
// only the shared-memory store is predicated on the slot being present.
template <int DIM, bool O2, typename S>
IBIS_HD void tile_recon_own_body(const FvArgs<S>& A, int c, S* st) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    const long cs = g.cstride;
    const int cap = g.tile_cap;
    int slot[MAXF], side[MAXF];
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        slot[j] = g.cf_slot[j * cs + c];
        side[j] = g.cf_out[j * cs + c] > 0 ? 0 : 1;
    }
    if (!O2) {                                                // copy_reconstruct :212-240
        Side<S> s;
        s.T = M::ldg(A.fs[F_T], c);
        s.p = M::ldg(A.fs[F_P], c);
        s.rho = M::ldg(A.fs[F_RHO], c);
        s.e = S(0.0);
        s.u = M::ldg(A.fs[F_VX], c);
        s.v = M::ldg(A.fs[F_VY], c);
        s.w = M::ldg(A.fs[F_VZ], c);
#pragma unroll
        for (int j = 0; j < MAXF; j++)
            if (slot[j] != 0xFFFF) tile_store_side(st, cap, side[j], slot[j], s);
        return;
    }
    S val[5], gx[5], gy[5], gz[5], lim[5];
#pragma unroll
    for (int v = 0; v < 5; v++) {
        val[v] = M::ldg(A.rc.val[v], c);
        if (DIM == 2 && v == 4) continue;                     // vz: no gradient in 2-D
        gx[v] = M::ldg(A.rc.g[v][0], c);
        gy[v] = M::ldg(A.rc.g[v][1], c);
        gz[v] = (DIM == 3) ? M::ldg(A.rc.g[v][2], c) : S(0.0);
        lim[v] = A.limiter_on ? M::ldg(A.rc.lim[v], c) : S(1.0);
    }
    double d[MAXF][3];
#pragma unroll
    for (int j = 0; j < MAXF; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) d[j][k] = g.cf_d[(j * 3 + k) * cs + c];
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        const double dx = d[j][0], dy = d[j][1], dz = d[j][2];
        S r[5];
#pragma unroll
        for (int v = 0; v < 5; v++) {
            if (DIM == 2 && v == 4) { r[v] = val[v]; continue; }
            r[v] = val[v] + lim[v] * (gx[v] * dx + gy[v] * dy + gz[v] * dz);   // linear_interpolate :243-255
        }
        Side<S> s;
        s.p = S(0.0);
        s.T = S(0.0);
        s.e = S(0.0);
        s.rho = r[0];
        if (A.thermo == THERMO_RHO_P) s.p = r[1];
        else if (A.thermo == THERMO_RHO_T) s.T = r[1];
        else s.e = r[1];
        thermo_close(A.thermo, A.gas, s.rho, s.p, s.T, s.e);
        s.u = r[2];
        s.v = r[3];
        s.w = r[4];
        if (slot[j] != 0xFFFF) tile_store_side(st, cap, side[j], slot[j], s);
    }
}

IBIS_HD int tile_num_slots(const GridDev& g, int tile) {
    return (g.home0[tile + 1] - g.home0[tile]) + (g.foreign_off[tile + 1] - g.foreign_off[tile]);
}
IBIS_HD int tile_face_of_slot(const GridDev& g, int tile, int slot) {
    const int h0 = g.home0[tile], nh = g.home0[tile + 1] - h0;
    return slot < nh ? h0 + slot : g.foreign[g.foreign_off[tile] + (slot - nh)];
}

// phase 1b: entry e of the tile's outside-side list (gathers; the offset is static)
template <int DIM, bool O2, typename S>
IBIS_HD void tile_recon_outside_body(const FvArgs<S>& A, int e, S* st) {
    const GridDev& g = A.g;
    const int packed = g.out_slot[e];
    const int slot = packed & 0x7FFF, side = packed >> 15;
    const int cell = g.out_cell[e];
    double dx = 0.0, dy = 0.0, dz = 0.0;
    if (O2) {
        dx = g.out_d[e];
        dy = g.out_d[g.n_out + e];
        dz = g.out_d[2 * g.n_out + e];
    }
    const Side<S> s = reconstruct_at<DIM, O2>(A, cell, dx, dy, dz);
    tile_store_side(st, g.tile_cap, side, slot, s);
}

// phase 2: interface in slot `slot`: st[0][0..4][slot] <- flux (global frame), st[0][5][slot] <- area
template <int DIM, int FLUX, typename S>
IBIS_HD void tile_flux_body(const FvArgs<S>& A, int tile, int slot, S* st) {
    const GridDev& g = A.g;
    const long fs_ = g.fstride;
    const int cap = g.tile_cap;
    const int f = tile_face_of_slot(g, tile, slot);
    Side<S> L, R_;
    {
        const S* p = st + slot;
        L.rho = p[0 * cap]; L.p = p[1 * cap]; L.T = p[2 * cap]; L.u = p[3 * cap]; L.v = p[4 * cap]; L.w = p[5 * cap];
        p += (size_t)6 * cap;
        R_.rho = p[0 * cap]; R_.p = p[1 * cap]; R_.T = p[2 * cap]; R_.u = p[3 * cap]; R_.v = p[4 * cap]; R_.w = p[5 * cap];
        L.e = S(0.0);
        R_.e = S(0.0);
    }
    const double nx = g.f_n[f], ny = g.f_n[fs_ + f], nz = g.f_n[2 * fs_ + f];
    const double t1x = g.f_t1[f], t1y = g.f_t1[fs_ + f], t1z = g.f_t1[2 * fs_ + f];
    const double t2x = g.f_t2[f], t2y = g.f_t2[fs_ + f], t2z = g.f_t2[2 * fs_ + f];
    {
        const S x = L.u * nx + L.v * ny + L.w * nz;           // util/vector3.cpp:142-153
        const S y = L.u * t1x + L.v * t1y + L.w * t1z;
        const S z = L.u * t2x + L.v * t2y + L.w * t2z;
        L.u = x; L.v = y; L.w = z;
    }
    {
        const S x = R_.u * nx + R_.v * ny + R_.w * nz;
        const S y = R_.u * t1x + R_.v * t1y + R_.w * t1z;
        const S z = R_.u * t2x + R_.v * t2y + R_.w * t2z;
        R_.u = x; R_.v = y; R_.w = z;
    }
    Flux5<S> F = flux_calc<FLUX>(L, R_, A.gas, A.delta);
    const S px = F.mx, py = F.my;
    const S pz = (DIM == 3) ? F.mz : S(0.0);
    S* q = st + slot;
    q[0 * cap] = F.mass;                                      // convective_flux.cpp:148-164
    q[1 * cap] = px * nx + py * t1x + pz * t2x;
    q[2 * cap] = px * ny + py * t1y + pz * t2y;
    q[3 * cap] = px * nz + py * t1z + pz * t2z;
    q[4 * cap] = F.energy;
    q[5 * cap] = S(g.f_area[f]);
}

// phase 3
// what phase 3 needs of a cell besides the staged fluxes; loaded ahead of the flux phase
struct CellFaces {
    int slot[MAXF];
    double out[MAXF];
    double vol;
};
IBIS_HD void tile_cell_load(const GridDev& g, int c, CellFaces& F) {
    const long cs = g.cstride;
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        F.slot[j] = g.cf_slot[j * cs + c];
        F.out[j] = (double)g.cf_out[j * cs + c];
    }
    F.vol = g.volume[c];
}
// st[slot * ss + k * ks]: flux component k (5 = area) of the tile's interface `slot`
template <int DIM, typename S>
IBIS_HD void tile_cell_sum(const FvArgs<S>& A, int c, const CellFaces& F, const S* st, long ss, long ks) {
    typedef mem<S> M;
    const S vol = S(F.vol);
    S d_mass = S(0.0), d_mx = S(0.0), d_my = S(0.0), d_mz = S(0.0), d_e = S(0.0);
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        if (F.slot[j] == 0xFFFF) continue;
        const S* q = st + F.slot[j] * ss;
        const double area = -re(q[5 * ks]) * F.out[j];
        d_mass += q[0 * ks] * area;
        d_mx += q[1 * ks] * area;
        d_my += q[2 * ks] * area;
        if (DIM == 3) d_mz += q[3 * ks] * area;
        d_e += q[4 * ks] * area;
    }
    // division by the volume stays a division by S (dual / dual rounds its dual part
    // differently from dual / double, util/dual.h:169-180)
    M::store(A.dudt[0], c, d_mass / vol);
    M::store(A.dudt[1], c, d_mx / vol);
    M::store(A.dudt[2], c, d_my / vol);
    if (DIM == 3) M::store(A.dudt[3], c, d_mz / vol);
    M::store(A.dudt[4], c, d_e / vol);
}
template <int DIM, typename S>
IBIS_HD void tile_cell_body(const FvArgs<S>& A, int c, const S* st, long ss, long ks) {
    CellFaces F;
    tile_cell_load(A.g, c, F);
    tile_cell_sum<DIM>(A, c, F, st, ss, ks);
}

// ---------------------------------------------------------------------------------------
// The "sides" path of the inviscid second-order residual (the default on the GPU):
//   pass G'  grad_sides_body   one thread per cell: WLS gradients (grid/gradient.h:66-116),
//            Barth-Jespersen limiters (limiter.cpp:24-63) AND the MUSCL extrapolation of the cell to
//            each of its interfaces with the thermodynamic closure (convective_flux.cpp:243-420,
//            gas_model.h:91-117), written as one 48-byte record (rho p T vx vy vz) per interface
//            side.  The limiter's delta_2 = grad . (x_f - x_c) IS the reconstruction increment, so
//            it is evaluated once.  Gradients and limiters never go to HBM.
//   pass F'  flux_from_sides + tile_cell_sum   one thread per interface of a tile: rotate,
//            flux calculator, rotate back, staged in shared memory; then one thread per cell:
//            flux_surface_integral (finite_volume.cpp:221-256).  No gathers of cell data at all.
// Same per-face arithmetic and in-cell summation order as the reference, hence the same bits.
// ---------------------------------------------------------------------------------------

// Barth-Jespersen limiter of one variable from the face increments delta[j] (0 for an absent
// slot): the reference's loop, one division per face (limiter.cpp:41-60).
template <int NF, typename S>
IBIS_HD S bj_phi(const S* delta, const S& ui, const S& umin, const S& umax, double eps) {
    S phi = S(1.0);
#pragma unroll
    for (int j = 0; j < NF; j++) {
        S delta_2 = delta[j];
        const int sgn = (re(delta_2) > 0.0) - (re(delta_2) < 0.0);
        delta_2 = (double)sgn * (abs_(delta_2) + eps);
        if (sgn > 0) {
            phi = min_(phi, (umax - ui) / delta_2);
        } else if (sgn < 0) {
            phi = min_(phi, (umin - ui) / delta_2);
        }
    }
    return phi;
}
// Real numbers: two divisions instead of NF, bit-identical (see bj_limiters above).
template <int NF>
IBIS_HD double bj_phi(const double* delta, const double& ui, const double& umin, const double& umax, double eps) {
    double dmax = 0.0, dmin = 0.0;
#pragma unroll
    for (int j = 0; j < NF; j++) {
        dmax = dmax < delta[j] ? delta[j] : dmax;
        dmin = delta[j] < dmin ? delta[j] : dmin;
    }
    double p = 1.0;
    if (dmax > 0.0) p = min_(p, div_z(umax - ui, dmax + eps));
    if (dmin < 0.0) p = min_(p, div_z(umin - ui, -(-dmin + eps)));
    return p;
}

// The static per cell-face tables (WLS weights cf_w, offsets cf_d) as the kernel sees them:
// element (j, k) of cell `idx` is base[(j*3 + k) * stride + idx].  On the device the tile's slice
// of both tables is staged in shared memory by bulk async copies (stride = tile size, idx = cell
// within the tile); the host emulation points at the global tables (stride = cstride, idx = cell).
template <int WSTRIDE>
struct CellTables {
    const double* w;
    const double* d;
    long stride_rt;
    int idx;
    // staged tables are re-read from shared memory at every use (volatile): keeping all 36
    // entries in registers across the variables is what made the first version spill
    IBIS_HD double W(int j, int k) const {
        if (WSTRIDE > 0) return *static_cast<const volatile double*>(w + (j * 3 + k) * WSTRIDE + idx);
        return w[(j * 3 + k) * stride_rt + idx];
    }
    IBIS_HD double D(int j, int k) const {
        if (WSTRIDE > 0) return *static_cast<const volatile double*>(d + (j * 3 + k) * WSTRIDE + idx);
        return d[(j * 3 + k) * stride_rt + idx];
    }
};

// one variable of one cell: gradient, limiter, extrapolation to the NF interfaces -> r[j]
// (TT: the static tables of the cell, W(j, k) WLS weights and D(j, k) face offsets)
template <int DIM, int NF, typename TT, typename S>
IBIS_HD void var_to_faces(const S& ui, const S* un, const TT& T, double eps, bool limit, S* r) {
    S gx = S(0.0), gy = S(0.0), gz = S(0.0), umin = ui, umax = ui;
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const S diff_u = un[j] - ui;                           // gradient.h:110-113
        gx += T.W(j, 0) * diff_u;
        gy += T.W(j, 1) * diff_u;
        gz += ((DIM == 3) ? T.W(j, 2) : 0.0) * diff_u;
        umin = min_(umin, un[j]);                              // limiter.cpp:37-38
        umax = max_(umax, un[j]);
    }
    S delta[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) delta[j] = gx * T.D(j, 0) + gy * T.D(j, 1) + gz * T.D(j, 2);
    const S phi = limit ? bj_phi<NF>(delta, ui, umin, umax, eps) : S(1.0);
#pragma unroll
    for (int j = 0; j < NF; j++) r[j] = ui + phi * delta[j];   // linear_interpolate, convective_flux.cpp:243-255
}

// ---- pass G' on the packed per-tile records (GridDev::t_wd / t_grec) ----
// entry of the tile-local neighbour table
constexpr int NBR_LOCAL_MASK = 0x3FFF, NBR_LEFT_BIT = 0x4000, NBR_GHOST_BIT = 0x8000;
// a view of one tile's record in memory (global memory in the host emulation, shared memory on the device)
struct GRec {
    const unsigned short* nbr;     // [nf][T]
    const unsigned short* stage;   // [nf][T]
    const int* hdr;                // n_halo, nR, n_own_l, n_stage, has_ghost (a cell of the tile has a boundary ghost neighbour), home
    const int* halo;               // [halo_cap]
};
IBIS_HD GRec grec_view(const unsigned char* rec, int nf, int T) {
    GRec R;
    R.nbr = reinterpret_cast<const unsigned short*>(rec);
    R.stage = R.nbr + (size_t)nf * T;
    R.hdr = reinterpret_cast<const int*>(R.stage + (size_t)nf * T);
    R.halo = R.hdr + 8;
    return R;
}
IBIS_HD int grec_bytes(int nf, int T, int halo_cap) { return 4 * nf * T + 32 + 4 * halo_cap; }

// A side record = what linear_reconstruct extrapolates to one side of an interface (convective_flux.cpp:243-420):
// rho, the second thermodynamic variable of the interpolator (p, T, or T = u / Cv: side_second below), vx, vy, vz.
// 40 bytes; the rest of the thermodynamic closure (gas_model.h:91-117) is evaluated by the consumer, pass F'.
constexpr int SIDE_REC = 5;
constexpr int SIDE_EVEN = 2;       // first records and counts of bulk transfers are multiples of this
// The thermodynamic closure of a side (gas_model.h:91-117) is split so that the consumer has no division left to do
// where that is possible.  The producer (pass G') stores, next to rho, what side_second returns for the second
// reconstructed variable: T = u / Cv for rho_u (update_thermo_from_rhou's first line), the variable itself for
// rho_T and rho_p.  The consumer (pass F') finishes with side_close: p = rho R T for rho_u and rho_T,
// T = p / (rho R) for rho_p.  Same calls on the same values as thermo_close, hence the same bits.
template <typename S>
IBIS_HD S side_second(int thermo, const Gas<S>& gas, const S& th2) {
    return thermo == THERMO_RHO_U ? th2 / gas.Cv : th2;
}
template <typename S>
IBIS_HD void side_close(int thermo, const Gas<S>& gas, const S& rho_in, const S& second, S& rho, S& p, S& Tj) {
    rho = rho_in;
    if (thermo == THERMO_RHO_P) {
        p = second;
        Tj = T_from_rhop(rho, p, gas.R);
    } else {
        Tj = second;
        p = p_from_rhoT(rho, Tj, gas.R);
    }
}
// staging index of the far side of a boundary interface (the ghost cell's own values): the far side
// sits at a fixed distance from the near side in the tile's staging buffer (layout.h)
IBIS_HD int far_stage_index(bool left, int stage, int nR, int n_own_l) {
    return left ? n_own_l + stage + nR : stage - n_own_l - nR;
}

// all five variables of tile-local cell i, one after the other: the host emulation of pass G' (the device kernel
// gives every (cell, variable) its own thread, fv_pipe.cu, with the same calls).  sv = the tile's state values
// [5][svs]: own cells at [0, T), halo-list entries from T on.  wd = the tile's slice of t_wd.
template <int DIM, int NF, typename S>
IBIS_HD void pipe_grad_cell(const FvArgs<S>& A, const double* wd, const GRec& R, int T, int i, const S* sv, int svs, S* stg) {
    const bool limit = A.limiter_on != 0;
    CellTables<0> Tb;
    Tb.w = wd;
    Tb.d = wd + (size_t)NF * 3 * T;
    Tb.stride_rt = T;
    Tb.idx = i;
    const int nR = R.hdr[1], n_own_l = R.hdr[2];
    int nl[NF], st[NF];
    bool left[NF], ghost[NF];
#pragma unroll
    for (int j = 0; j < NF; j++) {
        const int e = R.nbr[j * T + i];
        nl[j] = e & NBR_LOCAL_MASK;
        left[j] = (e & NBR_LEFT_BIT) != 0;
        ghost[j] = (e & NBR_GHOST_BIT) != 0;
        st[j] = R.stage[j * T + i];
    }
    S r[5][NF];
#pragma unroll
    for (int v = 0; v < 5; v++) {
        const S ui = sv[v * svs + i];
        if (DIM == 2 && v == 4) {
            // 2-D: vz has no gradient (convective_flux.cpp:206-208) and a unit limiter: value copied
#pragma unroll
            for (int j = 0; j < NF; j++) r[v][j] = ui;
            continue;
        }
        S un[NF];
#pragma unroll
        for (int j = 0; j < NF; j++) un[j] = sv[v * svs + nl[j]];
        var_to_faces<DIM, NF>(ui, un, Tb, A.eps, limit, r[v]);
    }
#pragma unroll
    for (int j = 0; j < NF; j++) {
        if (st[j] == 0xFFFF) continue;
        S* q = stg + (long)st[j] * SIDE_REC;
#pragma unroll
        for (int v = 0; v < 5; v++) q[v] = v == 1 ? side_second(A.thermo, A.gas, r[v][j]) : r[v][j];
        if (ghost[j]) {
            // the far side of a boundary interface is the ghost cell itself: zero gradient, unit limiter
            // (convective_flux.cpp:281-300), i.e. its values, closed thermodynamically like any other side
            S* f = stg + (long)far_stage_index(left[j], st[j], nR, n_own_l) * SIDE_REC;
#pragma unroll
            for (int v = 0; v < 5; v++) f[v] = v == 1 ? side_second(A.thermo, A.gas, sv[v * svs + nl[j]]) : sv[v * svs + nl[j]];
        }
    }
}

// The bulk transfers of a tile (layout.h).  Transfer q of tile b moves `count` interfaces starting
// at interface f0; in the staging buffer of pass G' they start at `stage`, in the flux slots of pass F'
// at `slot`.  Stores (pass G'): q = 0 home left range, 1 home right range, 2 + e entry e.
struct TileXfer {
    int f0, count, stage, slot;
    bool left;                    // stores: which array the tile's own sides go to
};
IBIS_HD int tile_n_own_left(const GridDev& g, int b) { return g.tile_info[TILE_INFO * b + 8]; }
// entry e of the tile whose descriptor is ti: inline in the descriptor when there are few
IBIS_HD const int* tile_entry(const GridDev& g, const int* ti, int e) {
    return (ti[7] <= TILE_INLINE_ENT ? ti + 12 : g.tile_ent + 5 * ti[6]) + 5 * e;
}
// first cell and cell count of a tile without touching memory: owned tiles are runs of `tile`
// cells from 0, the tiles of the inner halo ring runs from n_owned
IBIS_HD void tile_cells(const GridDev& g, int b, int& c0, int& nc) {
    const bool owned = b < g.n_tiles;
    c0 = owned ? b * g.tile : g.n_owned + (b - g.n_tiles) * g.tile;
    const int end = owned ? g.n_owned : g.n_grad;
    nc = end - c0 < g.tile ? end - c0 : g.tile;
}
IBIS_HD TileXfer tile_store_xfer(const GridDev& g, int b, int q, int n_own_l) {
    const int* ti = g.tile_info + TILE_INFO * b;
    TileXfer x;
    if (q == 0) { x.f0 = ti[2] + ti[4]; x.count = ti[3] - ti[4]; x.stage = 0; x.slot = ti[4]; x.left = true; }
    else if (q == 1) { x.f0 = ti[2]; x.count = ti[4] + ti[5]; x.stage = n_own_l; x.slot = 0; x.left = false; }
    else {
        const int* te = tile_entry(g, ti, q - 2);
        x.f0 = te[0]; x.count = te[1]; x.stage = te[4]; x.slot = te[3]; x.left = te[2] == 0;
    }
    return x;
}
// loads (pass F'): q = 0 the home range, 1 + e entry e; each moves left sides, right sides and frames
IBIS_HD TileXfer tile_load_xfer(const GridDev& g, int b, int q) {
    const int* ti = g.tile_info + TILE_INFO * b;
    TileXfer x;
    x.stage = 0;
    x.left = true;
    if (q == 0) { x.f0 = ti[2]; x.count = ti[3]; x.slot = 0; }
    else {
        const int* te = tile_entry(g, ti, q - 1);
        x.f0 = te[0]; x.count = te[1]; x.slot = te[3];
    }
    return x;
}

// pass F', one interface: Ls, Rs = its side states, fr = its frame t1 t2 (shared memory on the device, the
// global arrays in the host emulation); out[0..4] <- flux in the global frame.  out may alias Ls: everything is
// read before anything is written.  The normal is cross(t1, t2) as the reference computes it
// (grid/interface.h:250, util/vector3.h:153-160).
IBIS_HD void frame_normal(const double* fr, double& nx, double& ny, double& nz) {
    nx = fr[1] * fr[5] - fr[2] * fr[4];
    ny = fr[2] * fr[3] - fr[0] * fr[5];
    nz = fr[0] * fr[4] - fr[1] * fr[3];
}
// the operands of one interface: side states and frame, as loaded from the stage
template <typename S>
struct FluxOperands {
    Side<S> L, R;
    double t1x, t1y, t1z, t2x, t2y, t2z;
};
template <typename S>
IBIS_HD void flux_load_operands(const FvArgs<S>& A, const S* Ls, const S* Rs, const double* fr, FluxOperands<S>& O) {
    typedef mem<S> M;
    // 40-byte records: 8-byte alignment only
    const S la = M::load(Ls, 0), lb = M::load(Ls, 1), ra = M::load(Rs, 0), rb = M::load(Rs, 1);
    O.L.u = M::load(Ls, 2); O.L.v = M::load(Ls, 3); O.L.w = M::load(Ls, 4);
    O.R.u = M::load(Rs, 2); O.R.v = M::load(Rs, 3); O.R.w = M::load(Rs, 4);
    side_close(A.thermo, A.gas, la, lb, O.L.rho, O.L.p, O.L.T);
    side_close(A.thermo, A.gas, ra, rb, O.R.rho, O.R.p, O.R.T);
    O.L.e = S(0.0);
    O.R.e = S(0.0);
    mem<double>::ld2(fr, 0, O.t1x, O.t1y);
    mem<double>::ld2(fr, 2, O.t1z, O.t2x);
    mem<double>::ld2(fr, 4, O.t2y, O.t2z);
}
template <int DIM, int FLUX, typename S>
IBIS_HD void flux_from_operands(const FvArgs<S>& A, const FluxOperands<S>& O, S* out) {
    Side<S> L = O.L, R_ = O.R;
    const double t1x = O.t1x, t1y = O.t1y, t1z = O.t1z, t2x = O.t2x, t2y = O.t2y, t2z = O.t2z;
    const double nx = t1y * t2z - t1z * t2y;
    const double ny = t1z * t2x - t1x * t2z;
    const double nz = t1x * t2y - t1y * t2x;
    {
        const S x = L.u * nx + L.v * ny + L.w * nz;           // util/vector3.cpp:142-153
        const S y = L.u * t1x + L.v * t1y + L.w * t1z;
        const S z = L.u * t2x + L.v * t2y + L.w * t2z;
        L.u = x; L.v = y; L.w = z;
    }
    {
        const S x = R_.u * nx + R_.v * ny + R_.w * nz;
        const S y = R_.u * t1x + R_.v * t1y + R_.w * t1z;
        const S z = R_.u * t2x + R_.v * t2y + R_.w * t2z;
        R_.u = x; R_.v = y; R_.w = z;
    }
    Flux5<S> F = flux_calc<FLUX>(L, R_, A.gas, A.delta);
    const S px = F.mx, py = F.my;
    const S pz = (DIM == 3) ? F.mz : S(0.0);
    out[0] = F.mass;                                          // convective_flux.cpp:148-164
    out[1] = px * nx + py * t1x + pz * t2x;
    out[2] = px * ny + py * t1y + pz * t2y;
    out[3] = px * nz + py * t1z + pz * t2z;
    out[4] = F.energy;
}
template <int DIM, int FLUX, typename S>
IBIS_HD void flux_from_sides(const FvArgs<S>& A, const S* Ls, const S* Rs, const double* fr, S* out) {
    FluxOperands<S> O;
    flux_load_operands(A, Ls, Rs, fr, O);
    flux_from_operands<DIM, FLUX>(A, O, out);
}

// the per-tile interface record of pass F' (GridDev::t_freca)
struct FRecA {
    const unsigned short* lidx;    // [cap]
    const unsigned short* ridx;    // [cap]
    const double* frame;           // [cap][6]
};
IBIS_HD FRecA freca_view(const unsigned char* rec, int cap) {
    FRecA R;
    R.lidx = reinterpret_cast<const unsigned short*>(rec);
    R.ridx = R.lidx + cap;
    R.frame = reinterpret_cast<const double*>(R.ridx + cap);
    return R;
}
IBIS_HD int freca_bytes(int cap) { return 4 * cap + 48 * cap; }

// the per-tile cell record of pass F' (GridDev::t_frec)
struct FRec {
    const double* vol;             // [T]
    const double* sarea;           // [6][T]  -area * outsign of the cell's face j
    const unsigned short* slot;    // [6][T]  flux slot of the cell's face j in the tile, 0xFFFF past the last face
};
IBIS_HD FRec frec_view(const unsigned char* rec, int T) {
    FRec R;
    R.vol = reinterpret_cast<const double*>(rec);
    R.sarea = R.vol + T;
    R.slot = reinterpret_cast<const unsigned short*>(R.sarea + (size_t)MAXF * T);
    return R;
}
IBIS_HD int frec_bytes(int T) { return T * (8 + 8 * MAXF + 2 * MAXF); }

// flux_surface_integral of tile-local cell i (global cell c) from the staged fluxes st[slot * QS + k]
// (finite_volume.cpp:221-256): same products, same order over the cell's faces as tile_cell_sum
template <int DIM, int QS = 6, typename S>
IBIS_HD void pipe_cell_sum(const FvArgs<S>& A, int c, int i, int T, const FRec& R, const S* st) {
    typedef mem<S> M;
    S d_mass = S(0.0), d_mx = S(0.0), d_my = S(0.0), d_mz = S(0.0), d_e = S(0.0);
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        const int sl = R.slot[j * T + i];
        if (sl == 0xFFFF) continue;
        const S* q = st + (long)sl * QS;
        const double area = R.sarea[j * T + i];
        d_mass += q[0] * area;
        d_mx += q[1] * area;
        d_my += q[2] * area;
        if (DIM == 3) d_mz += q[3] * area;
        d_e += q[4] * area;
    }
    const S vol = S(R.vol[i]);
    M::store(A.dudt[0], c, d_mass / vol);
    M::store(A.dudt[1], c, d_mx / vol);
    M::store(A.dudt[2], c, d_my / vol);
    if (DIM == 3) M::store(A.dudt[3], c, d_mz / vol);
    M::store(A.dudt[4], c, d_e / vol);
}
// local time step of the cell (finite_volume.cpp:145-180, inviscid) from the tile's frames[slot * 6 + k]:
// the same normals and areas, in the same cell-face order, as local_dt_body reads
IBIS_HD double pipe_cell_dt(double vx, double vy, double vz, double Tc, double gas_R, double gamma, double signal_factor,
                            int i, int T, const FRec& R, const double* frames) {
    const double a = speed_of_sound(Tc, gas_R, gamma);
    double sr_c = 0.0;
    const double sr_v = 0.0;
#pragma unroll
    for (int j = 0; j < MAXF; j++) {
        const int sl = R.slot[j * T + i];
        if (sl == 0xFFFF) continue;
        double nx, ny, nz;
        frame_normal(frames + (long)sl * 6, nx, ny, nz);
        const double area = fabs(R.sarea[j * T + i]);
        const double dot = vx * nx + vy * ny + vz * nz;
        const double sig_vel = fabs(dot) + a;
        sr_c += sig_vel * area;
    }
    return R.vol[i] / (sr_c + signal_factor * sr_v);
}

// ---------------------------------------------------------------------------------------
// estimate_dt: local time step of one cell (finite_volume.cpp:145-180)
// ---------------------------------------------------------------------------------------
// The reduction takes real_part(local_dt); real parts of a dual evaluation equal the double
// evaluation, so this runs in double on the real parts for both scalar types.
template <typename S>
IBIS_HD double local_dt_body(const FvArgs<S>& A, int c) {
    typedef mem<S> M;
    const GridDev& g = A.g;
    const long cs = g.cstride, fs_ = g.fstride;
    const double vx = re(M::load(A.fs[F_VX], c)), vy = re(M::load(A.fs[F_VY], c)),
                 vz = re(M::load(A.fs[F_VZ], c));
    double sr_c = 0.0, sr_v = 0.0;
    const double volume = g.volume[c];
    const double R = re(A.gas.R), gamma = re(A.gas.gamma), Cp = re(A.gas.Cp);
    const double a = speed_of_sound(re(M::load(A.fs[F_T], c)), R, gamma);
    for (int j = 0; j < MAXF; j++) {
        if (j >= g.max_faces) break;
        const int f = g.cf_face[j * cs + c];
        if (f < 0) break;
        // the interface frame as one 80-byte record: normal and area in two sectors instead of four
        const double* gm = g.f_geom + (long)f * 10;
        const double area = gm[9];
        const double dot = vx * gm[0] + vy * gm[1] + vz * gm[2];
        const double sig_vel = fabs(dot) + a;
        sr_c += sig_vel * area;
        if (A.viscous) {
            const double Tf = A.face_T[f];
            const double mu = viscosity(Tf, re(A.mu0), re(A.T0), re(A.Ts));
            const double k = conductivity(mu, Cp, re(A.Pr));
            const double rho = A.face_rho[f];
            const double Pr = mu * Cp / k;
            const double tmp = (gamma / rho) * (mu / Pr) * area * area;
            sr_v += tmp / volume;
        }
    }
    return volume / (sr_c + A.signal_factor * sr_v);
}

// the same with `frames` = the tile's frames in GLOBAL memory (t_freca; pass F' has released its staged copy by the
// time the cells sum their faces): identical code, the loads go through L2

}  // namespace ib
