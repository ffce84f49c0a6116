// solver_bodies.h -- per-index bodies of the state-vector kernels: primitive/conserved
// conversion, Runge-Kutta updates, norms, bad-cell test, JFNK pack/unpack.
#pragma once

#include "fv_bodies.h"

namespace ib {

// conserved_to_primatives (finite_volume/primative_conserved_conversion.cpp:5-23); runs
// over ALL rows including ghosts (fs.gas.size()), where 0/0 gives NaN as in the reference.
template <int DIM, typename S>
IBIS_HD void prim_from_cons_store(const S* q /* mass mx my mz energy */, S* const* fs, const Gas<S>& gas, long i) {
    typedef mem<S> M;
    const S rho = q[0];
    const S vx = q[1] / rho;
    const S vy = q[2] / rho;
    const S vz = (DIM == 3) ? q[3] / rho : S(0.0);
    const S ke = 0.5 * (vx * vx + vy * vy + vz * vz);
    const S u = q[4] / rho - ke;
    M::store(fs[F_RHO], i, rho);
    M::store(fs[F_VX], i, vx);
    M::store(fs[F_VY], i, vy);
    M::store(fs[F_VZ], i, vz);
    M::store(fs[F_E], i, u);
    const S T = u / gas.Cv;                                   // update_thermo_from_rhou
    M::store(fs[F_T], i, T);
    M::store(fs[F_P], i, p_from_rhoT(rho, T, gas.R));
}
template <int DIM, typename S>
IBIS_HD void cons_to_prim_body(S* const* cq, S* const* fs, const Gas<S>& gas, long i) {
    typedef mem<S> M;
    S q[5];
#pragma unroll
    for (int k = 0; k < 5; k++) q[k] = (DIM == 2 && k == 3) ? S(0.0) : M::load(cq[k], i);
    prim_from_cons_store<DIM>(q, fs, gas, i);
}

// one cell of a native flow snapshot (io/native.cpp:280-336): T p vx vy vz in, thermodynamic
// closure update_thermo_from_pT (gas_model.h:91-95); row i of the arrays is reference cell i
template <typename S>
IBIS_HD void flow_read_body(const double* T_, const double* p_, const double* vx, const double* vy, const double* vz,
                            S* const* fs, const Gas<S>& gas, long i, long cell) {
    typedef mem<S> M;
    const S T = S(T_[i]), p = S(p_[i]);
    M::store(fs[F_T], cell, T);
    M::store(fs[F_P], cell, p);
    M::store(fs[F_RHO], cell, rho_from_pT(p, T, gas.R));
    M::store(fs[F_E], cell, gas.Cv * T);
    M::store(fs[F_VX], cell, S(vx[i]));
    M::store(fs[F_VY], cell, S(vy[i]));
    M::store(fs[F_VZ], cell, vz ? S(vz[i]) : S(0.0));
}
template <typename S>
IBIS_HD void flow_write_body(S* const* fs, double* T_, double* p_, double* vx, double* vy, double* vz, long i, long cell) {
    typedef mem<S> M;
    T_[i] = re(M::load(fs[F_T], cell));
    p_[i] = re(M::load(fs[F_P], cell));
    vx[i] = re(M::load(fs[F_VX], cell));
    vy[i] = re(M::load(fs[F_VY], cell));
    if (vz) vz[i] = re(M::load(fs[F_VZ], cell));
}

// primatives_to_conserved (primative_conserved_conversion.cpp:32-50)
template <int DIM, typename S>
IBIS_HD void prim_to_cons_body(S* const* cq, S* const* fs, long i) {
    typedef mem<S> M;
    const S vx = M::load(fs[F_VX], i), vy = M::load(fs[F_VY], i), vz = M::load(fs[F_VZ], i);
    const S rho = M::load(fs[F_RHO], i);
    M::store(cq[0], i, rho);
    M::store(cq[1], i, rho * vx);
    M::store(cq[2], i, rho * vy);
    if (DIM == 3) M::store(cq[3], i, rho * vz);
    const S ke = 0.5 * (vx * vx + vy * vy + vz * vz);
    M::store(cq[4], i, rho * (ke + M::load(fs[F_E], i)));
}

// U1 = U0 + dUdt*dt (conserved_quantities.cpp:80-94; in place: :26-38)
template <int DIM, typename S>
IBIS_HD void apply_dt_body(S* const* u0, S* const* u1, S* const* k, double dt, long i) {
    typedef mem<S> M;
#pragma unroll
    for (int q = 0; q < 5; q++) {
        if (DIM == 2 && q == 3) continue;
        M::store(u1[q], i, M::load(u0[q], i) + M::load(k[q], i) * dt);
    }
}

// One fused Runge-Kutta update of rk_step: the chain of apply_time_derivative calls of one stage
// (runge_kutta.cpp:155-169) or of the final update (:185-186), U <- U + k_t * dt_t term after term
// exactly as the separate kernels would round it, optionally stored (u_out), followed by the
// conserved_to_primatives of the result (:177, :193) -- one pass over the cells instead of up to
// five.  The intermediate conserved state of a stage (k_tmp_) is never needed in memory.
constexpr int RK_MAX_TERMS = 8;
template <typename S>
struct RkTerms {
    S* k[RK_MAX_TERMS][5];
    double scale[RK_MAX_TERMS];
    int n;
};
template <int DIM, typename S>
IBIS_HD void rk_update_body(S* const* u_in, const RkTerms<S>& T, double dt, S* const* u_out, S* const* fs,
                            const Gas<S>& gas, long i) {
    typedef mem<S> M;
    S q[5];
#pragma unroll
    for (int c = 0; c < 5; c++) q[c] = (DIM == 2 && c == 3) ? S(0.0) : M::load(u_in[c], i);
    for (int t = 0; t < T.n; t++) {
        const double dte = T.scale[t] * dt;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            if (DIM == 2 && c == 3) continue;
            q[c] = q[c] + M::load(T.k[t][c], i) * dte;
        }
    }
    if (u_out[0]) {
#pragma unroll
        for (int c = 0; c < 5; c++) {
            if (DIM == 2 && c == 3) continue;
            M::store(u_out[c], i, q[c]);
        }
    }
    prim_from_cons_store<DIM>(q, fs, gas, i);
}

// L2_norms contributions (conserved_quantities.cpp:43-58), real parts:
// out = {global, mass, mx, my, mz, energy}
template <int DIM, typename S>
IBIS_HD void l2_body(S* const* cq, long i, double* out) {
    typedef mem<S> M;
    const double m = re(M::load(cq[0], i)), mx = re(M::load(cq[1], i)), my = re(M::load(cq[2], i));
    const double mz = (DIM == 3) ? re(M::load(cq[3], i)) : 0.0;
    const double e = re(M::load(cq[4], i));
    out[1] = m * m;
    out[2] = mx * mx;
    out[3] = my * my;
    out[4] = mz * mz;
    out[5] = e * e;
    out[0] = m * m + mx * mx + my * my + mz * mz + e * e;
}

// count_bad_cells predicate (finite_volume.cpp:263-266; dual: isnan/isinf test both parts,
// util/dual.h:306-314)
IBIS_HD bool bad_cell(double T, double rho) {
    return T < 0.0 || rho < 0.0 || isnan(rho) || isinf(rho);
}
IBIS_HD bool bad_cell(const dual& T, const dual& rho) {
    return T.re < 0.0 || rho.re < 0.0 || isnan(rho.re) || isnan(rho.du) || isinf(rho.re) ||
           isinf(rho.du);
}

// the operations the reference's own Dual tests exercise (util/dual.cpp:4-222), evaluated with the
// operators of scalar.h: in = {x.re, x.du, y.re, y.du, r}, out = 22 (re, du) pairs
IBIS_HD void dual_kat_body(const double* in, double* out) {
    const dual x(in[0], in[1]), y(in[2], in[3]);
    const double r = in[4];
    dual t;
    int k = 0;
    auto put = [&](const dual& d) { out[2 * k] = d.re; out[2 * k + 1] = d.du; k++; };
    put(x + y);
    t = x; t += y; put(t);
    put(x + r);
    put(x - y);
    t = x; t -= y; put(t);
    put(x - r);
    put(r - x);
    put(x * y);
    t = x; t *= y; put(t);
    put(x * r);
    t = x; t *= r; put(t);
    put(r * x);
    put(x / y);
    t = x; t /= y; put(t);
    put(x / r);
    t = x; t /= r; put(t);
    put(r / x);
    put(sqrt_(x));
    put(abs_(x));
    put(abs_(x * x));
    put(x * x + 2.0 * x);
    put(x * x - sqrt_(x));
}

}  // namespace ib
