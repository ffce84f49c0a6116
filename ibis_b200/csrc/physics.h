// physics.h -- per-face / per-cell numerics of the residual, templated on the scalar
// S = double | ib::dual.  Each function cites the reference lines it evaluates; operation
// order and association follow the reference expression trees exactly (the library is
// built with --fmad=false so no contraction changes the rounding).
#pragma once

#include "scalar.h"

namespace ib {

enum { FLUX_HANEL = 0, FLUX_AUSMDV = 1, FLUX_LDFSS2 = 2, FLUX_RUSANOV = 3 };
enum { THERMO_RHO_P = 0, THERMO_RHO_T = 1, THERMO_RHO_U = 2, THERMO_P_T = 3 };

// IdealGas<T> (gas/gas_model.h:47-189); in the dual instantiation the constants are
// duals with zero dual part and enter the arithmetic as such.
template <typename S>
struct Gas {
    S R, Cv, Cp, gamma;
};

// FlowState of one side of a face, velocity already in the face frame
template <typename S>
struct Side {
    S rho, p, T, e, u, v, w;
};

template <typename S>
struct Flux5 {
    S mass, mx, my, mz, energy;
};

// gas_model.h:12-40
template <typename S> IBIS_HD S rho_from_pT(S p, S T, S R) { return p / (R * T); }
template <typename S> IBIS_HD S T_from_rhop(S rho, S p, S R) { return p / (rho * R); }
template <typename S> IBIS_HD S p_from_rhoT(S rho, S T, S R) { return rho * R * T; }
template <typename S> IBIS_HD S speed_of_sound(S T, S R, S gamma) { return sqrt_(gamma * R * T); }

// update_thermo_from_* on a single state (gas_model.h:91-117)
template <typename S>
IBIS_HD void thermo_close(int thermo, const Gas<S>& g, S& rho, S& p, S& T, S& e) {
    switch (thermo) {
        case THERMO_RHO_P:                       // update_thermo_from_rhop
            T = T_from_rhop(rho, p, g.R);
            e = g.Cv * T;
            break;
        case THERMO_RHO_T:                       // update_thermo_from_rhoT
            p = p_from_rhoT(rho, T, g.R);
            e = g.Cv * T;
            break;
        case THERMO_RHO_U:                       // update_thermo_from_rhou
            T = e / g.Cv;
            p = p_from_rhoT(rho, T, g.R);
            break;
        default:                                 // update_thermo_from_pT
            rho = rho_from_pT(p, T, g.R);
            e = g.Cv * T;
            break;
    }
}

// ---------------------------------------------------------------------------------------
// flux calculators: L/R in the face frame -> flux in the face frame
// ---------------------------------------------------------------------------------------

// finite_volume/flux_calculators/hanel.cpp:10-64
template <typename S>
IBIS_HD Flux5<S> flux_hanel(const Side<S>& L, const Side<S>& R_, const Gas<S>& g) {
    S rL = L.rho, pL = L.p;
    S eL = g.Cv * L.T;
    S aL = speed_of_sound(L.T, g.R, g.gamma);
    S uL = L.u, vL = L.v, wL = L.w;
    S keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    S pLrL = pL / rL;
    S HL = eL + pLrL + keL;

    S rR = R_.rho, pR = R_.p;
    S eR = g.Cv * R_.T;
    S aR = speed_of_sound(R_.T, g.R, g.gamma);
    S uR = R_.u, vR = R_.v, wR = R_.w;
    S keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    S pRrR = pR / rR;
    S HR = eR + pRrR + keR;

    S pLplus, uLplus;
    if (re(abs_(uL)) <= re(aL)) {
        uLplus = 1.0 / (4.0 * aL) * (uL + aL) * (uL + aL);
        pLplus = pL * uLplus * (1.0 / aL * (2.0 - uL / aL));
    } else {
        uLplus = 0.5 * (uL + abs_(uL));
        pLplus = pL * uLplus * (1.0 / uL);
    }
    S pRminus, uRminus;
    if (re(abs_(uR)) <= re(aR)) {
        uRminus = -1.0 / (4.0 * aR) * (uR - aR) * (uR - aR);
        pRminus = pR * uRminus * (1.0 / aR * (-2.0 - uR / aR));
    } else {
        uRminus = 0.5 * (uR - abs_(uR));
        pRminus = pR * uRminus * (1.0 / uR);
    }
    S p_half = pLplus + pRminus;
    Flux5<S> f;
    f.mass = uLplus * rL + uRminus * rR;
    f.mx = uLplus * rL * uL + uRminus * rR * uR + p_half;
    f.my = uLplus * rL * vL + uRminus * rR * vR;
    f.mz = uLplus * rL * wL + uRminus * rR * wR;
    f.energy = uLplus * rL * HL + uRminus * rR * HR;
    return f;
}

// finite_volume/flux_calculators/ausmdv.cpp:12-131
template <typename S>
IBIS_HD Flux5<S> flux_ausmdv(const Side<S>& L, const Side<S>& R_, const Gas<S>& g) {
    // the quotients of the reference's expressions, two at a time where they are independent or share their
    // denominator (scalar.h::div_pair / div_shared: the same quotients, bit for bit)
    S rL = L.rho, pL = L.p;
    S rR = R_.rho, pR = R_.p;
    S pLrL, pRrR;
    div_pair(pL, rL, pR, rR, pLrL, pRrR);
    S uL = L.u, vL = L.v, wL = L.w;
    S eL = g.Cv * L.T;
    S aL = speed_of_sound(L.T, g.R, g.gamma);
    S keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    S HL = eL + pLrL + keL;

    S uR = R_.u, vR = R_.v, wR = R_.w;
    S eR = g.Cv * R_.T;
    S aR = speed_of_sound(R_.T, g.R, g.gamma);
    S keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    S HR = eR + pRrR + keR;

    S am = max_(aL, aR);
    S ML, MR;
    div_shared(uL, uR, am, ML, MR);
    // alphaL, alphaR (ausmdv.cpp:38-39) and (uL + am)^2 / (4 am), -(uR - am)^2 / (4 am) are read by the subsonic
    // branches only: a supersonic interface (both sides) skips their three reciprocals
    const bool subL = re(abs_(ML)) <= 1.0, subR = re(abs_(MR)) <= 1.0;
    S alphaL = S(0.0), alphaR = S(0.0), qLs = S(0.0), qRs = S(0.0);
    if (subL || subR) {
        div_shared(2.0 * pLrL, 2.0 * pRrR, pLrL + pRrR, alphaL, alphaR);
        div_shared((uL + am) * (uL + am), -(uR - am) * (uR - am), 4.0 * am, qLs, qRs);
    }

    S pLplus, uLplus;
    S duL = 0.5 * (uL + abs_(uL));
    if (subL) {
        pLplus = pL * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) * 0.25;
        uLplus = alphaL * (qLs - duL) + duL;
    } else {
        pLplus = div_z(pL * duL, uL);
        uLplus = duL;
    }
    S pRminus, uRminus;
    S duR = 0.5 * (uR - abs_(uR));
    if (subR) {
        pRminus = pR * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) * 0.25;
        uRminus = alphaR * (qRs - duR) + duR;
    } else {
        pRminus = div_z(pR * duR, uR);
        uRminus = duR;
    }
    S ru_half = uLplus * rL + uRminus * rR;
    S p_half = pLplus + pRminus;

    S dp = pL - pR;
    const S K_SWITCH = 10.0;
    dp = div_z(K_SWITCH * abs_(dp), min_(pL, pR));
    S s = 0.5 * min_(1.0, dp);
    S ru2_AUSMV = uLplus * rL * uL + uRminus * rR * uR;
    S ru2_AUSMD = 0.5 * (ru_half * (uL + uR) - abs_(ru_half) * (uR - uL));
    S ru2_half = (0.5 + s) * ru2_AUSMV + (0.5 - s) * ru2_AUSMD;

    Flux5<S> f;
    f.mass = ru_half;
    if (re(ru_half) >= 0.0) {
        f.mx = (ru2_half + p_half);
        f.my = (ru_half * vL);
        f.mz = (ru_half * wL);
        f.energy = ru_half * HL;
    } else {
        f.mx = (ru2_half + p_half);
        f.my = (ru_half * vR);
        f.mz = (ru_half * wR);
        f.energy = ru_half * HR;
    }

    const S C_EFIX = 0.125;
    bool caseA = (re(uL - aL) < 0.0) && (re(uR - aR) > 0.0);
    bool caseB = (re(uL + aL) < 0.0) && (re(uR + aR) > 0.0);
    S d_ua = 0.0;
    if (caseA && !caseB) d_ua = C_EFIX * ((uR - aR) - (uL - aL));
    if (caseB && !caseA) d_ua = C_EFIX * ((uR + aR) - (uL + aL));
    if (re(d_ua) != 0.0) {
        f.mass -= d_ua * (rR - rL);
        f.mx -= d_ua * (rR * uR - rL * uL);
        f.my -= d_ua * (rR * vR - rL * vL);
        f.mz -= d_ua * (rR * wR - rL * wL);
        f.energy -= d_ua * (rR * HR - rL * HL);
    }
    return f;
}

// finite_volume/flux_calculators/ldfss2.cpp:14-89
template <typename S>
IBIS_HD Flux5<S> flux_ldfss2(const Side<S>& L, const Side<S>& R_, const Gas<S>& g, double delta) {
    S rL = L.rho, pL = L.p;
    S pLrL = pL / rL;
    S uL = L.u, vL = L.v, wL = L.w;
    S eL = g.Cv * L.T;
    S aL = speed_of_sound(L.T, g.R, g.gamma);
    S keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    S HL = eL + pLrL + keL;

    S rR = R_.rho, pR = R_.p;
    S pRrR = pR / rR;
    S uR = R_.u, vR = R_.v, wR = R_.w;
    S eR = g.Cv * R_.T;
    S aR = speed_of_sound(R_.T, g.R, g.gamma);
    S keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    S HR = eR + pRrR + keR;

    S am = 0.5 * (aL + aR);
    S ML = uL / am;
    S MR = uR / am;

    S MpL = 0.25 * (ML + 1.0) * (ML + 1.0);
    S MmR = -0.25 * (MR - 1.0) * (MR - 1.0);

    S alphaL = 0.5 * (1.0 + copysign_(S(1.0), ML));
    S alphaR = 0.5 * (1.0 - copysign_(S(1.0), MR));

    S betaL = -max_(0.0, 1.0 - floor_(abs_(ML)));
    S betaR = -max_(0.0, 1.0 - floor_(abs_(MR)));

    S PL = 0.25 * (ML + 1.0) * (ML + 1.0) * (2.0 - ML);
    S PR = 0.25 * (MR - 1.0) * (MR - 1.0) * (2.0 + MR);

    S DL = alphaL * (1.0 + betaL) - betaL * PL;
    S DR = alphaR * (1.0 + betaR) - betaR * PR;

    S Mhalf = 0.25 * betaL * betaR * pow2_((sqrt_(0.5 * (ML * ML + MR * MR)) - 1.0));

    S MhalfL = Mhalf * (1.0 - (div_z(pL - pR, pL + pR) + delta * div_z(abs_(pL - pR), pL)));
    S MhalfR = Mhalf * (1.0 + (div_z(pL - pR, pL + pR) - delta * div_z(abs_(pL - pR), pR)));

    S CL = alphaL * (1.0 + betaL) * ML - betaL * MpL - MhalfL;
    S CR = alphaR * (1.0 + betaR) * MR - betaR * MmR + MhalfR;

    S ru_half = am * rL * CL + am * rR * CR;
    S ru2_half = am * rL * CL * uL + am * rR * CR * uR;
    S p_half = DL * pL + DR * pR;
    Flux5<S> f;
    f.mass = ru_half;
    f.mx = ru2_half + p_half;
    f.my = am * rL * CL * vL + am * rR * CR * vR;
    f.mz = am * rL * CL * wL + am * rR * CR * wR;
    f.energy = am * rL * CL * HL + am * rR * CR * HR;
    return f;
}

// finite_volume/flux_calculators/rusanov.cpp:13-52
template <typename S>
IBIS_HD Flux5<S> flux_rusanov(const Side<S>& L, const Side<S>& R_, const Gas<S>& g) {
    S rL = L.rho;
    S eL = g.Cv * L.T;
    S pL = L.p;
    S uL = L.u, vL = L.v, wL = L.w;
    S ruL = rL * uL;
    S keL = 0.5 * (uL * uL + vL * vL + wL * wL);
    S HL = eL + pL / rL + keL;
    S aL = speed_of_sound(L.T, g.R, g.gamma);

    S rR = R_.rho;
    S eR = g.Cv * R_.T;
    S pR = R_.p;
    S uR = R_.u, vR = R_.v, wR = R_.w;
    S ruR = rR * uR;
    S keR = 0.5 * (uR * uR + vR * vR + wR * wR);
    S HR = eR + pR / rR + keR;
    S aR = speed_of_sound(R_.T, g.R, g.gamma);

    S S_plus = max_(abs_(uL - aL), abs_(uR - aR));
    S_plus = max_(S_plus, abs_(uL + aL));
    S_plus = max_(S_plus, abs_(uR + aR));

    Flux5<S> f;
    f.mass = 0.5 * (ruL + ruR) - 0.5 * S_plus * (rR - rL);
    f.mx = 0.5 * (ruL * uL + pL + ruR * uR + pR) - 0.5 * S_plus * (rR * uR - rL * uL);
    f.my = 0.5 * (ruL * vL + ruR * vR) - 0.5 * S_plus * (rR * vR - rL * vL);
    f.mz = 0.5 * (ruL * wL + ruR * wR) - 0.5 * S_plus * (rR * wR - rL * wL);
    f.energy = 0.5 * (ruL * HL + ruR * HR) - 0.5 * S_plus * (rR * HR - rL * HL);
    return f;
}

template <int FLUX, typename S>
IBIS_HD Flux5<S> flux_calc(const Side<S>& L, const Side<S>& R_, const Gas<S>& g, double delta) {
    if (FLUX == FLUX_HANEL) return flux_hanel(L, R_, g);
    if (FLUX == FLUX_AUSMDV) return flux_ausmdv(L, R_, g);
    if (FLUX == FLUX_LDFSS2) return flux_ldfss2(L, R_, g, delta);
    return flux_rusanov(L, R_, g);
}

// Sutherland viscosity (gas/transport_properties.h:29-35) and constant-Prandtl
// conductivity (transport_properties.h:62-67)
template <typename S>
IBIS_HD S viscosity(S T, S mu0, S T0, S Ts) {
    return mu0 * pow15_(T / T0) * (T0 + Ts) / (T + Ts);
}
template <typename S>
IBIS_HD S conductivity(S mu, S Cp, S Pr) {
    return Cp * mu / Pr;
}

}  // namespace ib
