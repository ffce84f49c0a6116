"""oracle/restatement.py -- TEST INFRASTRUCTURE, not product code.

A plain-numpy restatement of the reference's residual, FiniteVolume<real>::compute_dudt
(/root/reference/src/finite_volume/finite_volume/finite_volume.cpp:57-96), written from the reference's
sources function by function (each block cites the file:line it follows) and independent of the product's
kernel bodies.  It exists next to the compiled reference (oracle/_ref, the primary oracle) for two reasons:
it travels as source, and it is a second, independently written statement of the same arithmetic.

Pinned: tests/test_oracle_restatement.py holds it BIT FOR BIT to the golden vectors the compiled reference
produced (tests/golden/*.npz: residual, ghost fill, every gradient and limiter, dt, conversions; and
moving_*.npz) on all 26 static and 8 moving-grid cases of tests/cases.py -- inviscid and viscous; hexes, mixed
cells, tets, quads, triangles; ausmdv / hanel / rusanov / ldfss2; limited and unlimited; rho_u / rho_p / rho_T;
supersonic, subsonic, wall, profile and constant-flux boundaries; first and second order.  Every operation
below is an individually rounded IEEE operation in the reference's association order (numpy does not contract
to FMA; pow is glibc's on both sides), which is why equality is exact.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Covered: all eight ghost-cell actions (boundary_layer_profile with CubicSpline) and the constant_flux flux
action; WLS r-factors and gradients; Barth-Jespersen / unlimited; MUSCL reconstruction with the three working
thermo interpolators; the four flux calculators; the viscous flux (pre_viscous_grad actions, Hasselbacher face
gradients, Sutherland / constant-Prandtl transport); the surface integral; the moving-grid terms (interface
velocities, relative fluxes, geometric conservation law); estimate_dt with its viscous term, count_bad_cells,
L2 norms, primitive <-> conserved -- all of it on dual numbers too (oracle/dualnum.py: one exact Jacobian-vector
product direction of the steady solver); and the solvers: Solver::solve + RungeKutta on a static grid (bit for bit) and SteadyState / Jfnk /
Gmres / FGmres with dual-number matrix-vector products (1e-9), held to the final snapshots of the reference's own runs
(tests/golden/solve_*.npz), moving grids with the geometry recomputed every stage included.
"""
import numpy as np

from dualnum import Dual, as_T, dmax, dmin

R_AIR = 287.0


def _min(a, b):
    """Ibis::min -> Kokkos::min(a, b) = (b < a) ? b : a for doubles (util/real.h:15-65); dual.h:284-298 for duals"""
    if isinstance(a, Dual) or isinstance(b, Dual):
        return dmin(a, b)
    return np.where(b < a, b, a)


def _max(a, b):
    """Kokkos::max(a, b) = (a < b) ? b : a for doubles; dual.h:269-282 for duals"""
    if isinstance(a, Dual) or isinstance(b, Dual):
        return dmax(a, b)
    return np.where(a < b, b, a)


class _Ragged:
    """cell -> (faces, neighbours, outsigns) in cell-face order, as padded (n_cells, max_faces) tables + mask"""

    def __init__(self, g):
        off = np.asarray(g.cell_face_offsets, np.int64)
        n = len(off) - 1
        cnt = np.diff(off)
        self.maxf = int(cnt.max()) if n else 0
        slot = np.arange(self.maxf)[None, :]
        self.mask = slot < cnt[:, None]
        idx = np.where(self.mask, off[:-1, None] + slot, 0)
        self.face = np.where(self.mask, np.asarray(g.cell_face_ids, np.int64)[idx], 0)
        self.nbr = np.where(self.mask, np.asarray(g.cell_neighbours, np.int64)[idx], 0)
        self.out = np.where(self.mask, np.asarray(g.cell_outsigns, np.int64)[idx], 0)


class CubicSpline:
    """util/cubic_spline.cpp:9-89 (constructor: assembly + Thomas sweep for the interior second derivatives) and
    util/cubic_spline.h:14-45 (eval).  The reference stores y'' in an array of n - 2 entries but indexes it by
    knot, i.e. it reads past the end in the last two intervals; here y'' holds the n values those formulas
    address (natural spline) -- identical wherever the reference is defined."""

    def __init__(self, x, y):
        x, y = [float(v) for v in x], [float(v) for v in y]
        n = len(x)
        if n < 4:
            raise ValueError("Insufficient points for cubic spline")
        if len(y) != n:
            raise ValueError("Cubic spline has different number of x and y points")
        ne = n - 2
        diag, upper, rhs = [0.0] * ne, [0.0] * ne, [0.0] * ne
        for i in range(ne + 1):
            dxi, dyi = x[i + 1] - x[i], y[i + 1] - y[i]
            if i != ne:
                diag[i] += dxi
                rhs[i] -= dyi / dxi
            if i == 0:
                continue
            rhs[i - 1] += dyi / dxi
            diag[i - 1] += dxi
            upper[i - 1] += dxi
        rhs = [v * 6 for v in rhs]
        diag = [v * 2 for v in diag]
        for i in range(1, ne):
            w = upper[i - 1] / diag[i - 1]
            diag[i] = diag[i] - w * upper[i - 1]
            rhs[i] = rhs[i] - w * rhs[i - 1]
        ydd = [0.0] * n
        ydd[n - 2] = rhs[ne - 1] / diag[ne - 1]
        for i in range(ne - 2, -1, -1):
            ydd[i + 1] = (rhs[i] - upper[i] * ydd[i + 2]) / diag[i]
        self.x, self.y, self.ydd = x, y, ydd

    def eval(self, q):
        x, y, ydd = self.x, self.y, self.ydd
        if q <= x[0]:
            return y[0]
        if q >= x[-1]:
            return y[-1]
        idx = next(i for i in range(len(x)) if x[i] >= q) - 1
        dx_plus, dx_minus, delta = x[idx + 1] - q, q - x[idx], x[idx + 1] - x[idx]
        a = (ydd[idx] * dx_plus * dx_plus * dx_plus + ydd[idx + 1] * dx_minus * dx_minus * dx_minus) / (6 * delta)
        b = (y[idx] / delta - ydd[idx] * delta / 6) * dx_plus
        c = (y[idx + 1] / delta - ydd[idx + 1] * delta / 6) * dx_minus
        return a + b + c


def apply_pre_reconstruction(g, config, prim, phase="pre_reconstruction"):
    """finite_volume.cpp:187-218: tags in alphabetical order, actions in list order (boundary.cpp:517-543);
    phase = "pre_reconstruction" or "pre_viscous_grad"""
    fs = prim
    n = g.n_cells
    left, right = np.asarray(g.face_left), np.asarray(g.face_right)
    for tag in sorted(g.tags):
        faces = np.asarray(g.tags[tag]["faces"], np.int64)
        l, r = left[faces], right[faces]
        ghost = np.where(l < n, r, l)
        valid = np.where(l < n, l, r)
        for act in config["grid"]["boundaries"][tag][phase]:
            kind = act["type"]
            if kind == "flow_state_copy":                        # boundary.cpp:15-45
                s = act["flow_state"]
                fs[ghost] = [s["rho"], s["p"], s["T"], s["energy"], s["vx"], s["vy"], s["vz"]]
            elif kind == "internal_copy":                        # boundary.cpp:100-136
                fs[ghost] = fs[valid]
            elif kind == "internal_copy_reflect_normal":         # boundary.cpp:139-206
                fs[ghost, :4] = fs[valid, :4]
                x, y, z = fs[valid, 4], fs[valid, 5], fs[valid, 6]
                nrm, t1, t2 = g.face_norm[faces], g.face_tan1[faces], g.face_tan2[faces]
                xs = -(x * nrm[:, 0] + y * nrm[:, 1] + z * nrm[:, 2])
                ys = x * t1[:, 0] + y * t1[:, 1] + z * t1[:, 2]
                zs = x * t2[:, 0] + y * t2[:, 1] + z * t2[:, 2]
                fs[ghost, 4] = xs * nrm[:, 0] + ys * t1[:, 0] + zs * t2[:, 0]
                fs[ghost, 5] = xs * nrm[:, 1] + ys * t1[:, 1] + zs * t2[:, 1]
                fs[ghost, 6] = xs * nrm[:, 2] + ys * t1[:, 2] + zs * t2[:, 2]
            elif kind == "internal_vel_copy_reflect":            # boundary.cpp:209-240
                fs[ghost, 4], fs[ghost, 5], fs[ghost, 6] = -fs[valid, 4], -fs[valid, 5], -fs[valid, 6]
            elif kind == "fix_temperature":                      # boundary.cpp:243-276
                fs[ghost, 2] = 2.0 * float(act["temperature"]) - fs[valid, 2]
            elif kind == "boundary_layer_profile":               # boundary.cpp:51-95
                prof = act["profile"]
                gm = config["gas_model"]
                vs, Ts = CubicSpline(prof["height"], prof["v"]), CubicSpline(prof["height"], prof["T"])
                pos = np.asarray(g.cell_centroid)[ghost, 1]
                T = np.array([Ts.eval(float(q)) for q in pos])
                p = float(prof["p"])
                fs[ghost, 1], fs[ghost, 2] = p, T
                fs[ghost, 0] = p / (float(gm["R"]) * T)          # update_thermo_from_pT (gas_model.h:91-96)
                fs[ghost, 3] = float(gm["Cv"]) * T
                fs[ghost, 4] = [vs.eval(float(q)) for q in pos]
                fs[ghost, 5], fs[ghost, 6] = 0.0, 0.0
            elif kind in ("subsonic_inflow", "subsonic_outflow"):  # boundary.cpp:279-342, 345-400
                gm = config["gas_model"]
                nrm = g.face_norm[faces]
                nx, ny, nz = nrm[:, 0], nrm[:, 1], nrm[:, 2]
                a = np.sqrt(float(gm["gamma"]) * float(gm["R"]) * fs[valid, 2])
                rho_ref, p_v = fs[valid, 0], fs[valid, 1]
                vx, vy, vz = fs[valid, 4], fs[valid, 5], fs[valid, 6]
                if kind == "subsonic_inflow":
                    s = act["flow_state"]
                    vel_correct = nx * (s["vx"] - vx) + ny * (s["vy"] - vy) + nz * (s["vz"] - vz)
                    p_in = s["p"]
                    p_face = 0.5 * (p_in + p_v - rho_ref * a * vel_correct)
                    rho_face = s["rho"] + (p_face - p_in) / a / a
                    vx_face = s["vx"] - nx * (p_in - p_face) / (rho_ref * a)
                    vy_face = s["vy"] - ny * (p_in - p_face) / (rho_ref * a)
                    vz_face = s["vz"] - nz * (p_in - p_face) / (rho_ref * a)
                else:
                    p_face = float(act["pressure"])
                    rho_face = rho_ref + (p_face - p_v) / a / a
                    vx_face = vx + nx * (p_v - p_face) / (rho_ref * a)
                    vy_face = vy + ny * (p_v - p_face) / (rho_ref * a)
                    vz_face = vz + nz * (p_v - p_face) / (rho_ref * a)
                pg = 2.0 * p_face - p_v
                rg = 2.0 * rho_face - rho_ref
                Tg = pg / (rg * float(gm["R"]))                  # update_thermo_from_rhop (gas_model.h:105-110)
                fs[ghost, 1], fs[ghost, 0], fs[ghost, 2], fs[ghost, 3] = pg, rg, Tg, float(gm["Cv"]) * Tg
                fs[ghost, 4], fs[ghost, 5], fs[ghost, 6] = 2.0 * vx_face - vx, 2.0 * vy_face - vy, 2.0 * vz_face - vz
            else:
                raise NotImplementedError(f"boundary action {kind}: use the compiled reference (oracle/_ref)")
    return fs


def wls_r(g, rag):
    """WLSGradient::compute_weights (grid/gradient.h:119-162) -> r11 r12 r22 r13 r23 r33"""
    n = g.n_cells
    c = np.asarray(g.cell_centroid)
    s = [np.zeros(n) for _ in range(6)]
    for j in range(rag.maxf):
        m, nb = rag.mask[:, j], rag.nbr[:, j]
        dx, dy, dz = c[nb, 0] - c[:n, 0], c[nb, 1] - c[:n, 1], c[nb, 2] - c[:n, 2]
        for k, t in enumerate((dx * dx, dx * dy, dx * dz, dy * dy, dy * dz, dz * dz)):
            s[k] = np.where(m, s[k] + t, s[k])
    sxx, sxy, sxz, syy, syz, szz = s
    with np.errstate(all="ignore"):
        r11 = np.sqrt(sxx)
        r12 = 1.0 / r11 * sxy
        r22 = np.sqrt(syy - r12 * r12)
        r13 = 1.0 / r11 * sxz
        r23 = 1.0 / r22 * (syz - r12 / r11 * sxz)
        r33 = np.sqrt(szz - (r13 * r13 + r23 * r23))
    if g.dim == 2:                                               # gradient.h:156-160: only stored in 3-D
        r13, r23, r33 = np.zeros(n), np.zeros(n), np.zeros(n)
    return r11, r12, r22, r13, r23, r33


def wls_gradient(g, rag, r, values):
    """WLSGradient::compute_gradients (grid/gradient.h:66-116)"""
    n = g.n_cells
    c = np.asarray(g.cell_centroid)
    r11, r12, r22, r13, r23, r33 = r
    gx, gy, gz = np.zeros(n), np.zeros(n), np.zeros(n)
    with np.errstate(all="ignore"):
        beta = (r12 * r23 - r13 * r23) / (r11 * r22)
        for j in range(rag.maxf):
            m, nb = rag.mask[:, j], rag.nbr[:, j]
            diff_u = values[nb] - values[:n]
            dx, dy, dz = c[nb, 0] - c[:n, 0], c[nb, 1] - c[:n, 1], c[nb, 2] - c[:n, 2]
            alpha_1 = dx / (r11 * r11)
            alpha_2 = 1.0 / (r22 * r22) * (dy - r12 / r11 * dx)
            alpha_3 = np.zeros(n)
            if g.dim == 3:
                alpha_3 = 1.0 / (r33 * r33) * (dz - r23 / r22 * dy + beta * dx)
            w_1 = alpha_1 - r12 / r11 * alpha_2 + beta * alpha_3
            w_2 = alpha_2 - r23 / r22 * alpha_3
            w_3 = alpha_3
            gx = np.where(m, gx + w_1 * diff_u, gx)
            gy = np.where(m, gy + w_2 * diff_u, gy)
            gz = np.where(m, gz + w_3 * diff_u, gz)
    return np.stack([gx, gy, gz], axis=1)


def barth_jespersen(g, rag, values, grad, epsilon):
    """BarthJespersen::calculate_limiters (finite_volume/limiter.cpp:24-63)"""
    n = g.n_cells
    ui = values[:n]
    umin, umax = ui.copy(), ui.copy()
    for j in range(rag.maxf):
        m, un = rag.mask[:, j], values[rag.nbr[:, j]]
        umin = np.where(m, _min(umin, un), umin)
        umax = np.where(m, _max(umax, un), umax)
    phi = np.ones(n)
    c, fc = np.asarray(g.cell_centroid), np.asarray(g.face_centre)
    with np.errstate(all="ignore"):
        for j in range(rag.maxf):
            m, f = rag.mask[:, j], rag.face[:, j]
            dx, dy, dz = fc[f, 0] - c[:n, 0], fc[f, 1] - c[:n, 1], fc[f, 2] - c[:n, 2]
            d2 = grad[:, 0] * dx + grad[:, 1] * dy + grad[:, 2] * dz
            sign = (d2 > 0).astype(np.int64) - (d2 < 0).astype(np.int64)
            d2 = sign * (np.abs(d2) + epsilon)
            pos = _min(phi, (umax - ui) / d2)
            neg = _min(phi, (umin - ui) / d2)
            phi = np.where(m & (sign > 0), pos, np.where(m & (sign < 0), neg, phi))
    return phi


def _thermo_close(kind, rho, th2, R, Cv):
    """update_thermo_from_rhop / rhou / rhoT on the interface states (gas/gas_model.h:91-117) -> p, T, e"""
    if kind == "rho_p":
        p = th2
        T = p / (rho * R)
        return p, T, Cv * T
    if kind == "rho_u":
        e = th2
        T = e / as_T(Cv, e)                                      # temp_from_energy(u, Cv_): Cv_ is a T
        return rho * R * T, T, e
    if kind == "rho_T":
        T = th2
        return rho * R * T, T, Cv * T
    raise NotImplementedError("thermo_interpolator p_T is broken upstream (convective_flux.cpp:53-55)")


def reconstruct(g, config, fs, rag, want):
    """copy_reconstruct (convective_flux.cpp:212-240) or linear_reconstruct (:258-420) -> left, right states
    (rho p T e vx vy vz per interface) and the gradients / limiters by name"""
    n = g.n_cells
    cf = config["convective_flux"]
    gm = config["gas_model"]
    left, right = np.asarray(g.face_left), np.asarray(g.face_right)
    if not want:
        return fs[left].copy(), fs[right].copy(), {}
    kind = cf["thermo_interpolator"]
    col2, name2 = {"rho_p": (1, "p"), "rho_u": (3, "u"), "rho_T": (2, "T")}[kind]
    r = wls_r(g, rag)
    # compute_convective_gradient (:181-209): thermo pair, then vx, vy and (3-D only) vz
    var_cols = [("rho", 0), (name2, col2), ("vx", 4), ("vy", 5), ("vz", 6)]
    inter = {}
    grads, lims = {}, {}
    limited = cf["limiter"]["type"] == "barth_jespersen"
    for name, col in var_cols:
        if name == "vz" and g.dim == 2:
            grads[name] = np.zeros((n, 3))                       # never computed in 2-D: the Views stay zero
        else:
            grads[name] = wls_gradient(g, rag, r, fs[:, col])
        inter["grad_" + name] = grads[name]
        if limited:                                              # compute_limiters (:423-462)
            lims[name] = barth_jespersen(g, rag, fs[:, col], grads[name], float(cf["limiter"].get("epsilon", 1e-25)))
            inter["lim_" + name] = lims[name]
    c, fc = np.asarray(g.cell_centroid), np.asarray(g.face_centre)
    sides = []
    for cell in (left, right):
        valid = cell < n
        safe = np.where(valid, cell, 0)
        dx, dy, dz = fc[:, 0] - c[cell, 0], fc[:, 1] - c[cell, 1], fc[:, 2] - c[cell, 2]

        def interp(name, col):                                   # linear_interpolate (:243-255)
            gr = np.where(valid[:, None], grads[name][safe], 0.0)
            lim = np.where(valid, lims[name][safe], 1.0) if limited else np.ones(len(cell))
            return fs[cell, col] + lim * (gr[:, 0] * dx + gr[:, 1] * dy + gr[:, 2] * dz)

        rho = interp("rho", 0)
        th2 = interp(name2, col2)
        p, T, e = _thermo_close(kind, rho, th2, float(gm["R"]), float(gm["Cv"]))
        sides.append(np.stack([rho, p, T, e, interp("vx", 4), interp("vy", 5), interp("vz", 6)], axis=1))
    return sides[0], sides[1], inter


def _to_local(v, nrm, t1, t2):
    """transform_to_local_frame (util/vector3.cpp:142-153)"""
    x = v[:, 0] * nrm[:, 0] + v[:, 1] * nrm[:, 1] + v[:, 2] * nrm[:, 2]
    y = v[:, 0] * t1[:, 0] + v[:, 1] * t1[:, 1] + v[:, 2] * t1[:, 2]
    z = v[:, 0] * t2[:, 0] + v[:, 1] * t2[:, 1] + v[:, 2] * t2[:, 2]
    return x, y, z


def _unpack(s, vel, gm):
    rho, p, T = s[:, 0], s[:, 1], s[:, 2]
    u, v, w = vel
    e = float(gm["Cv"]) * T                                      # gm.internal_energy (gas_model.h:164-170)
    a = np.sqrt(float(gm["gamma"]) * float(gm["R"]) * T)         # speed_of_sound_ (gas_model.h:27-31)
    ke = 0.5 * (u * u + v * v + w * w)
    return rho, p, u, v, w, e, a, ke


def flux_ausmdv(L, R_, vl, vr, gm, three_d):
    """Ausmdv::compute_flux (flux_calculators/ausmdv.cpp:8-132) -> mass, mx, my, mz, energy"""
    rL, pL, uL, vL, wL, eL, aL, keL = _unpack(L, vl, gm)
    rR, pR, uR, vR, wR, eR, aR, keR = _unpack(R_, vr, gm)
    pLrL, pRrR = pL / rL, pR / rR
    HL, HR = eL + pLrL + keL, eR + pRrR + keR
    alphaL = 2.0 * pLrL / (pLrL + pRrR)
    alphaR = 2.0 * pRrR / (pLrL + pRrR)
    am = _max(aL, aR)
    ML, MR = uL / am, uR / am
    duL = 0.5 * (uL + np.abs(uL))
    subL = np.abs(ML) <= 1.0
    pLplus = np.where(subL, pL * (ML + 1.0) * (ML + 1.0) * (2.0 - ML) * 0.25, pL * duL / uL)
    uLplus = np.where(subL, alphaL * ((uL + am) * (uL + am) / (4.0 * am) - duL) + duL, duL)
    duR = 0.5 * (uR - np.abs(uR))
    subR = np.abs(MR) <= 1.0
    pRminus = np.where(subR, pR * (MR - 1.0) * (MR - 1.0) * (2.0 + MR) * 0.25, pR * duR / uR)
    uRminus = np.where(subR, alphaR * (-(uR - am) * (uR - am) / (4.0 * am) - duR) + duR, duR)
    ru_half = uLplus * rL + uRminus * rR
    p_half = pLplus + pRminus
    dp = pL - pR
    dp = 10.0 * np.abs(dp) / _min(pL, pR)
    s = 0.5 * _min(np.ones_like(dp), dp)
    ru2_AUSMV = uLplus * rL * uL + uRminus * rR * uR
    ru2_AUSMD = 0.5 * (ru_half * (uL + uR) - np.abs(ru_half) * (uR - uL))
    ru2_half = (0.5 + s) * ru2_AUSMV + (0.5 - s) * ru2_AUSMD
    fromL = ru_half >= 0.0
    mass = ru_half.copy()
    mx = ru2_half + p_half
    my = np.where(fromL, ru_half * vL, ru_half * vR)
    mz = np.where(fromL, ru_half * wL, ru_half * wR)
    en = np.where(fromL, ru_half * HL, ru_half * HR)
    caseA = ((uL - aL) < 0.0) & ((uR - aR) > 0.0)
    caseB = ((uL + aL) < 0.0) & ((uR + aR) > 0.0)
    d_ua = np.zeros_like(mass)
    d_ua = np.where(caseA & ~caseB, 0.125 * ((uR - aR) - (uL - aL)), d_ua)
    d_ua = np.where(caseB & ~caseA, 0.125 * ((uR + aR) - (uL + aL)), d_ua)
    fix = d_ua != 0.0
    mass = np.where(fix, mass - d_ua * (rR - rL), mass)
    mx = np.where(fix, mx - d_ua * (rR * uR - rL * uL), mx)
    my = np.where(fix, my - d_ua * (rR * vR - rL * vL), my)
    mz = np.where(fix, mz - d_ua * (rR * wR - rL * wL), mz)
    en = np.where(fix, en - d_ua * (rR * HR - rL * HL), en)
    return mass, mx, my, (mz if three_d else np.zeros_like(mz)), en


def flux_hanel(L, R_, vl, vr, gm, three_d):
    """Hanel::compute_flux (flux_calculators/hanel.cpp:7-65)"""
    rL, pL, uL, vL, wL, eL, aL, keL = _unpack(L, vl, gm)
    rR, pR, uR, vR, wR, eR, aR, keR = _unpack(R_, vr, gm)
    HL, HR = eL + pL / rL + keL, eR + pR / rR + keR
    subL = np.abs(uL) <= aL
    uLplus = np.where(subL, 1.0 / (4.0 * aL) * (uL + aL) * (uL + aL), 0.5 * (uL + np.abs(uL)))
    pLplus = np.where(subL, pL * uLplus * (1.0 / aL * (2.0 - uL / aL)), pL * uLplus * (1.0 / uL))
    subR = np.abs(uR) <= aR
    uRminus = np.where(subR, -1.0 / (4.0 * aR) * (uR - aR) * (uR - aR), 0.5 * (uR - np.abs(uR)))
    pRminus = np.where(subR, pR * uRminus * (1.0 / aR * (-2.0 - uR / aR)), pR * uRminus * (1.0 / uR))
    p_half = pLplus + pRminus
    mass = uLplus * rL + uRminus * rR
    mx = uLplus * rL * uL + uRminus * rR * uR + p_half
    my = uLplus * rL * vL + uRminus * rR * vR
    mz = uLplus * rL * wL + uRminus * rR * wR
    en = uLplus * rL * HL + uRminus * rR * HR
    return mass, mx, my, (mz if three_d else np.zeros_like(mz)), en


def flux_rusanov(L, R_, vl, vr, gm, three_d):
    """Rusanov::compute_flux (flux_calculators/rusanov.cpp:8-53)"""
    rL, pL, uL, vL, wL, eL, aL, keL = _unpack(L, vl, gm)
    rR, pR, uR, vR, wR, eR, aR, keR = _unpack(R_, vr, gm)
    ruL, ruR = rL * uL, rR * uR
    HL, HR = eL + pL / rL + keL, eR + pR / rR + keR
    S = _max(np.abs(uL - aL), np.abs(uR - aR))
    S = _max(S, np.abs(uL + aL))
    S = _max(S, np.abs(uR + aR))
    mass = 0.5 * (ruL + ruR) - 0.5 * S * (rR - rL)
    mx = 0.5 * (ruL * uL + pL + ruR * uR + pR) - 0.5 * S * (rR * uR - rL * uL)
    my = 0.5 * (ruL * vL + ruR * vR) - 0.5 * S * (rR * vR - rL * vL)
    mz = 0.5 * (ruL * wL + ruR * wR) - 0.5 * S * (rR * wR - rL * wL)
    en = 0.5 * (ruL * HL + ruR * HR) - 0.5 * S * (rR * HR - rL * HL)
    return mass, mx, my, (mz if three_d else np.zeros_like(mz)), en


def flux_ldfss2(L, R_, vl, vr, gm, three_d, delta=2.0):
    """Ldfss2::compute_flux (flux_calculators/ldfss2.cpp:9-90)"""
    rL, pL, uL, vL, wL, eL, aL, keL = _unpack(L, vl, gm)
    rR, pR, uR, vR, wR, eR, aR, keR = _unpack(R_, vr, gm)
    HL, HR = eL + pL / rL + keL, eR + pR / rR + keR
    am = 0.5 * (aL + aR)
    ML, MR = uL / am, uR / am
    MpL = 0.25 * (ML + 1.0) * (ML + 1.0)
    MmR = -0.25 * (MR - 1.0) * (MR - 1.0)
    alphaL = 0.5 * (1.0 + np.copysign(1.0, ML))
    alphaR = 0.5 * (1.0 - np.copysign(1.0, MR))
    zero = np.zeros_like(ML)
    betaL = -_max(zero, 1.0 - np.floor(np.abs(ML)))
    betaR = -_max(zero, 1.0 - np.floor(np.abs(MR)))
    PL = 0.25 * (ML + 1.0) * (ML + 1.0) * (2.0 - ML)
    PR = 0.25 * (MR - 1.0) * (MR - 1.0) * (2.0 + MR)
    DL = alphaL * (1.0 + betaL) - betaL * PL
    DR = alphaR * (1.0 + betaR) - betaR * PR
    Mhalf = 0.25 * betaL * betaR * np.power(np.sqrt(0.5 * (ML * ML + MR * MR)) - 1.0, 2.0)
    MhalfL = Mhalf * (1.0 - ((pL - pR) / (pL + pR) + delta * (np.abs(pL - pR) / pL)))
    MhalfR = Mhalf * (1.0 + ((pL - pR) / (pL + pR) - delta * (np.abs(pL - pR) / pR)))
    CL = alphaL * (1.0 + betaL) * ML - betaL * MpL - MhalfL
    CR = alphaR * (1.0 + betaR) * MR - betaR * MmR + MhalfR
    mass = am * rL * CL + am * rR * CR
    ru2_half = am * rL * CL * uL + am * rR * CR * uR
    p_half = DL * pL + DR * pR
    mx = ru2_half + p_half
    my = am * rL * CL * vL + am * rR * CR * vR
    mz = am * rL * CL * wL + am * rR * CR * wR
    en = am * rL * CL * HL + am * rR * CR * HR
    return mass, mx, my, (mz if three_d else np.zeros_like(mz)), en


FLUX = {"ausmdv": flux_ausmdv, "hanel": flux_hanel, "rusanov": flux_rusanov, "ldfss2": flux_ldfss2}


def compute_dudt(g, config, prim, allow_reconstruction=True, vertex_vel=None, cq=None):
    """FiniteVolume<real>::compute_dudt (finite_volume.cpp:57-96).  With vertex_vel (n_vertices, 3) and cq
    (conserved quantities of prim) it is the moving-grid overload: interface velocities (grid.h:571-592),
    interface-relative fluxes with the inertial-frame correction (convective_flux.cpp:105-140,166-176) and the
    geometric conservation law (finite_volume.cpp:99-128); inter["face_vel"] is GridBlock::face_vel() afterwards.
    prim: (n_cells + n_ghost, 7) rho p T e vx vy vz.  -> dudt (n_cells, dim + 2), prim after the ghost fill,
    dict of intermediates (grad_* (n_cells, 3), lim_* (n_cells,))."""
    cf = config["convective_flux"]
    if cf["flux_calculator"]["type"] not in FLUX:
        raise NotImplementedError("flux calculator " + cf["flux_calculator"]["type"])
    fs = apply_pre_reconstruction(g, config, prim.copy() if isinstance(prim, Dual) else np.array(prim, dtype=np.float64))
    rag = _Ragged(g)
    order2 = allow_reconstruction and int(cf["reconstruction_order"]) == 2
    with np.errstate(all="ignore"):
        L, R_, inter = reconstruct(g, config, fs, rag, order2)
        nrm, t1, t2 = np.asarray(g.face_norm), np.asarray(g.face_tan1), np.asarray(g.face_tan2)
        moving = vertex_vel is not None
        face_vel = np.zeros((g.n_faces, 3))
        if moving:                                               # GridBlock::compute_face_vel (grid.h:571-592)
            fv_ids, nv = np.asarray(g.face_vertices), np.asarray(g.face_nv)
            acc = np.zeros((g.n_faces, 3))
            for k in range(4):
                mk = (nv > k)[:, None]
                acc = np.where(mk, acc + vertex_vel[np.maximum(fv_ids[:, k], 0)], acc)
            face_vel = acc / nv[:, None].astype(np.float64)
            L, R_ = L.copy(), R_.copy()
            L[:, 4:7] = L[:, 4:7] - face_vel                     # subtract (convective_flux.cpp:105-108)
            R_[:, 4:7] = R_[:, 4:7] - face_vel
        vl, vr = _to_local(L[:, 4:7], nrm, t1, t2), _to_local(R_[:, 4:7], nrm, t1, t2)        # convective_flux.cpp:110-113
        if moving:
            lx, ly, lz = _to_local(face_vel, nrm, t1, t2)        # :114-117
        three_d = g.dim == 3
        fc = cf["flux_calculator"]
        extra = {"delta": float(fc.get("delta", 2.0))} if fc["type"] == "ldfss2" else {}
        mass, px, py, pz, en = FLUX[fc["type"]](L, R_, vl, vr, config["gas_model"], three_d, **extra)
        if moving:                                               # flux::grid_motion (convective_flux.cpp:123-141)
            v_sqr = lx * lx + ly * ly + lz * lz
            en = en + (0.5 * mass * v_sqr + px * lx + py * ly)
            if three_d:
                en = en + pz * lz
            px = px + lx * mass
            py = py + ly * mass
            if three_d:
                pz = pz + lz * mass
            # the interface velocity goes back to the global frame through the same rotation (:166-176)
            face_vel = np.stack([lx * nrm[:, 0] + ly * t1[:, 0] + lz * t2[:, 0],
                                 lx * nrm[:, 1] + ly * t1[:, 1] + lz * t2[:, 1],
                                 lx * nrm[:, 2] + ly * t1[:, 2] + lz * t2[:, 2]], axis=1)
            inter["face_vel"] = face_vel
        # flux::transform_to_global (convective_flux.cpp:148-164)
        mx = px * nrm[:, 0] + py * t1[:, 0] + pz * t2[:, 0]
        my = px * nrm[:, 1] + py * t1[:, 1] + pz * t2[:, 1]
        mz = px * nrm[:, 2] + py * t1[:, 2] + pz * t2[:, 2]
        # apply_post_convective_flux_bc (finite_volume.cpp:201-210): ConstantFlux::apply (boundary.cpp:405-444) on a
        # static grid, its `vy * n.x` in the energy flux included
        for tag in sorted(g.tags):
            for act in config["grid"]["boundaries"][tag].get("post_convective_flux", []):
                if act["type"] != "constant_flux":
                    raise NotImplementedError("flux action " + act["type"])
                s_, gm = act["flow_state"], config["gas_model"]
                f = np.asarray(g.tags[tag]["faces"], np.int64)
                nx, ny, nz = nrm[f, 0], nrm[f, 1], nrm[f, 2]
                u_ = float(gm["Cv"]) * s_["T"]
                mf = s_["rho"] * ((s_["vx"] - face_vel[f, 0]) * nx + (s_["vy"] - face_vel[f, 1]) * ny +
                                  (s_["vz"] - face_vel[f, 2]) * nz)
                mass[f] = mf
                mx[f] = s_["p"] * nx + s_["vx"] * mf
                my[f] = s_["p"] * ny + s_["vy"] * mf
                mz[f] = s_["p"] * nz + s_["vz"] * mf
                en[f] = mf * (u_ + 0.5 * (s_["vx"] * s_["vx"] + s_["vy"] * s_["vy"] + s_["vz"] * s_["vz"])) + \
                    s_["p"] * (s_["vx"] * nx + s_["vy"] * nx + s_["vz"] * nz)
        if config["viscous_flux"]["enabled"]:                    # finite_volume.cpp:73-77
            fs = apply_pre_reconstruction(g, config, fs, "pre_viscous_grad")
            by_name = {"mx": mx, "my": my, "mz": mz, "en": en}
            vgrads, face_fs = viscous_flux(g, config, fs, rag, by_name)
            # cell_grad_ is shared: the viscous gradients overwrite the convective ones (finite_volume.h:116-141)
            for name, gr in vgrads.items():
                inter["grad_" + name] = gr
            inter["face_fs"] = face_fs
        comps = [mass, mx, my] + ([mz] if three_d else []) + [en]
        # flux_surface_integral (finite_volume.cpp:221-256): in cell-face order
        n = g.n_cells
        acc = [np.zeros(n) for _ in comps]
        area_all = np.asarray(g.face_area)
        for j in range(rag.maxf):
            m, f = rag.mask[:, j], rag.face[:, j]
            area = -area_all[f] * rag.out[:, j]
            for k, comp in enumerate(comps):
                acc[k] = np.where(m, acc[k] + comp[f] * area, acc[k])
        vol = np.asarray(g.cell_volume)
        dudt = np.stack([a / as_T(vol, a) for a in acc], axis=1)
        if moving:                                               # apply_geometric_conservation_law (finite_volume.cpp:99-128)
            dVdt = np.zeros(n)
            for j in range(rag.maxf):
                m, f = rag.mask[:, j], rag.face[:, j]
                area = area_all[f] * rag.out[:, j]
                dVdt = np.where(m, dVdt + (face_vel[f, 0] * nrm[f, 0] + face_vel[f, 1] * nrm[f, 1] +
                                           face_vel[f, 2] * nrm[f, 2]) * area, dVdt)
            dudt = dudt - cq[:n] * dVdt[:, None] / as_T(vol[:, None], dudt)
    return dudt, fs, inter


def _viscosity(T, tp):
    """ViscosityModel::viscosity (gas/transport_properties.h:29-42): Sutherland"""
    v = tp["viscosity"]
    mu0, T0, Ts = float(v["mu_0"]), float(v["T_0"]), float(v["T_s"])
    return mu0 * np.power(T / as_T(T0, T), 3.0 / 2.0) * (T0 + Ts) / (T + Ts)


def viscous_flux(g, config, fs, rag, comps_by_name):
    """ViscousFlux::compute_viscous_flux (finite_volume/viscous_flux.cpp:122-184) with
    compute_viscous_properties_at_faces (:83-107) and the Hasselbacher average (:37-80): subtracts the viscous
    flux in place; -> (viscous cell gradients by name, face_fs T and rho for estimate_dt)"""
    n = g.n_cells
    gm, tp = config["gas_model"], config["transport_properties"]
    r = wls_r(g, rag)
    grads = {name: wls_gradient(g, rag, r, fs[:, col]) for name, col in (("T", 2), ("vx", 4), ("vy", 5), ("vz", 6))}
    left, right = np.asarray(g.face_left), np.asarray(g.face_right)
    lv, rv = left < n, right < n
    both = lv & rv
    interior = np.where(lv, left, right)
    ls, rs = np.where(both, left, 0), np.where(both, right, 0)
    c, nrm = np.asarray(g.cell_centroid), np.asarray(g.face_norm)
    ex, ey, ez = c[right, 0] - c[left, 0], c[right, 1] - c[left, 1], c[right, 2] - c[left, 2]
    len_e = np.sqrt(ex * ex + ey * ey + ez * ez)
    ehx, ehy, ehz = ex / len_e, ey / len_e, ez / len_e
    nx, ny, nz = nrm[:, 0], nrm[:, 1], nrm[:, 2]
    ehat_dot_n = ehx * nx + ehy * ny + ehz * nz
    face_grad = {}
    for name, col in (("vx", 4), ("vy", 5), ("vz", 6), ("T", 2)):
        gr = grads[name]
        ax, ay, az = 0.5 * (gr[ls, 0] + gr[rs, 0]), 0.5 * (gr[ls, 1] + gr[rs, 1]), 0.5 * (gr[ls, 2] + gr[rs, 2])
        corr = (ax * ehx + ay * ehy + az * ehz) - (fs[right, col] - fs[left, col]) / as_T(len_e, ax)
        den = as_T(ehat_dot_n, ax)
        hx, hy, hz = ax - corr * nx / den, ay - corr * ny / den, az - corr * nz / den
        gi = gr[np.where(both, 0, interior)]
        face_grad[name] = (np.where(both, hx, gi[:, 0]), np.where(both, hy, gi[:, 1]), np.where(both, hz, gi[:, 2]))
    # average_flow_states_pT + update_thermo_from_pT (gas/flow_state.h:89-91, gas_state.h:108-113)
    p = 0.5 * (fs[left, 1] + fs[right, 1])
    T = 0.5 * (fs[left, 2] + fs[right, 2])
    rho = p / (float(gm["R"]) * T)
    vx, vy, vz = (0.5 * (fs[left, k] + fs[right, k]) for k in (4, 5, 6))
    mu = _viscosity(T, tp)
    k = float(gm["Cp"]) * mu / as_T(float(tp["thermal_conductivity"]["Pr"]), mu)
    lam = -2.0 / 3.0 * mu
    gvx, gvy, gvz, gT = face_grad["vx"], face_grad["vy"], face_grad["vz"], face_grad["T"]
    bulk = lam * (gvx[0] + gvy[1] + gvz[2])
    tau_xx = 2.0 * mu * gvx[0] + bulk
    tau_yy = 2.0 * mu * gvy[1] + bulk
    tau_zz = 2.0 * mu * gvz[2] + bulk
    tau_xy = mu * (gvx[1] + gvy[0])
    tau_xz = mu * (gvx[2] + gvz[0])
    tau_yz = mu * (gvy[2] + gvz[1])
    theta_x = vx * tau_xx + vy * tau_xy + vz * tau_xz + k * gT[0]
    theta_y = vx * tau_xy + vy * tau_yy + vz * tau_yz + k * gT[1]
    theta_z = vx * tau_xz + vy * tau_yz + vz * tau_zz + k * gT[2]
    comps_by_name["mx"] -= nx * tau_xx + ny * tau_xy + nz * tau_xz
    comps_by_name["my"] -= nx * tau_xy + ny * tau_yy + nz * tau_yz
    if g.dim == 3:
        comps_by_name["mz"] -= nx * tau_xz + ny * tau_yz + nz * tau_zz
    comps_by_name["en"] -= nx * theta_x + ny * theta_y + nz * theta_z
    return grads, (T, rho)


def estimate_dt(g, config, fs, face_fs=None):
    """FiniteVolume::estimate_dt (finite_volume.cpp:131-184), inviscid: min over the valid cells of
    V / sum_f (|v.n| + a) A, summed in cell-face order"""
    n = g.n_cells
    gm = config["gas_model"]
    rag = _Ragged(g)
    nrm, area = np.asarray(g.face_norm), np.asarray(g.face_area)
    vx, vy, vz = fs[:n, 4], fs[:n, 5], fs[:n, 6]
    a = np.sqrt(float(gm["gamma"]) * float(gm["R"]) * fs[:n, 2])
    sr, sr_v = np.zeros(n), np.zeros(n)
    vol = np.asarray(g.cell_volume)
    viscous = bool(config["viscous_flux"]["enabled"])
    if viscous and face_fs is None:
        raise ValueError("viscous estimate_dt reads the face states of the last viscous flux evaluation")
    with np.errstate(all="ignore"):
        for j in range(rag.maxf):
            m, f = rag.mask[:, j], rag.face[:, j]
            dot = vx * nrm[f, 0] + vy * nrm[f, 1] + vz * nrm[f, 2]
            sr = np.where(m, sr + (np.abs(dot) + a) * area[f], sr)
            if viscous:                                          # finite_volume.cpp:166-174
                Tf, rhof = face_fs[0][f], face_fs[1][f]
                mu = _viscosity(Tf, config["transport_properties"])
                k = float(gm["Cp"]) * mu / float(config["transport_properties"]["thermal_conductivity"]["Pr"])
                Pr = mu * float(gm["Cp"]) / k
                tmp = (float(gm["gamma"]) / rhof) * (mu / Pr) * area[f] * area[f]
                sr_v = np.where(m, sr_v + tmp / vol, sr_v)
        local = vol / (sr + float(config["viscous_flux"]["signal_factor"]) * sr_v)
    dt = np.inf                                                  # Kokkos::Min identity; min(a, b) = (b < a) ? b : a
    for v in local:
        dt = v if v < dt else dt
    return float(dt)


def count_bad_cells(g, fs):
    """FiniteVolume::count_bad_cells (finite_volume.cpp:258-270)"""
    n = g.n_cells
    T, rho = fs[:n, 2], fs[:n, 0]
    return int(((T < 0.0) | (rho < 0.0) | np.isnan(rho) | np.isinf(rho)).sum())


def primatives_to_conserved(fs, dim):
    """primative_conserved_conversion.cpp:34-57 -> (n, dim + 2) mass, momentum, energy"""
    rho, e, vx, vy, vz = fs[:, 0], fs[:, 3], fs[:, 4], fs[:, 5], fs[:, 6]
    ke = 0.5 * (vx * vx + vy * vy + vz * vz)
    cols = [rho, rho * vx, rho * vy] + ([rho * vz] if dim == 3 else []) + [rho * (ke + e)]
    return np.stack(cols, axis=1)


def conserved_to_primatives(cq, dim, gm):
    """primative_conserved_conversion.cpp:5-26 with update_thermo_from_rhou -> (n, 7) rho p T e vx vy vz"""
    with np.errstate(all="ignore"):
        rho = cq[:, 0]
        vx, vy = cq[:, 1] / rho, cq[:, 2] / rho
        vz = cq[:, 3] / rho if dim == 3 else np.zeros_like(rho)
        ke = 0.5 * (vx * vx + vy * vy + vz * vz)
        u = cq[:, -1] / rho - ke
        T = u / as_T(float(gm["Cv"]), u)
        p = rho * float(gm["R"]) * T
    return np.stack([rho, p, T, u, vx, vy, vz], axis=1)


def l2_norms(dudt, dim):
    """ConservedQuantities::L2_norms (conserved_quantities.cpp:41-69) over the rows given, summed in index order
    (one OpenMP thread): global, mass, momentum x y z, energy"""
    acc = np.zeros(6)
    for row in dudt:
        m, mx, my = row[0], row[1], row[2]
        mz = row[3] if dim == 3 else 0.0
        en = row[-1]
        acc[1] += m * m
        acc[2] += mx * mx
        acc[3] += my * my
        acc[4] += mz * mz
        acc[5] += en * en
        acc[0] += m * m + mx * mx + my * my + mz * mz + en * en
    return np.sqrt(acc)


def _cfl(schedule, t):
    """ConstantSchedule / LinearSchedule::eval (solvers/cfl.cpp:11-40)"""
    if schedule["type"] == "constant":
        return float(schedule["value"])
    if schedule["type"] == "linear_interpolate":
        times, cfls = schedule["times"], schedule["cfls"]
        if t <= times[0]:
            return cfls[0]
        if t >= times[-1]:
            return cfls[-1]
        for i in range(len(times)):
            if times[i] <= t < times[i + 1]:
                frac = (t - times[i]) / (times[i + 1] - times[i])
                return cfls[i] + frac * (cfls[i + 1] - cfls[i])
    raise NotImplementedError("cfl schedule " + schedule["type"])


def _area_of_triangle(pos, a, b, c):
    """Ibis::area_of_triangle (grid/geom.h:30-45)"""
    ab, ac = pos[b] - pos[a], pos[c] - pos[a]
    cx = ab[:, 1] * ac[:, 2] - ab[:, 2] * ac[:, 1]
    cy = ab[:, 2] * ac[:, 0] - ab[:, 0] * ac[:, 2]
    cz = ab[:, 0] * ac[:, 1] - ab[:, 1] * ac[:, 0]
    return 0.5 * np.sqrt(cx * cx + cy * cy + cz * cz)


def compute_geometric_data(g, mesh, pos):
    """GridBlock::compute_geometric_data (grid/grid.h:169-178) after the vertices moved to pos: interface centres
    (interface.h:304-327) and areas (:257-302), cell centroids (cell.h:236-254) and volumes (cell.h:263-303,
    geom.h:67-83), ghost-cell centres (grid.h:364-403).  The interface orientations are NOT recomputed, as in
    the reference; the WLS r-factors follow from the centroids (they are recomputed on every evaluation here).
    Rewrites the arrays of g in place."""
    n = g.n_cells
    fv, nv = np.asarray(g.face_vertices), np.asarray(g.face_nv)
    acc = np.zeros((g.n_faces, 3))
    for k in range(4):
        acc = np.where((nv > k)[:, None], acc + pos[np.maximum(fv[:, k], 0)], acc)
    g.face_centre = acc / nv[:, None].astype(np.float64)
    v0, v1, v2, v3 = (np.maximum(fv[:, k], 0) for k in range(4))
    d = pos[v1] - pos[v0]
    line = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
    tri = _area_of_triangle(pos, v0, v1, v2)
    quad = _area_of_triangle(pos, v0, v1, v2) + _area_of_triangle(pos, v0, v2, v3)
    g.face_area = np.where(nv == 2, line, np.where(nv == 3, tri, quad))
    ct, cv = np.asarray(mesh.cell_types), np.asarray(mesh.cell_vertices)
    ncv = (cv >= 0).sum(axis=1)
    acc = np.zeros((n, 3))
    for k in range(8):
        acc = np.where((ncv > k)[:, None], acc + pos[np.maximum(cv[:, k], 0)], acc)
    cen = np.array(g.cell_centroid)
    cen[:n] = acc / ncv[:, None].astype(np.float64)
    rag = _Ragged(g)
    if g.dim == 2:
        a, b, c, dd = (np.maximum(cv[:, k], 0) for k in range(4))
        g.cell_volume = np.where(ct == 5, _area_of_triangle(pos, a, b, c),
                                 _area_of_triangle(pos, a, b, c) + _area_of_triangle(pos, a, c, dd))
    else:
        vol = np.zeros(n)
        fc, fn = g.face_centre, np.asarray(g.face_norm)
        for j in range(rag.maxf):
            m, f, out = rag.mask[:, j], rag.face[:, j], rag.out[:, j]
            dot = fc[f, 0] * out * fn[f, 0] + fc[f, 1] * out * fn[f, 1] + fc[f, 2] * out * fn[f, 2]
            vol = np.where(m, vol + dot * g.face_area[f], vol)
        g.cell_volume = vol / 3.0
    left, right = np.asarray(g.face_left), np.asarray(g.face_right)
    for tag in sorted(g.tags):
        f = np.asarray(g.tags[tag]["faces"], np.int64)
        l, r = left[f], right[f]
        valid, ghost = np.where(l < n, l, r), np.where(l < n, r, l)
        dxyz = g.face_centre[f] - cen[valid]
        cen[ghost] = g.face_centre[f] + dxyz
    g.cell_centroid = cen


def runge_kutta(g, config, prim0, mesh=None):
    """Solver::solve (solvers/solver.cpp:18-64) driving RungeKutta::initialise / estimate_dt / take_step / stop_now
    (solvers/runge_kutta.cpp:82-119,131-203,258-262), time-based plotting off.  The stage arrays k_[i] are ONE
    array in the reference (runge_kutta.cpp:49-50 copies one view handle into every slot), so every k_[j] below is
    the most recent evaluation.  With grid.motion enabled (mesh = the su2.Su2Mesh of g; rigid_body_translation
    driver) the vertices move: `init_vertex_pos_ = grid_.vertices().positions()` (:136) copies a view HANDLE and
    vertex_vel_ is one array in every stage slot (:64-65), so every add_scaled_vector (:158,172,189) displaces
    the grid's own vertices cumulatively, and the geometry is recomputed before every later stage and at the end
    of the step (:178-180,194-196).  g is modified in place then.  -> (final primitive state, t, steps taken)"""
    import sys
    sc, gm, dim = config["solver"], config["gas_model"], g.dim
    a_rows, b = sc["butcher_tableau"]["a"], sc["butcher_tableau"]["b"]
    n_stages = len(b)
    a = lambda i, j: float(a_rows[i - 1][j])                                        # noqa: E731
    flow = np.array(prim0, dtype=np.float64)
    cq = primatives_to_conserved(flow, dim)
    dt = float(sc["dt_init"]) if sc["dt_init"] > 0 else sys.float_info.max
    t, max_time, max_step = 0.0, float(sc["max_time"]), int(sc["max_step"])
    viscous = bool(config["viscous_flux"]["enabled"])
    motion = config["grid"].get("motion", {})
    moving = bool(motion.get("enabled"))
    if moving:
        if motion["type"] != "rigid_body_translation":
            raise NotImplementedError("grid motion driver " + motion["type"])
        pos = np.array(mesh.vertices, dtype=np.float64)
        vv = motion["velocity"]
        vertex_vel = np.tile([float(vv["x"]), float(vv["y"]), float(vv["z"])], (len(pos), 1))

    def evaluate(fs, cq_of_fs=None):
        """function_eval_: dudt over all rows (ghost rows stay zero), the ghost-filled state, face_fs"""
        if moving:
            d, fs_out, inter = compute_dudt(g, config, fs, True, vertex_vel=vertex_vel, cq=cq_of_fs)
        else:
            d, fs_out, inter = compute_dudt(g, config, fs, True)
        full = np.zeros((len(fs), dim + 2))
        full[:g.n_cells] = d
        return full, fs_out, inter.get("face_fs")

    k, flow, face_fs = evaluate(flow, cq)                                           # initialise(): :93
    steps = 0
    for step in range(max_step):
        k, flow, face_fs = evaluate(flow, cq)                                       # take_step: :143
        stable = estimate_dt(g, config, flow, face_fs if viscous else None)        # :108-119
        dt_startup = min(_cfl(sc["cfl"], t) * stable, 1.5 * dt)
        dt = min(dt_startup, max_time - t)
        for i in range(1, n_stages):
            k_tmp = cq + k * (a(i, 0) * dt)                                         # :155
            if moving:
                pos = pos + vertex_vel * (a(i, 0) * dt)                             # :157-160
            for j in range(1, i):
                if a(i, j) < 1e-14:
                    continue
                k_tmp = k_tmp + k * (a(i, j) * dt)                                  # :169
                if moving:
                    pos = pos + vertex_vel * (a(i, j) * dt)                         # :170-173
            with np.errstate(all="ignore"):
                flow_tmp = conserved_to_primatives(k_tmp, dim, gm)                  # :177
            if moving:
                compute_geometric_data(g, mesh, pos)                                # :178-180
            k, _, face_fs = evaluate(flow_tmp, k_tmp)                               # :181 (flow_tmp's ghosts are its own)
        for i in range(n_stages):
            cq = cq + k * (float(b[i]) * dt)                                        # :185-186
            if moving:
                pos = pos + vertex_vel * (float(b[i]) * dt)                         # :188-190
        with np.errstate(all="ignore"):
            flow = conserved_to_primatives(cq, dim, gm)                             # :193
        if moving:
            compute_geometric_data(g, mesh, pos)                                    # :194-196
        t += dt
        steps += 1
        if count_bad_cells(g, flow) > 0:                                            # solver.cpp:38-43
            break
        if step >= max_step - 1 or t >= max_time:                                   # stop_now
            break
    return flow, t, steps


def _seqsum(a):
    """Kokkos::parallel_reduce(Sum) with one thread: index order"""
    return float(np.cumsum(a)[-1]) if len(a) else 0.0


def gmres(matvec, rhs, x0, max_iters, tol, precond=None):
    """Gmres::solve (linear_algebra/gmres.cpp:119-182) with compute_r0_ (:11-20) and
    apply_rotations_to_hessenberg_ (:22-74: full (j+2)^2 matrix products on the host, kept as such); dot / norm2
    summed in index order.  With precond (v -> z, an inner Gmres on the preconditioner system) it is
    FGmres::solve (:234-303): w = A z_j, and the update is built from the z_j.
    -> (x, success, n_iters, relative residual)"""
    m = min(max_iters, len(rhs))
    Q0, Q1, Om = np.eye(m + 1), np.eye(m + 1), np.eye(m + 1)
    H0, H1 = np.zeros((m + 1, m + 1)), np.zeros((m + 1, m + 1))
    g0, g1 = np.zeros(m + 1), np.zeros(m + 1)
    x = np.array(x0, dtype=np.float64)
    r0 = rhs - matvec(x)
    beta = np.sqrt(_seqsum(r0 * r0))
    if beta < tol:
        return x, True, 0, beta
    g0[0] = beta
    v = r0 * (1.0 / beta)
    V = [v.copy()]
    Z = []
    success, n_iters, rel = False, 0, beta

    def gemm(A, B):                                              # dense_linear_algebra.h:262-279
        out = np.zeros((A.shape[0], B.shape[1]))
        for r in range(A.shape[0]):
            for c in range(B.shape[1]):
                dot = 0.0
                for i in range(A.shape[1]):
                    dot += A[r, i] * B[i, c]
                out[r, c] = dot
        return out

    def gemv(A, xv):                                             # :244-260
        out = np.zeros(A.shape[0])
        for r in range(A.shape[0]):
            dot = 0.0
            for c in range(A.shape[1]):
                dot += A[r, c] * xv[c]
            out[r] = dot
        return out

    for j in range(m):
        if precond is not None:
            z = precond(v)
            Z.append(z.copy())
            w = matvec(z)
        else:
            w = matvec(v)
        for i in range(j + 1):
            H0[i, j] = _seqsum(w * V[i])
            w = w + V[i] * (-H0[i, j])
        H0[j + 1, j] = np.sqrt(_seqsum(w * w))
        v = w * (1.0 / H0[j + 1, j])
        V.append(v.copy())
        if j != 0:
            H0[:j + 1, j] = gemv(Q0[:j + 1, :j + 1], H0[:j + 1, j].copy())
        Om[...] = np.eye(m + 1)
        denom = np.sqrt(H0[j, j] * H0[j, j] + H0[j + 1, j] * H0[j + 1, j])
        si, ci = H0[j + 1, j] / denom, H0[j, j] / denom
        Om[j, j], Om[j, j + 1], Om[j + 1, j], Om[j + 1, j + 1] = ci, si, -si, ci
        k = j + 2
        H1[:k, :k] = gemm(Om[:k, :k], H0[:k, :k])
        g1[:k] = gemv(Om[:k, :k], g0[:k])
        Q1[:k, :k] = Om[:k, :k] if j == 0 else gemm(Om[:k, :k], Q0[:k, :k])
        g0[:k] = g1[:k]
        Q0[:k, :k] = Q1[:k, :k]
        H0[:k, :k] = H1[:k, :k]
        residual = abs(g0[j + 1])
        rel, n_iters = residual / beta, j + 1
        if residual < tol * beta:
            success = True
            break
    nv = n_iters
    ym = np.zeros(nv)                                            # upper_triangular_solve (:280-297)
    ym[nv - 1] = g0[nv - 1] / H0[nv - 1, nv - 1]
    for i in range(nv - 2, -1, -1):
        s = g0[i]
        for jj in range(i + 1, nv):
            s -= H0[i, jj] * ym[jj]
        ym[i] = s / H0[i, i]
    w = np.zeros_like(x)
    basis = V if precond is None else Z
    for c in range(nv):                                          # gemv(V | Z, ym, w): row-wise, columns in order
        w = w + basis[c] * ym[c]
    return x + w * 1.0, success, n_iters, rel


def steady_state(g, config, prim0):
    """Solver::solve driving SteadyState::take_step -> Jfnk::step (solvers/steady_state.cpp:47-170,222-300,
    solvers/jfnk.cpp:25-93) with the unpreconditioned Gmres: the matrix-vector product is one dual-number residual
    evaluation.  -> (final primitive state (real), steps taken, relative global residual)"""
    sc = config["solver"]
    ls = sc["linear_solver"]
    if ls["type"] not in ("gmres", "fgmres"):
        raise NotImplementedError("linear solver " + ls["type"])
    gm, dim = config["gas_model"], g.dim
    n, nc = g.n_cells, g.dim + 2
    viscous = bool(config["viscous_flux"]["enabled"])
    fs = Dual(np.array(prim0, dtype=np.float64))
    cq = primatives_to_conserved(fs, dim)

    def residual(fs_in):
        d, fs_out, inter = compute_dudt(g, config, fs_in, True)
        ff = inter.get("face_fs")
        return d, fs_out, (None if ff is None else (ff[0].re, ff[1].re))

    def norms_of(d):
        return l2_norms(d.re, dim)

    res, fs, face_fs = residual(fs)                              # Jfnk::initialise: eval_rhs
    rhs = res.re.reshape(-1).copy()
    norms = norms_of(res)
    init = norms.copy()
    cflc = sc["cfl"]
    prev_cfl, prev_res = float(cflc.get("start_cfl", 0.0)), 1.0
    steps = 0
    for step in range(int(sc["max_steps"])):
        stable = estimate_dt(g, config, fs.re, face_fs if viscous else None)
        if cflc["type"] == "residual_based":                     # ResidualBasedCfl::eval (cfl.cpp:57-70)
            t = norms[0] / init[0]
            if t > float(cflc["growth_threshold"]):
                prev_res, cfl = t, prev_cfl
            else:
                new = prev_cfl * np.power(prev_res / t, float(cflc["power"]))
                new = float(cflc["max_cfl"]) if float(cflc["max_cfl"]) < new else new          # Ibis::min(new, max_cfl)
                new = new if float(cflc["start_cfl"]) < new else float(cflc["start_cfl"])      # Ibis::max(start_cfl, new)
                prev_res, prev_cfl, cfl = t, new, new
        else:
            cfl = _cfl(cflc, float(step))
        dt_star = cfl * stable

        def matvec(vec, allow=True):                             # SteadyStateLinearisation::matrix_vector_product
            cq_tmp = Dual(np.zeros(cq.re.shape))
            cq_tmp.re[:n] = cq.re[:n]
            cq_tmp.du[:n] = vec.reshape(n, nc)
            with np.errstate(all="ignore"):
                fs_tmp = conserved_to_primatives(cq_tmp, dim, gm)
            d = compute_dudt(g, config, fs_tmp, allow)[0]
            return 1 / dt_star * vec - d.du.reshape(-1)

        if ls["type"] == "fgmres":
            # the preconditioner is the same linearisation without reconstruction (steady_state.cpp:47-50), solved
            # by an inner Gmres from a zero guess (gmres.cpp:258-262)
            def precond(v):
                return gmres(lambda x: matvec(x, False), v, np.zeros(n * nc), int(ls["max_preconditioner_iters"]),
                             float(ls["preconditioner_tolerance"]))[0]
            dU, ok, iters, rel = gmres(matvec, rhs, np.zeros(n * nc), int(ls["max_iters"]), float(ls["tolerance"]),
                                       precond)
        else:
            dU, ok, iters, rel = gmres(matvec, rhs, np.zeros(n * nc), int(ls["max_iters"]), float(ls["tol"]))
        cq.re[:n] += dU.reshape(n, nc)                           # Jfnk::apply_update_
        cq.du[:n] = 0.0
        with np.errstate(all="ignore"):
            fs = conserved_to_primatives(cq, dim, gm)
        res, fs, face_fs = residual(fs)
        rhs = res.re.reshape(-1).copy()
        norms = norms_of(res)
        steps += 1
        if count_bad_cells(g, fs.re) > 0:
            break
        if step >= int(sc["max_steps"]) - 1 or norms[0] / init[0] < float(sc["tolerance"]):
            break
    return fs.re, steps, norms[0] / init[0]
