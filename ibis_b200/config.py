"""The reference's ``config/config.json`` schema, unchanged.

``ibis prep`` turns a user's ``job.py`` into ``config/config.json``
(/root/reference/src/ibis/ibis/commands/prep/prep.py); the C++ solvers read that JSON in
their constructors.  ibis_b200 consumes the same JSON.  The helpers here build such
dictionaries directly (same keys, same defaults as /root/reference/share/defaults/*.json)
for tests and benchmarks, and provide the boundary-condition factories of prep.py:541-614.
"""
import copy

AIR_R = 287.0

# share/defaults/runge_kutta.json + prep.py:845-859
BUTCHER = {
    "euler": {"a": [[]], "b": [1.0], "c": []},
    "midpoint": {"a": [[0.5]], "b": [0.0, 1.0], "c": [0.5]},
    "rk3": {"a": [[0.5], [-1.0, 2.0]], "b": [1 / 6, 2 / 3, 1 / 6], "c": [0.5, 1.0]},
    "ssp-rk3": {"a": [[1], [0.25, 0.25]], "b": [1 / 6, 1 / 6, 2 / 3], "c": [1, 0.5]},
    "rk4": {"a": [[0.5], [0.0, 0.5], [0.0, 0.0, 1.0]], "b": [1 / 6, 1 / 3, 1 / 3, 1 / 6],
            "c": [0.5, 0.5, 1.0]},
}


def ideal_gas(R=AIR_R):
    # gas/gas_model.h:51-56: Cv = 5/2 R, Cp = 7/2 R, gamma = Cp/Cv
    cv, cp = 5.0 / 2.0 * R, 7.0 / 2.0 * R
    return {"type": "ideal_gas", "R": R, "Cv": cv, "Cp": cp, "gamma": cp / cv}


def air_transport():
    # share/species_database/air.json
    return {"viscosity": {"type": "sutherland", "mu_0": 1.716e-05, "T_0": 273.0, "T_s": 110.0},
            "thermal_conductivity": {"type": "constant_prandtl_number", "Pr": 0.72}}


def flow_state(rho=None, T=None, p=None, vx=0.0, vy=0.0, vz=0.0, R=AIR_R):
    """FlowState.as_dict() of ibis_py_utils.py:34-39; two of rho/T/p given."""
    cv = 5.0 / 2.0 * R
    if rho is not None and T is not None:
        p = rho * R * T                       # update_thermo_from_rhoT (gas_model.h:24)
    elif p is not None and T is not None:
        rho = p / (R * T)                     # update_thermo_from_pT   (gas_model.h:14)
    elif rho is not None and p is not None:
        T = p / (rho * R)                     # update_thermo_from_rhop (gas_model.h:19)
    else:
        raise ValueError("need two of rho, T, p")
    return {"p": p, "T": T, "rho": rho, "energy": cv * T, "vx": vx, "vy": vy, "vz": vz}


def _bc(pre, post=(), visc=()):
    return {"pre_reconstruction": list(pre), "post_convective_flux": list(post),
            "pre_viscous_grad": list(visc), "ghost_cells": True}


def supersonic_inflow(fs):
    return _bc([{"type": "flow_state_copy", "flow_state": dict(fs)}])


def supersonic_outflow():
    return _bc([{"type": "internal_copy"}])


def slip_wall():
    return _bc([{"type": "internal_copy_reflect_normal"}])


def adiabatic_no_slip_wall():
    return _bc([{"type": "internal_copy_reflect_normal"}],
               visc=[{"type": "internal_vel_copy_reflect"}])


def fixed_temperature_no_slip_wall(temperature):
    return _bc([{"type": "internal_copy_reflect_normal"}],
               visc=[{"type": "internal_vel_copy_reflect"},
                     {"type": "fix_temperature", "temperature": temperature}])


def subsonic_inflow(fs):
    return _bc([{"type": "subsonic_inflow", "flow_state": dict(fs)}])


def subsonic_outflow(pressure):
    return _bc([{"type": "subsonic_outflow", "pressure": pressure}])


def boundary_layer_inflow(height, velocity_profile, temperature_profile, pressure):
    # prep.py:549-557
    return _bc([{"type": "boundary_layer_profile",
                 "profile": {"height": list(height), "v": list(velocity_profile),
                             "T": list(temperature_profile), "p": pressure}}])


def constant_flux(fs):
    # prep.py:609-617 (also exported there as bow_shock_fit)
    return _bc([{"type": "flow_state_copy", "flow_state": dict(fs)}],
               post=[{"type": "constant_flux", "flow_state": dict(fs)}])


def rigid_body_translation(x=0.0, y=0.0, z=0.0):
    # prep.py:309-328 (RigidBodyTranslation.as_dict)
    return {"enabled": True, "type": "rigid_body_translation", "velocity": {"x": x, "y": y, "z": z}}


def runge_kutta(method="ssp-rk3", cfl=0.5, max_step=100, max_time=1.0, dt_init=-1.0,
                print_frequency=1000, plot_frequency=-1, plot_every_n_steps=-1,
                residual_frequency=-1, residuals_every_n_steps=-1):
    cfl_d = cfl if isinstance(cfl, dict) else {"type": "constant", "value": cfl}
    return {"name": "runge_kutta", "cfl": cfl_d, "max_time": max_time, "max_step": max_step,
            "print_frequency": print_frequency, "plot_frequency": plot_frequency,
            "plot_every_n_steps": plot_every_n_steps, "dt_init": dt_init,
            "butcher_tableau": copy.deepcopy(BUTCHER[method]),
            "residual_frequency": residual_frequency,
            "residuals_every_n_steps": residuals_every_n_steps}


def steady_state(cfl=10.0, max_steps=50, tolerance=1e-10, linear_solver=None,
                 print_frequency=1000, plot_frequency=100000, diagnostics_frequency=-1):
    cfl_d = cfl if isinstance(cfl, dict) else {"type": "constant", "value": cfl}
    if linear_solver is None:
        linear_solver = {"type": "gmres", "max_iters": 50, "tol": 1e-2}
    return {"name": "steady_state", "cfl": cfl_d, "max_steps": max_steps,
            "print_frequency": print_frequency, "plot_frequency": plot_frequency,
            "diagnostics_frequency": diagnostics_frequency, "tolerance": tolerance,
            "linear_solver": dict(linear_solver)}


def make_config(dim, boundaries, flux_calculator="hanel", reconstruction_order=2,
                limiter="barth_jespersen", epsilon=1e-25, thermo_interpolator="rho_u",
                viscous=False, signal_factor=4.0, solver=None, R=AIR_R, ldfss_delta=2.0, motion=None):
    fc = {"type": flux_calculator}
    if flux_calculator == "ldfss2":
        fc["delta"] = ldfss_delta
    lim = {"type": limiter}
    if limiter == "barth_jespersen":
        lim["epsilon"] = epsilon
    return {
        "convective_flux": {"flux_calculator": fc, "reconstruction_order": reconstruction_order,
                            "limiter": lim, "thermo_interpolator": thermo_interpolator},
        "viscous_flux": {"enabled": bool(viscous), "signal_factor": signal_factor},
        "solver": copy.deepcopy(solver) if solver is not None else runge_kutta(),
        "grid": {"boundaries": copy.deepcopy(boundaries), "dimensions": dim,
                 "motion": copy.deepcopy(motion) if motion else {"enabled": False}},
        "gas_model": ideal_gas(R),
        "transport_properties": air_transport(),
        "io": {"flow_format": "native_binary"},
    }
