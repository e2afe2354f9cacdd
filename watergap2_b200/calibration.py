"""Objective evaluation for parameter sweeps (plumbing above the C ABI).

The reference calibrates gamma of one basin by bisection, one model run per step (calibration.cpp:266-527), and
judges a run by its annual station discharge: the 1 % criterion on the summed difference (:317-330) and the
Nash-Sutcliffe coefficient over the evaluated years (:531-546), all in single precision.  With the parameter
sets as members of one context (and the members sharded over the GPUs) a whole sweep is ONE run: every member
records the daily discharge of the station cells on the device (wgk_record_cells), and the functions here turn
those records into the reference's criteria for all parameter sets at once.

parity unpinned: calibration.cpp needs the station / basin files of a calibration set-up and is not run by the
harness; the formulas follow the cited lines.
"""
import numpy as np


def annual_runoff_km3(record, days_per_year=365):
    """record [ndays, nstation] of daily discharge in km3/day (wgk_get_record) -> float32 [nyears, nstation]
    in km3/year, the unit and precision of calibGammaClass::setRunoff (calibration.cpp:253-256)"""
    r = np.asarray(record, np.float64)
    ny = r.shape[0] // days_per_year
    return r[:ny * days_per_year].reshape(ny, days_per_year, -1).sum(1).astype(np.float32)


def measured_km3_per_year(m3_per_s):
    """observed mean annual discharge in m3/s -> km3/year (calibration.cpp:233-235); missing years stay -99"""
    q = np.asarray(m3_per_s, np.float32)
    prod = q * np.float32(365) * np.float32(24) * np.float32(60) * np.float32(60)  # float * int stays float ...
    return np.where(q > -1, (prod.astype(np.float64) / 1000000000.0).astype(np.float32),  # ... / double literal, stored as float
                    np.float32(-99)).astype(np.float32)


def criteria(measured, simulated):
    """measured, simulated: float32 [nyears] in km3/year, measured < -1 = no observation.
    -> dict(nse, sum_of_differences, rel_difference, measured_avg, years) as calibration.cpp:300-330, 531-546"""
    f = np.float32
    m, s = np.asarray(measured, f), np.asarray(simulated, f)
    n, sod, msum = 0, f(0), f(0)
    for i in range(m.size):
        if m[i] > -1:
            n += 1
            sod = f(sod + f(s[i] - m[i]))
            msum = f(msum + m[i])
    avg = f(msum / f(n)) if n else f(np.nan)
    if m.size > 1:
        sum1, sum2 = f(0), f(0)
        for i in range(m.size):
            if m[i] > -1:
                sum1 = f(sum1 + f(m[i] - avg) * f(m[i] - avg))
                sum2 = f(sum2 + f(s[i] - m[i]) * f(s[i] - m[i]))
        nse = f((sum1 - sum2) / sum1)
    else:
        nse = f(-99)
    return {"nse": float(nse), "sum_of_differences": float(sod), "years": n, "measured_avg": float(avg),
            "rel_difference": float(abs(sod / f(f(n) * avg))) if n else float("nan")}


def correction_factor(measured_avg, sim_runoff_sum, sim_water_use_sum, sim_inflow_sum, years):
    """cell correction factor CFA of a basin whose gamma hit a limit (calibration.cpp:586-588)"""
    f = np.float32
    y = f(years)
    return float(f(f(measured_avg) + f(f(sim_water_use_sum) - f(sim_inflow_sum)) / y)
                 / f(f(f(sim_runoff_sum) + f(sim_water_use_sum) - f(sim_inflow_sum)) / y))


def sweep_criteria(model, measured, nyears, days_per_year=365):
    """criteria() of every member (parameter set) of `model` at every recorded station after `nyears` recorded years.
    measured: [nyears, nstation] km3/year.  -> list over members of list over stations of criteria dicts"""
    out = []
    for mem in range(model.nmember):
        sim = annual_runoff_km3(model.get_record(nyears * days_per_year, mem), days_per_year)
        out.append([criteria(np.asarray(measured)[:, k], sim[:, k]) for k in range(sim.shape[1])])
    return out


def best_member(crit, station=0):
    """index of the parameter set with the smallest 1 %-criterion value at `station` (what the bisection converges to)"""
    return int(np.argmin([c[station]["rel_difference"] for c in crit]))


def gather_criteria(crit, group=None):
    """all ranks' sweep_criteria lists concatenated in rank order (members are sharded contiguously, ensemble.shard_members)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return crit
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, crit, group=group)
    return [c for p in parts for c in p]
