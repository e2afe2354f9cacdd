"""Objective evaluation for parameter sweeps (plumbing above the C ABI).

The reference calibrates gamma of one basin by bisection, one model run per step (calibration.cpp:266-527), and
judges a run by its annual station discharge: the 1 % criterion on the summed difference (:317-330) and the
Nash-Sutcliffe coefficient over the evaluated years (:531-546), all in single precision.  With the parameter
sets as members of one context (and the members sharded over the GPUs) a whole sweep is ONE run: every member
records the daily discharge of the station cells on the device (wgk_record_cells), and the functions here turn
those records into the reference's criteria for all parameter sets at once.

The gamma search itself (calibGammaClass: bisection, 1 % / 10 % criteria, CFA, correction grid, CFS) is GammaCalibration
below.  Pinned against the compiled reference: `ref_harness calib` runs the reference's calibGammaClass in the calibration
loop of integrate_wghm_ on seven scenarios (tests/golden/ref_calibration.json, tests/test_calibration.py); GammaCalibration
reproduces its gamma sequences, CALIBRATION.OUT / STAT_CORR_FACTOR.OUT lines and G_CORR_FACTOR grid exactly, and criteria()
gives the same Nash-Sutcliffe / sum-of-differences values.
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


def gather_annual_runoff(annual, nsets_total, group=None, device=None):
    """annual [sets_on_this_rank, nyears, nstation] (km3/year per recorded station) of every rank -> the table of all
    `nsets_total` sets in set order (sets are sharded contiguously, ensemble.shard_members): ONE all-gather of a padded
    f64 tensor (NCCL on the GPUs, gloo in the CPU tests) instead of pickled objects."""
    import torch
    import torch.distributed as dist
    a = torch.as_tensor(np.ascontiguousarray(annual, np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return a.numpy()
    world = dist.get_world_size(group)
    cap = (nsets_total + world - 1) // world
    pad = torch.zeros((cap,) + tuple(a.shape[1:]), dtype=torch.float64)
    pad[:a.shape[0]] = a
    if device is not None:
        pad = pad.to(device)
    out = torch.empty((world * cap,) + tuple(a.shape[1:]), dtype=torch.float64, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.cpu().numpy().reshape((world, cap) + tuple(a.shape[1:]))
    base, extra = divmod(nsets_total, world)
    return np.concatenate([out[r, :base + (1 if r < extra else 0)] for r in range(world)], axis=0)


def gather_criteria(crit, group=None):
    """all ranks' sweep_criteria lists concatenated in rank order (members are sharded contiguously, ensemble.shard_members)"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return crit
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, crit, group=group)
    return [c for p in parts for c in p]


# ----------------------------------------------------------------------------------------------------------------
# The reference's gamma search itself (calibGammaClass, calibration.cpp:45-801), as host logic above the C ABI:
# a calibration run of integrate_wghm_ (integrateWGHM.cpp:213-264, 967-972, 1091-1116) is
#     cal = GammaCalibration(eval_start_year, end_year); cal.read_observed(...)
#     loop: run the evaluation years on the GPU with gamma, cal.set_runoff / set_upst_inflow / set_water_use per year,
#           gamma = cal.find_new_gamma(gamma)  until it returns -99, then one test run and cal.write_corr_factors(...)
# Everything is single precision in the reference's operand order; the CALIBRATION.OUT / STAT_CORR_FACTOR.OUT lines
# have the reference's layout.  Pinned against the compiled reference (`ref_harness calib`, tests/test_calibration.py,
# golden tests/golden/ref_calibration.json).
# ----------------------------------------------------------------------------------------------------------------
_f = np.float32


def _g(x):
    """operator<< of a float / double in the default ostream format (%g, 6 significant digits)"""
    return "%g" % float(x)


class GammaCalibration:
    GAMMA_UPPER, GAMMA_LOWER = _f(5.), _f(0.1)  # calibration.cpp:45-49

    def __init__(self, eval_start_year, end_year, station_number=0):
        self.y0, self.y1 = int(eval_start_year), int(end_year)
        n = self.y1 - self.y0 + 1
        self.years = [self.y0 + i for i in range(n)]
        self.measured = np.full(n, -99, _f)        # calibGammaClass::init (:239-249)
        self.sim_runoff = np.full(n, -99, _f)
        self.sim_inflow = np.full(n, -99, _f)
        self.sim_water_use = np.full(n, -99, _f)
        self.station_number = int(station_number)
        self.cell_corr_factor = _f(1.)
        self.cell_corr_factor_ind = 1
        self.gamma_cond = 0
        self.calib_status = 0
        # function-local statics of findNewGamma (:277-284)
        self.call_counter = 0
        self.gamma_low, self.gamma_high = _f(-99), _f(-99)
        self.sum_of_differences_old = _f(0)
        self.gamma_of_previous_run = _f(0)
        self.result_lines, self.log = [], []
        self.runoff_generated_in_basin = None

    def read_observed(self, years, m3_per_s):
        """RIVER.DAT: year, observed discharge in m3/s -> km3/year (readObservedData, :212-237)"""
        for y, q in zip(years, m3_per_s):
            self.measured[int(y) - self.y0] = measured_km3_per_year(_f(q))

    def set_runoff(self, year, value):
        self.sim_runoff[year - self.y0] = _f(value)

    def set_upst_inflow(self, year, value):
        self.sim_inflow[year - self.y0] = _f(value)

    def set_water_use(self, year, value):
        self.sim_water_use[year - self.y0] = _f(value)

    def find_new_gamma(self, gamma_old):
        """calibGammaClass::findNewGamma (:266-636) -> gamma of the next run, or -99 when the search has ended"""
        gamma_old = _f(gamma_old)
        self.call_counter += 1
        cc = self.call_counter
        UP, LO = self.GAMMA_UPPER, self.GAMMA_LOWER
        sod, msum, isum, usum, rsum, n = _f(0), _f(0), _f(0), _f(0), _f(0), 0
        for i in range(self.measured.size):  # :300-313
            if self.measured[i] > -1:
                n += 1
                sod = _f(sod + _f(self.sim_runoff[i] - self.measured[i]))
                msum = _f(msum + self.measured[i])
                rsum = _f(rsum + self.sim_runoff[i])
                isum = _f(isum + self.sim_inflow[i])
                usum = _f(usum + self.sim_water_use[i])
        avg = _f(msum / _f(n))
        crit = abs(float(_f(sod / _f(_f(n) * avg))))
        gamma = None  # the reference leaves `gamma` uninitialised on the paths that do not assign it
        half = lambda: _f((float(_f(self.gamma_low + self.gamma_high))) / 2.0)
        if crit < 0.01:  # the 1 % criterion, identical in all three branches (:317-330, 369-379, 452-462)
            gamma, self.gamma_cond, self.calib_status = _f(-99), 1, 1
        elif cc == 1:    # first call: go to a limit (:331-361)
            if sod > 0:
                if gamma_old >= UP:
                    gamma = _f(-99)
                else:
                    self.gamma_low, gamma = gamma_old, UP
            else:
                if gamma_old <= LO:
                    gamma = _f(-99)
                else:
                    self.gamma_high, gamma = gamma_old, LO
        else:            # bisection (:380-446 second call, :463-526 later calls)
            if sod > 0:
                if gamma_old == UP:
                    if cc == 2:
                        gamma, self.gamma_high = gamma_old, UP
                    else:
                        gamma, self.gamma_cond = _f(-99), 2
                elif gamma_old < UP:
                    if self.gamma_high < 0:
                        self.gamma_low, gamma = gamma_old, _f(float(gamma_old) * 2.0)
                    else:
                        self.gamma_low = gamma_old
                        gamma = half()  # second call: guarded by `gamma != gammaUpperLimit` on an uninitialised gamma (:404)
            else:
                if gamma_old == LO:
                    if cc == 2:
                        gamma, self.gamma_low = gamma_old, LO
                    else:
                        gamma, self.gamma_cond = _f(-99), 2
                elif gamma_old > LO:
                    if self.gamma_low < 0:
                        gamma, self.gamma_high = _f(float(gamma_old) / 2.0), gamma_old
                    else:
                        self.gamma_high = gamma_old
                        gamma = half()
        # Nash-Sutcliffe (:529-546)
        if self.y1 - self.y0 > 0:
            s1, s2 = _f(0), _f(0)
            for i in range(self.measured.size):
                if self.measured[i] > -1:
                    s1 = _f(s1 + _f(_f(self.measured[i] - avg) * _f(self.measured[i] - avg)))
                    s2 = _f(s2 + _f(_f(self.sim_runoff[i] - self.measured[i]) * _f(self.sim_runoff[i] - self.measured[i])))
            with np.errstate(all="ignore"):  # all observed years equal: 0 / 0 or x / 0, as in the reference
                nse = _f(_f(s1 - s2) / s1)
        else:
            nse = _f(-99)
        # 10 % criterion at a limit (:551-574)
        avg_adapt = avg
        sim_avg = _f(rsum / _f(n))
        for limit, factor, ok in ((UP, 1.1, lambda: sim_avg < avg_adapt), (LO, 0.9, lambda: sim_avg > avg_adapt)):
            if gamma_old == limit and self.gamma_cond == 2:
                avg_adapt = _f(float(avg) * factor)
                if ok():
                    self.calib_status = 2
                else:
                    self.cell_corr_factor_ind = 99
        # CFA (:583-592)
        if self.cell_corr_factor_ind == 99:
            self.cell_corr_factor = _f(_f(avg_adapt + _f(_f(usum - isum) / _f(n))) / _f(_f(_f(rsum + usum) - isum) / _f(n)))
            self.runoff_generated_in_basin = _f(_f(_f(rsum / _f(n)) + _f(usum / _f(n))) - _f(isum / _f(n)))
        # CALIBRATION.OUT line (:594-607); column 13 is uninitialised in the reference unless CFA was computed
        rgb = "?" if self.runoff_generated_in_basin is None else _g(self.runoff_generated_in_basin)
        self.result_lines.append("\t".join([_g(gamma_old), _g(nse), _g(sod), _g(self.gamma_low), _g(self.gamma_high), _g(avg_adapt), str(n),
                                            _g(_f(avg_adapt / sim_avg)), _g(_f(isum / _f(n))), _g(_f(usum / _f(n))), "0", _g(sim_avg), rgb,
                                            _g(self.cell_corr_factor)]) + "\t")
        if gamma is not None:
            if gamma > 0 and float(gamma) < 0.99 * float(LO) and cc > 2:  # :609-613
                gamma = _f(-99)
            if gamma > 0 and sod != 0 and abs(float(_f(_f(self.sum_of_differences_old / sod) - _f(1)))) < 0.0001:  # :614-619
                gamma = _f(-99)
        self.sum_of_differences_old = sod
        self.last = {"sim_avg": sim_avg, "avg_adapt": avg_adapt, "nse": nse, "sum_of_differences": sod, "years": n}
        if gamma is not None and gamma < 0 and abs(float(_f(self.cell_corr_factor - _f(1)))) > 0.01:  # :623-627
            self.calib_status = 3
            self.correction_grid_due = (sim_avg, avg_adapt)  # createCorrectionGrid(evalStartYear, end_year, sim, measured)
        self.gamma_of_previous_run = gamma_old
        return gamma if gamma is not None else _f(np.nan)

    def correction_grid(self, annual_pot_cell_runoff, sbasin, cell_corr_fact):
        """createCorrectionGrid (:730-801): annual_pot_cell_runoff float32 [nyears][ng] (the G_POT_CELL_RUNOFF_<year> grids),
        sbasin int16 [ng]; updates and returns cell_corr_fact [ng] (G_cellCorrFact) for the cells of the station's basin"""
        sim, meas = self.correction_grid_due
        a = np.asarray(annual_pot_cell_runoff, _f)
        mean = np.zeros(a.shape[1], _f)
        for y in range(a.shape[0]):
            mean = (mean + a[y]).astype(_f)
        mean = (mean / _f(a.shape[0])).astype(_f)
        inb = np.asarray(sbasin) == self.station_number
        tot = np.cumsum(np.abs(mean[inb]), dtype=_f)[-1] if inb.any() else _f(0)  # sequential single-precision sum, cell order
        out = np.array(cell_corr_fact, np.float64, copy=True)
        d = _f(sim - meas)
        v = (_f(1) - (np.sign(mean[inb]).astype(_f) * d).astype(_f) / tot).astype(_f)
        out[inb] = np.clip(v.astype(np.float64), 0.5, 1.5)
        return out

    def write_corr_factors(self, gamma, cell_corr_fact_ind=None):
        """writeCorrFactors (:660-728): the station correction factor CFS -> (cfs, STAT_CORR_FACTOR.OUT data line)"""
        ind = self.cell_corr_factor_ind if cell_corr_fact_ind is None else cell_corr_fact_ind
        gamma = _f(gamma)
        msum, rsum, n = _f(0), _f(0), 0
        for i in range(self.measured.size):
            if self.measured[i] > -1:
                n += 1
                msum = _f(msum + self.measured[i])
                rsum = _f(rsum + self.sim_runoff[i])
        if gamma == self.GAMMA_UPPER and ind == 99:
            msum = _f(float(msum) * 1.1)
        if gamma == self.GAMMA_LOWER and ind == 99:
            msum = _f(float(msum) * 0.9)
        cfs = _f(msum / rsum) if ind == 99 else _f(1.0)
        if (1.0 < float(cfs) < 1.01) or (0.99 < float(cfs) < 1.0):
            cfs = _f(1.0)
        if cfs > 1.0 or cfs < 1.0:
            self.calib_status = 4
        line = "\t".join([_g(_f(msum / _f(n))), _g(_f(rsum / _f(n))), _g(self.gamma_of_previous_run), _g(self.cell_corr_factor), _g(cfs)])
        return float(cfs), line


def upstream_basin(downstream_cell, station_cell):
    """bool [ncell]: the station cell and every cell draining into it (what G_sbasin == station marks, calib_basins.cpp);
    downstream_cell holds 1-based cell numbers (G_OUTFLC), station_cell is 0-based"""
    dc = np.asarray(downstream_cell, np.int64) - 1
    inb = np.zeros(dc.size, bool)
    inb[station_cell] = True
    while True:
        new = (dc >= 0) & inb[np.clip(dc, 0, None)] & ~inb
        if not new.any():
            return inb
        inb |= new


def calibrate_gamma(run_years, cal, gamma0, max_runs=60):
    """The calibration loop of integrate_wghm_ (integrateWGHM.cpp:289, 967-972, 1091-1116) around GammaCalibration `cal`:
    run_years(gamma) simulates the evaluation years with gamma in the station's basin and returns, per evaluation year,
    (annual station discharge, satisfied water use, upstream-station inflow) in km3/year.  The search ends when
    find_new_gamma returns -99; one more run with the last gamma (the reference's test run) fixes CFS.
    -> dict(gamma, cfa, cfs, calib_status, runs, stat_corr_factor_line)"""
    gamma, test_run, runs = np.float32(gamma0), False, 0
    while runs < max_runs:
        res = run_years(float(gamma))
        runs += 1
        for i, (q, use, inflow) in enumerate(res):
            cal.set_runoff(cal.y0 + i, q)
            cal.set_water_use(cal.y0 + i, use)
            cal.set_upst_inflow(cal.y0 + i, inflow)
        if test_run:
            cfs, line = cal.write_corr_factors(gamma)
            return {"gamma": float(gamma), "cfa": float(cal.cell_corr_factor), "cfs": cfs, "calib_status": cal.calib_status, "runs": runs,
                    "stat_corr_factor_line": line}
        gamma_old = gamma
        gamma = cal.find_new_gamma(gamma)
        if gamma < 0:
            test_run, gamma = True, gamma_old
    raise RuntimeError("calibrate_gamma: no termination within max_runs")
