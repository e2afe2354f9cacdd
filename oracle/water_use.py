"""TEST INFRASTRUCTURE - first piece of the restatement for SURVEY 8f row 4 (water use; NOT on the product path yet).

  daily_net_abstraction()   routingClass::dailyNUInit (routing.cpp:884-977, subtract_use 2, aggrNUsGloLakResOpt 0) followed by
                            calcNextDay_M (:7414-7440): the monthly net abstractions from surface water / groundwater in m3 per
                            month, times the cell's multiplier (M_NETABSSW / M_NETABSGW), as km3 per day of the given month

  update_net_abstraction_gw()  routingClass::updateNetAbstractionGW (:5503-5572)

Pinned against the compiled reference: tests/golden/ref_ng1000_wateruse.npz (a 59-day run with net abstractions, and the unit
vectors of `ref_harness wu_unit`; tests/test_oracle_golden.py).  The use satisfaction inside routing() (:2193-2296, 2922-2946,
3590-3920) is not restated yet.
Only tests/ may import this module.
"""
import numpy as np

NDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def daily_net_abstraction(monthly_m3, multiplier, month):
    """monthly_m3: float32 [ng][12] as read from G_NETUSE_*_m3_<year>.12.UNF0 into a Grid<double>; multiplier [ng] (double);
    month 0..11 -> km3/day [ng]: (multiplier * value) / (1000000000. * (double) days)"""
    v = np.asarray(monthly_m3, np.float32).reshape(-1, 12)[:, month].astype(np.float64)
    return (np.asarray(multiplier, np.float64) * v) / (1000000000. * float(NDAYS[month]))


def update_net_abstraction_gw(rem, wusi, cusi, frgi, uns_irr, uns_oth, red_rf, nug):
    """routingClass::updateNetAbstractionGW (routing.cpp:5503-5572) for every cell at once: the net abstraction from
    groundwater adapted to the surface-water use that stayed unsatisfied (rem > 0: return flows reduced) or was satisfied late
    (rem < 0: return flows reintroduced).  Inputs [ng] doubles: dailyRemainingUse, withdrawalIrrigFromSwb,
    consumptiveUseIrrigFromSwb, fractreturngw_irrig, unsatisfiedNAsFromIrrig, unsatisfiedNAsFromOtherSectors, reducedReturnFlow,
    dailydailyNUg.  -> (NAg returned, rem, uns_irr, uns_oth, red_rf) after the call; nug itself is not modified by the function"""
    rem, wusi, cusi, frgi = (np.array(x, np.float64) for x in (rem, wusi, cusi, frgi))
    uns_irr, uns_oth, red_rf, nug = (np.array(x, np.float64) for x in (uns_irr, uns_oth, red_rf, nug))
    out = nug.copy()
    for n in range(rem.size):
        if rem[n] > 1.e-12:
            if wusi[n] > 0.:
                eff = cusi[n] / wusi[n]
                factor = 1 - (1 - frgi[n]) * (1 - eff)
                wusi_new = 1 / factor * (wusi[n] * factor - rem[n])
                if wusi_new < 0.:
                    wusi_new = 0.
                    uns_irr[n] += wusi[n] * factor
                    uns_oth[n] += rem[n] - (wusi[n] * factor)
                else:
                    uns_irr[n] += rem[n]
                change = frgi[n] * (1 - eff) * (wusi_new - wusi[n])
                red_rf[n] += change
                out[n] = nug[n] - change
                rem[n] = 0.
            else:
                uns_oth[n] += rem[n]
        elif rem[n] < -1.e-12:
            from_irr = rem[n] + uns_oth[n]
            if from_irr < 0.:
                uns_oth[n] = 0.
            else:
                uns_oth[n] += rem[n]
                rem[n] = 0.
                continue
            if uns_irr[n] == 0.:
                rem[n] = 0.
                continue
            ratio = from_irr / uns_irr[n]
            if ratio < -1.:
                ratio = -1.
                from_irr = uns_irr[n] * -1.
            change = ratio * red_rf[n]
            uns_irr[n] += from_irr
            red_rf[n] += change
            out[n] = nug[n] - change
            rem[n] = 0.
    return out, rem, uns_irr, uns_oth, red_rf


def month_inputs(files, params, month):
    """the water-use inputs of one month for the oracle / the kernels, km3 per day: dict of wu_nus_month, wu_nug_month (net
    abstractions times M_NETABSSW / M_NETABSGW, dailyNUInit), wu_wusi_month, wu_cusi_month (irrigation withdrawal / consumptive
    use from surface water, routing.cpp:3907-3908: no multiplier).  `files` maps the input file names (G_NETUSE_SW_m3_..,
    G_NETUSE_GW_m3_.., G_IRRIG_WITHDRAWAL_USE_SW_m3_.., G_IRRIG_CONS_USE_SW_m3_..) to their float32 [ng][12] contents."""
    par = np.asarray(params, np.float64).reshape(26, -1)
    one = np.ones(par.shape[1])
    key = lambda stem: next(v for k, v in files.items() if stem in k)
    return {"wu_nus_month": daily_net_abstraction(key("G_NETUSE_SW_m3"), par[23], month),
            "wu_nug_month": daily_net_abstraction(key("G_NETUSE_GW_m3"), par[24], month),
            "wu_wusi_month": daily_net_abstraction(key("G_IRRIG_WITHDRAWAL_USE_SW_m3"), one, month),
            "wu_cusi_month": daily_net_abstraction(key("G_IRRIG_CONS_USE_SW_m3"), one, month)}
