"""TEST INFRASTRUCTURE - first piece of the restatement for SURVEY 8f row 4 (water use; NOT on the product path yet).

  daily_net_abstraction()   routingClass::dailyNUInit (routing.cpp:884-977, subtract_use 2, aggrNUsGloLakResOpt 0) followed by
                            calcNextDay_M (:7414-7440): the monthly net abstractions from surface water / groundwater in m3 per
                            month, times the cell's multiplier (M_NETABSSW / M_NETABSGW), as km3 per day of the given month

Pinned against the compiled reference: tests/golden/ref_ng1000_wateruse.npz (tests/test_oracle_golden.py).  The use
satisfaction inside routing() (:2193-2296, 2922-2946, 3590-3920) and updateNetAbstractionGW (:5503-5572) are not restated yet.
Only tests/ may import this module.
"""
import numpy as np

NDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def daily_net_abstraction(monthly_m3, multiplier, month):
    """monthly_m3: float32 [ng][12] as read from G_NETUSE_*_m3_<year>.12.UNF0 into a Grid<double>; multiplier [ng] (double);
    month 0..11 -> km3/day [ng]: (multiplier * value) / (1000000000. * (double) days)"""
    v = np.asarray(monthly_m3, np.float32).reshape(-1, 12)[:, month].astype(np.float64)
    return (np.asarray(multiplier, np.float64) * v) / (1000000000. * float(NDAYS[month]))
