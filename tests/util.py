"""Shared helpers of the test-suite: tolerance policy and golden access."""
import numpy as np

# Parity tolerance (BASELINE.json north_star: <= 1e-10 relative on daily storages and
# discharge).  The oracle is bit-identical to the compiled reference; the GPU path uses CUDA's
# libm (exp/pow within 2 ulp) instead of glibc's, so its results differ in the last bits.
# The error of a value is measured relative to max(|ref|, |got|, FLOOR), where FLOOR only keeps
# physically empty remainders (a river emptied to 1e-13 km3 = 100 litres, a corrected AET of
# 1e-15 mm left over from terms of order 1 mm) from being judged by their last bits:
#   1e-9 km3 (one cubic metre) for storage volumes and discharge,
#   1e-6 mm (one nanometre of water) for depths and daily fluxes,
#   1e-6 for reduction factors and area fractions (percent).
# Everything above those floors is held to 1e-10 RELATIVE.  (Round 1 used floors of 1e-3 km3 /
# 1 mm / 1, under which most discharges of the test worlds were effectively compared absolutely.)
RTOL = 1e-10
KM3 = {"gw", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor",
       "discharge", "cell_runoff", "river_evapo"}
SCALE_KM3, SCALE_MM, SCALE_DIMLESS = 1e-9, 1e-6, 1e-6
# land_aet / land_aet_uncorr are CLOSURE RESIDUALS of the cell's daily balance (daily.cpp:1135-1141, 1222-1239: when the
# soil runs dry "dailyAET += soil" leaves the difference of terms of 1-100 mm): in ~1 % of the cell-days the value is an
# ulp-sized remainder (1e-16 .. 2e-15 mm, also between two CPU builds of the same source), so they are judged against
# the scale of their operands, a micrometre
SCALE_RESIDUAL_MM = 1e-3


# How many values may exceed RTOL (they are always LISTED, never hidden):
#  one step from identical state: discharge = inflow + prevR - Sr - E and cell_runoff = discharge - upstream inflow are
#    differences of storages, so the <= 2 ulp between CUDA's and glibc's exp()/pow() in Sr reappear magnified by Sr / discharge
#    in slow reaches; measured on B200 (tools/parity_report.py, profiles/r2_parity.md): 0.9 - 1.5 values per million, worst 1.4e-8
#  free runs of 20 - 60 days additionally let those last bits grow in the cells of DESIGN.md 6: measured 2 - 8 per million on
#    the 3000 / 67420-cell worlds (worst 3.4e-9), 275 per million on the 1000-cell golden world (one cell, 1.03e-7)
ONE_STEP_PPM, ONE_STEP_MAX_REL = 5.0, 1e-7
FREE_RUN_PPM, FREE_RUN_MAX_REL = 50.0, 1e-6


def floor_of(name):
    if name in KM3:
        return SCALE_KM3
    if name in ("land_aet", "land_aet_uncorr"):
        return SCALE_RESIDUAL_MM
    if name.startswith("red_") or "frac" in name or name.startswith("fswb") or name == "k_release":
        return SCALE_DIMLESS
    return SCALE_MM


def rel_err(name, a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor_of(name))


def pure_rel_err(a, b):
    """|a - b| / max(|a|, |b|) with no floor at all (0 where both are 0)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.abs(a), np.abs(b))
    return np.where(den > 0, np.abs(a - b) / np.where(den > 0, den, 1.0), 0.0)


def assert_parity(name, ref, got, rtol=RTOL, max_flips=0, outlier_rtol=1e-6, flips=None, tag=None):
    """Integer fields bit-exact; floating point within rtol except for at most `max_flips`
    values (free runs only: cells whose dynamics amplify ulp noise), which must
    still be within `outlier_rtol`.  Every value beyond rtol is appended to `flips` (if given) as
    (tag, field, index, ref, got, rel) so that callers can list them instead of hiding them.
    Returns the number of values beyond rtol."""
    ref = np.asarray(ref)
    got = np.asarray(got)
    assert ref.shape == got.shape, (name, ref.shape, got.shape)
    if ref.dtype.kind != "f":
        bad = np.nonzero(ref != got)[0]
        assert bad.size == 0, f"{name}: {bad.size} integer mismatches, first at {bad[:5]}"
        return 0
    e = rel_err(name, ref, got)
    bad = np.nonzero(~(e <= rtol))[0]
    if flips is not None:
        for k in bad:
            flips.append((tag, name, int(k), float(ref[k]), float(got[k]), float(e[k])))
    if bad.size > max_flips or (bad.size and not (e[bad] <= outlier_rtol).all()):
        k = bad[np.argmax(e[bad])]
        raise AssertionError(f"{name}: {bad.size} values beyond rtol={rtol:g} (allowed {max_flips}); worst index {k}: "
                             f"ref {ref[k]!r} got {got[k]!r} rel {e[k]:.3e}")
    return int(bad.size)


class ParityReport:
    """worst error with the floors and without any floor, and the list of values beyond RTOL, over a
    sequence of field comparisons (smoke(), the bench parity samples and the parity tests print it)"""

    def __init__(self):
        self.worst, self.worst_pure, self.worst_at, self.flips, self.nvalues = 0.0, 0.0, None, [], 0

    def add(self, name, ref, got, tag=None):
        ref = np.asarray(ref)
        got = np.asarray(got)
        if ref.dtype.kind != "f":
            assert np.array_equal(ref, got), f"{name}: integer mismatch"
            return
        e = rel_err(name, ref, got)
        pe = pure_rel_err(ref, got)
        self.nvalues += int(e.size)
        if e.size:
            k = int(np.argmax(e))
            if float(e[k]) > self.worst:
                self.worst, self.worst_at = float(e[k]), (tag, name, k, float(ref[k]), float(got[k]))
            self.worst_pure = max(self.worst_pure, float(pe.max()))
        for k in np.nonzero(~(e <= RTOL))[0]:
            self.flips.append((tag, name, int(k), float(ref[k]), float(got[k]), float(e[k])))

    def check(self, max_ppm, max_rel, min_allowed=0, what=""):
        """the policy of DESIGN.md 6: at most max(min_allowed, max_ppm per million compared values) values beyond
        1e-10, each within max_rel; a failure lists the offending values (worst first)"""
        allowed = max(min_allowed, int(max_ppm * self.nvalues / 1e6))
        worst = max((f[5] for f in self.flips), default=0.0)
        if len(self.flips) > allowed or not (worst <= max_rel):
            top = sorted(self.flips, key=lambda f: -f[5])[:12]
            raise AssertionError(f"{what}: {len(self.flips)} of {self.nvalues} values beyond {RTOL:g} (allowed {allowed}), worst {worst:.3e} "
                                 f"(allowed {max_rel:g}); (tag, field, index, ref, got, rel): {top}")
        return len(self.flips)

    def summary(self):
        return {"values": self.nvalues, "worst_rel": self.worst, "worst_rel_no_floor": self.worst_pure,
                "beyond_1e-10": len(self.flips), "worst_at": self.worst_at}


def golden_day(golden, day):
    pre = f"d{day}/"
    return {k[len(pre):]: v for k, v in golden.items() if k.startswith(pre)}
