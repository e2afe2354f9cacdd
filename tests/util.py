"""Shared helpers of the test-suite: tolerance policy and golden access."""
import numpy as np

# Parity tolerance (BASELINE.json north_star: <= 1e-10 relative on daily storages and
# discharge).  The oracle is bit-identical to the compiled reference; the GPU path uses CUDA's
# libm (exp/pow within 2 ulp) instead of glibc's, so its results differ in the last bits.
# Where a value is the difference of nearly equal numbers (a river emptied to 1e-12 km3, a
# corrected AET of 1e-15 mm left over from terms of order 1 mm) those ulps are a large
# fraction of a physically meaningless remainder.  The error is therefore measured relative to
# max(|ref|, |got|, SCALE) with SCALE the characteristic magnitude of the field:
#   1 mm for water depths and daily fluxes, 1e-3 km3 (a million m3) for storage volumes and
#   discharge, 1 for reduction factors and area fractions (percent),
# i.e. |ref - got| <= 1e-10 * max(|ref|, |got|) + 1e-10 mm / 1e-13 km3 / 1e-10.
RTOL = 1e-10
KM3 = {"gw", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor",
       "discharge", "cell_runoff", "river_evapo"}
SCALE_KM3, SCALE_MM, SCALE_DIMLESS = 1e-3, 1.0, 1.0


def floor_of(name):
    if name in KM3:
        return SCALE_KM3
    if name.startswith("red_") or "frac" in name or name.startswith("fswb") or name == "k_release":
        return SCALE_DIMLESS
    return SCALE_MM


def rel_err(name, a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor_of(name))


def assert_parity(name, ref, got, rtol=RTOL, max_flips=0, outlier_rtol=1e-6):
    """Integer fields bit-exact; floating point within rtol except for at most `max_flips`
    cells (free runs only: cells whose dynamics amplify ulp noise), which must
    still be within `outlier_rtol`."""
    ref = np.asarray(ref)
    got = np.asarray(got)
    assert ref.shape == got.shape, (name, ref.shape, got.shape)
    if ref.dtype.kind != "f":
        bad = np.nonzero(ref != got)[0]
        assert bad.size == 0, f"{name}: {bad.size} integer mismatches, first at {bad[:5]}"
        return 0
    e = rel_err(name, ref, got)
    bad = np.nonzero(~(e <= rtol))[0]
    if bad.size > max_flips or (bad.size and not (e[bad] <= outlier_rtol).all()):
        k = bad[np.argmax(e[bad])]
        raise AssertionError(f"{name}: {bad.size} cells beyond rtol={rtol:g} (allowed {max_flips}); worst cell {k}: "
                             f"ref {ref[k]!r} got {got[k]!r} rel {e[k]:.3e}")
    return int(bad.size)


def golden_day(golden, day):
    pre = f"d{day}/"
    return {k[len(pre):]: v for k, v in golden.items() if k.startswith(pre)}
