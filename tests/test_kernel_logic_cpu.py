"""CPU: the product's CUDA kernel SOURCE (csrc/wgk_kernels.cuh), compiled for the host with a
shim (tests/emu) and executed thread by thread with glibc's libm, against the oracle.

This separates "is the kernel logic right" (checked here without a GPU) from "how far does
CUDA's libm move the result" (checked on the GPU).  The kernels deliberately deviate from the
reference's expression shapes in four places to shorten instruction chains:
pow(T,4) as two squarings, pow(x,2/3) as cbrt(x*x), the per-band S*a/b as a Markstein
reciprocal-correction quotient, and 1/(M*roughness) precomputed; each is within ~1 ulp of the
original, so a 31-day free run must stay within 2e-11 of the oracle, integer state exactly equal.
(Before those four substitutions the same harness was bit-identical to the oracle over 120 days
on the 67 420-cell world.)"""
import numpy as np
import pytest

from tests.util import rel_err


@pytest.mark.parametrize("tail_level0", [-1, 3])
def test_kernel_source_on_host_matches_oracle(world3000, oracle_lib, tail_level0):
    from oracle import synth_world as sw, wg_init
    from tests.emu import Emu
    wgo = oracle_lib
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    o = wgo.Oracle(w.ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    e = Emu(w.ng, topo["rout_order"], topo["outflow_cell"])
    e.load(ini, o)
    names = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS
    worst = 0.0
    for sd in range(1, 32):
        doy, mon, dom = wgo.calendar(sd)
        if dom == 1:
            f = sw.forcing_month(w, 1901, mon + 1)
            o.set_forcing_month(f)
            e.set_forcing(f)
        o.step_day(doy, mon, dom)
        e.day(doy, mon, dom, dom - 1, tail_level0)
        if sd in (1, 2, 15, 31):
            for nm in names:
                ref = o.field(nm)
                got = e.get(nm, ref)
                if ref.dtype.kind != "f":
                    assert np.array_equal(ref, got), f"day {sd} {nm}"
                else:
                    d = float(rel_err(nm, ref, got).max())
                    worst = max(worst, d)
                    assert d < 2e-11, f"day {sd} {nm}: {d:.2e}"
    assert worst > 0  # the four substitutions are really in effect
