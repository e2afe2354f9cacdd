"""CPU: the oracle (and the reference-loop replay of the harness) against the COMPILED
REFERENCE run live on a 3000-cell world.  Uses oracle/_ref/ref_harness_3000 (built by
__graft_entry__.build() where /root/reference exists; the binary travels to the GPU box)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness_3000")
NG, NDAYS = 3000, 45


@pytest.fixture(scope="module")
def ref_run(tmp_path_factory, world3000):
    if not os.path.exists(HARNESS):
        if os.path.isdir(os.environ.get("WG_REF_SRC", "/root/reference/source")):
            subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
        else:
            pytest.skip("compiled reference (oracle/_ref/ref_harness_3000) not available")
    from oracle import synth_world as sw
    tmp = str(tmp_path_factory.mktemp("refworld"))
    sw.write_world(world3000, tmp, (1901, 1901), (1, 2))
    dump = os.path.join(tmp, "dump.wgd")
    log = open(os.path.join(tmp, "replay.log"), "w")
    subprocess.check_call([HARNESS, "replay", os.path.join(tmp, "config.txt"), dump, "--days", f"1-{NDAYS}",
                           "--snow-days", f"{NDAYS}-{NDAYS}", "--final-state", os.path.join(tmp, "output", "replay")],
                          stdout=log, stderr=log, cwd=tmp)
    return tmp, dump


def test_replay_equals_reference_driver(ref_run):
    """the harness' replay of the day loop must leave exactly the state files that the
    reference's own entry points initialize_wghm/integrate_wghm write (16-digit txt)."""
    tmp, _ = ref_run
    log = open(os.path.join(tmp, "driver.log"), "w")
    subprocess.check_call([HARNESS, "driver", os.path.join(tmp, "config.txt")], stdout=log, stderr=log, cwd=tmp)
    out = os.path.join(tmp, "output")
    assert filecmp.cmp(os.path.join(out, "wghm_state_lastday.txt"), os.path.join(out, "replay_state.txt"), shallow=False)
    assert filecmp.cmp(os.path.join(out, "snow_lastday.txt"), os.path.join(out, "replay_snow.txt"), shallow=False)
    assert filecmp.cmp(os.path.join(out, "additional_lastday.txt"), os.path.join(out, "replay_additional.txt"), shallow=False)


def test_oracle_bit_exact_vs_live_reference(ref_run, oracle_lib, world3000):
    wgo = oracle_lib
    from oracle import synth_world as sw
    _, dump = ref_run
    check_days = {1, 2, 3, 10, 20, 31, 32, 40, NDAYS}
    recs = wgo.read_dump(dump, days=check_days | {0})
    o = wgo.Oracle(NG)
    assert o.load_records(recs, 0) > 70
    curm, n = -1, 0
    for sd in range(1, NDAYS + 1):
        doy, mon, dom = wgo.calendar(sd)
        if mon != curm:
            o.set_forcing_month(sw.forcing_month(world3000, 1901, mon + 1))
            curm = mon
        o.step_day(doy, mon, dom)
        if sd in check_days:
            for (name, d), ref in recs.items():
                if d == sd and o.has(name):
                    assert np.array_equal(ref, o.field(name)), f"day {sd} {name}"
                    n += 1
    assert n > 250


def test_init_restatement_vs_live_reference(ref_run, oracle_lib, world3000):
    from oracle import wg_init
    _, dump = ref_run
    recs = oracle_lib.read_dump(dump, days={0})
    d = wg_init.derive(world3000)
    n = 0
    for (name, _), ref in recs.items():
        if name in d:
            assert np.array_equal(ref, np.asarray(d[name]).ravel().astype(ref.dtype)), name
            n += 1
    assert n >= 75
