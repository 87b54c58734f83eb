"""Pins the plain-C oracle (oracle/stable_fluids_oracle.c, SFO_SEM_CPU) bit-for-bit against
(a) the UNMODIFIED reference CPU solver compiled out of /root/reference (oracle/_ref/libref_cpu.so)
and (b) the anchors SURVEY.md Appendix D / BASELINE.md section 5 recorded from that solver.
The reference ships no tests or golden vectors of its own."""
import json
import os

import numpy as np
import pytest

from util import DIFFUSION_RATE, DT, VISCOSITY, assert_bitwise, rng_fields

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def refcpu():
    from oracle import refs

    if not refs.have_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return refs.ref_cpu()


def test_canonical_fields_match_anchor_step0(sfo):
    anchors = json.load(open(os.path.join(GOLDEN, "anchors_ref_cpu_256.json")))
    d, u, v, sd, su, sv = sfo.canonical_fields(256)
    for name, a in (("d", d), ("u", u), ("v", v)):
        assert "%016x" % sfo.fnv1a64(a) == anchors["0"][name]["fnv"], name
    assert sd.sum() > 0 and su.sum() == 0 and sv.sum() > 0


@pytest.mark.parametrize("steps", [1, 10, 100])
def test_oracle_cpu_semantics_reproduces_anchors(sfo, steps):
    """fluid_solver_cpu::solve anchors (N=256, K=20, no smooth) from the oracle alone: this check
    also runs on the GPU box where /root/reference does not exist."""
    anchors = json.load(open(os.path.join(GOLDEN, "anchors_ref_cpu_256.json")))
    d, u, v, sd, su, sv = sfo.canonical_fields(256)
    d, u, v = sfo.steps(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 20, 20, smooth=False,
                        sem=sfo.SEM_CPU, nsteps=steps)
    for name, a in (("d", d), ("u", u), ("v", v)):
        st = sfo.field_stats(a)
        want = anchors[str(steps)][name]
        assert st["fnv"] == want["fnv"], (name, st, want)
        assert abs(st["sum"] - want["sum"]) <= 1e-9 * max(1.0, abs(want["sum"]))


def test_ref_cpu_reproduces_anchors(sfo, refcpu):
    anchors = json.load(open(os.path.join(GOLDEN, "anchors_ref_cpu_256.json")))
    d, u, v, sd, su, sv = sfo.canonical_fields(256)
    d, u, v = refcpu.solve(d, sd, DIFFUSION_RATE, u, v, su, sv, VISCOSITY, DT, 10)
    for name, a in (("d", d), ("u", u), ("v", v)):
        assert "%016x" % sfo.fnv1a64(a) == anchors["10"][name]["fnv"], name


@pytest.mark.parametrize("n,seed", [(16, 1), (33, 2), (64, 3), (100, 4)])
def test_stages_bitwise_vs_ref_cpu(sfo, refcpu, n, seed):
    d, u, v, sd, su, sv = rng_fields(n, seed, vel_cells=4.0)
    for kind in (sfo.BND_CONTINUOUS, sfo.BND_OPPOSITE_HORIZONTAL, sfo.BND_OPPOSITE_VERTICAL):
        assert_bitwise(sfo.set_bnd(u, kind, sfo.SEM_CPU), refcpu.set_bnd(u, kind), "set_bnd kind %d" % kind)
        for rate in (DIFFUSION_RATE, VISCOSITY, 1e-4):
            assert_bitwise(sfo.diffuse(d, kind, rate, DT, 7, sfo.SEM_CPU), refcpu.diffuse(d, kind, rate, DT, 7),
                           "diffuse kind %d rate %g" % (kind, rate))
        assert_bitwise(sfo.advect_gather(d, u, v, kind, DT, sfo.SEM_CPU), refcpu.advect(d, u, v, kind, DT, False),
                       "advect gather")
        assert_bitwise(sfo.advect_scatter(d, u, v, kind, DT, sfo.SEM_CPU), refcpu.advect(d, u, v, kind, DT, True),
                       "advect scatter")
    assert_bitwise(sfo.add_sources(d, sd, DT, sfo.SEM_CPU), refcpu.add_sources(d, sd, DT), "add_sources")
    ou, ov = sfo.project(u, v, 9, sfo.SEM_CPU)
    ru, rv = refcpu.project(u, v, 9)
    assert_bitwise(ou, ru, "project u")
    assert_bitwise(ov, rv, "project v")


@pytest.mark.parametrize("n,kd,kp,steps", [(48, 20, 20, 5), (64, 7, 11, 3), (130, 40, 40, 2)])
def test_steps_bitwise_vs_ref_cpu(sfo, refcpu, n, kd, kp, steps):
    d, u, v, sd, su, sv = rng_fields(n, 100 + n)
    od, ou, ov = sfo.steps(d, sd, 1e-4, u, v, su, sv, 1e-4, DT, kd, kp, smooth=False, sem=sfo.SEM_CPU, nsteps=steps)
    rd, ru, rv = refcpu.step_k(d, sd, 1e-4, u, v, su, sv, 1e-4, DT, kd, kp, nsteps=steps)
    assert_bitwise(od, rd, "d")
    assert_bitwise(ou, ru, "u")
    assert_bitwise(ov, rv, "v")
