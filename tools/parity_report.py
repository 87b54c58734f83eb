"""Per-field, per-step parity report: the CUDA path against the UNMODIFIED reference solvers run live
on the same box (oracle/_ref/libref_gpu.so, libref_cpu.so).  Writes markdown to the path given (default
gpurun_out/parity_report.md).  TEST TOOLING: uses oracle/ only as the checker."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np  # noqa: E402

import fluid2d_b200 as f2d  # noqa: E402
from oracle import refs, sfo  # noqa: E402
from util import err  # noqa: E402

DT, RATE, VISC = 0.02, 0.5, 1e-6


def fmt(e):
    return "%.2e / %.2e / %d" % (e["max_abs"], e["rel_l2"], e["n_diff"])


def gpu_semantics(out, n, kd, kp, steps, literal):
    """free-running: both solvers advance their own state from the same start"""
    g = refs.ref_gpu()
    f = sfo.canonical_fields(n)
    sd, su, sv = f[3], f[4], f[5]
    ref = [a.copy() for a in f[:3]]
    ref2 = [a.copy() for a in f[:3]]  # a second run of the reference: its own run-to-run noise (atomics)
    dflt = [a.copy() for a in f[:3]]
    exact = [a.copy() for a in f[:3]]
    out.append("\n### %dx%d, Kd=%d, Kp=%d, smooth on, canonical fields, %d free-running steps (%s)\n" %
               (n, n, kd, kp, steps, "fluid_solver_gpu::solve literally" if literal else "reference stage methods with these K"))
    out.append("max-abs / rel-L2 / cells whose bits differ, against the live `fluid_solver_gpu`\n")
    out.append("| step | field | this repo, default (fp32-corrected divide) | this repo, `F2D_DIV_F64` | reference vs itself (2nd run) |")
    out.append("|---|---|---|---|---|")
    if n > 8192:
        return gpu_semantics_big(out, g, f, n, kd, kp)
    with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp) as sa, \
            f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, divide_mode=f2d.DIV_F64) as sb:
        for s in range(1, steps + 1):
            if literal:
                ref = list(g.solve(ref[0], sd, RATE, ref[1], ref[2], su, sv, VISC, DT, 1)[:3])
                ref2 = list(g.solve(ref2[0], sd, RATE, ref2[1], ref2[2], su, sv, VISC, DT, 1)[:3])
            else:
                ref = list(g.step_k(ref[0], sd, RATE, ref[1], ref[2], su, sv, VISC, DT, kd, kp, True, 1)[:3])
                ref2 = list(g.step_k(ref2[0], sd, RATE, ref2[1], ref2[2], su, sv, VISC, DT, kd, kp, True, 1)[:3])
            sa.solve(dflt[0], sd, RATE, dflt[1], dflt[2], su, sv, VISC, DT)
            sb.solve(exact[0], sd, RATE, exact[1], exact[2], su, sv, VISC, DT)
            if s in (1, 2, steps // 2, steps):
                for k, name in enumerate("duv"):
                    out.append("| %d | %s | %s | %s | %s |" % (s, name, fmt(err(dflt[k], ref[k])), fmt(err(exact[k], ref[k])),
                                                            fmt(err(ref2[k], ref[k]))))


def gpu_semantics_big(out, g, f, n, kd, kp):
    """one step, one solver resident at a time (a 16384^2 solver holds 19 GiB, the reference 7 GiB)"""
    sd, su, sv = f[3], f[4], f[5]
    ref = list(g.step_k(f[0].copy(), sd, RATE, f[1].copy(), f[2].copy(), su, sv, VISC, DT, kd, kp, True, 1)[:3])
    ref2 = list(g.step_k(f[0].copy(), sd, RATE, f[1].copy(), f[2].copy(), su, sv, VISC, DT, kd, kp, True, 1)[:3])
    res = {}
    for tag, kw in (("dflt", {}), ("exact", {"divide_mode": f2d.DIV_F64})):
        h = [a.copy() for a in f[:3]]
        with f2d.FluidSolverB200(n, n, diffuse_iters=kd, project_iters=kp, **kw) as s:
            s.solve(h[0], sd, RATE, h[1], h[2], su, sv, VISC, DT)
        res[tag] = h
    for k, name in enumerate("duv"):
        out.append("| 1 | %s | %s | %s | %s |" % (name, fmt(err(res["dflt"][k], ref[k])), fmt(err(res["exact"][k], ref[k])),
                                                fmt(err(ref2[k], ref[k]))))


def cpu_semantics(out, n, steps):
    c = refs.ref_cpu()
    f = sfo.canonical_fields(n)
    sd, su, sv = f[3], f[4], f[5]
    ref = [a.copy() for a in f[:3]]
    mine = [a.copy() for a in f[:3]]
    out.append("\n### %dx%d, `F2D_SEM_CPU` against the live `fluid_solver_cpu::solve` (K=20, no smooth), %d free-running steps\n" % (n, n, steps))
    out.append("| step | field | max-abs / rel-L2 / cells whose bits differ |")
    out.append("|---|---|---|")
    with f2d.FluidSolverB200.cpu_compatible(n, n) as s:
        for st in range(1, steps + 1):
            ref = list(c.solve(ref[0], sd, RATE, ref[1], ref[2], su, sv, VISC, DT, 1))
            s.solve(mine[0], sd, RATE, mine[1], mine[2], su, sv, VISC, DT)
            if st in (1, steps // 2, steps):
                for k, name in enumerate("duv"):
                    out.append("| %d | %s | %s |" % (st, name, fmt(err(mine[k], ref[k]))))


def main(path, big=False):
    out = ["# Parity report, round 2 (B200; produced by `tools/parity_report.py`)", "",
           "Both reference solvers are the UNMODIFIED translation units of `/root/reference/src`, compiled by `oracle/Makefile`",
           "and run live on the same box through `oracle/refs.py`.  Stated tolerances: DESIGN.md section 3."]
    if refs.have_gpu():
        gpu_semantics(out, 256, 15, 20, 10, True)
        gpu_semantics(out, 1024, 40, 40, 4, False)
        if big:
            gpu_semantics(out, 4096, 80, 80, 2, False)
            gpu_semantics(out, 16384, 80, 80, 1, False)
    else:
        out.append("\n(libref_gpu.so did not travel to this box)")
    if refs.have_cpu():
        cpu_semantics(out, 256, 10)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out[-12:]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_report.md", big="--big" in sys.argv)
