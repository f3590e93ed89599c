#!/usr/bin/env python3
"""Per-phase breakdown of a step_kernel capture: ncu source page (cuda,sass correlation; needs -lineinfo + --import-source on)
-> static SASS instructions, executed warp instructions and stall samples per phase of the kernel source.
Usage: python tools/ncu_phase_breakdown.py gpurun_out/r02a_step_65536.ncu-rep > profiles/r02a_phase_breakdown_65536.txt"""
import csv
import io
import re
import subprocess
import sys

# phase = first marker (substring of a source line of rexsim_kernel.cu) at or before the line; order matters
MARKERS = [
    ("tables/TMA + state load", "__device__ __forceinline__ void tma_load_tables"),
    ("6x6 inverse, quat/euler helpers", "__device__ __forceinline__ M3 quat_to_mat"),
    ("ground query / tile", "// ground query ---"),
    ("motor model", "// motor model + overheat"),
    ("link damping helper", "__device__ __forceinline__ void link_damping"),
    ("forward kinematics", "// ---- forward kinematics"),
    ("velocities / bias / inertias", "// ---- velocities and bias terms"),
    ("ABA inward", "// ---- ABA inward pass"),
    ("base reduce + inverse", "// ---- base: reduce the 4 legs"),
    ("ABA outward + clamp", "// ---- ABA outward pass"),
    ("contact candidates", "// ---- contact candidates"),
    ("limits / flags", "// contact while the distance is below"),
    ("fast path: rows + Delassus", "// ================= fast path"),
    ("fast path: rhs", "// right-hand sides (btMultiBodyConstraintSolver"),
    ("fast path: PGS", "// ---- PGS in impulse space"),
    ("fast path: apply impulse", "// ---- apply the net contact impulse"),
    ("generic path", "// ================= generic path"),
    ("integrate", "// ---- integrate (btMultiBody::stepPositionsMultiDof)"),
    ("apply_action (motor, overheat)", "// Rex.ApplyAction + stepSimulation"),
    ("gait planner + IK", "// gait planner + IK for the own leg"),
    ("task command", "// <task>._transform_action_to_motor_command"),
    ("state load/store, reset, obs", "// state load / store (SoA"),
    ("step kernel body (actions, reward, done, io)", "// the fused step kernel"),
    ("other kernels", "// reset kernel: BatchEnv.reset"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    files, cur, hdr = {}, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]; files.setdefault(cur, []); continue
        if r[0] == "Function Name":
            kernel = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if cur is not None and hdr is not None:
            files[cur].append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    col = lambda r, name: float(r[ix[name]]) if r[ix[name]] not in ("-", "") else 0.0
    src = open([f for f in files if f.endswith("rexsim_kernel.cu")][0].replace("/root/repo", ".")).read().splitlines() if False else None
    import os
    kpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rex_gym_b200", "csrc", "rexsim_kernel.cu")
    lines = open(kpath).read().splitlines()
    marks = []
    for name, key in MARKERS:
        for n, l in enumerate(lines, 1):
            if key in l:
                marks.append((n, name)); break
    marks.sort()

    def phase_of(fn, line):
        if not fn.endswith("rexsim_kernel.cu"):
            return "helpers: " + os.path.basename(fn)
        p = "top of file"
        for n, name in marks:
            if n <= line:
                p = name
        return p
    agg = {}
    tot = [0.0, 0.0, 0.0, 0.0]
    for fn, data in files.items():
        line = None
        for r in data:
            if r[0] != "":
                line = int(r[0]); continue          # the per-line aggregate row; the SASS rows below carry the same numbers
            if r[ix["Address"]] in ("...", "-", ""):
                continue
            a = agg.setdefault(phase_of(fn, line or 0), [0.0, 0.0, 0.0, 0.0])
            v = [1.0, col(r, "Instructions Executed"), col(r, "# Samples"), col(r, "stall_no_inst")]
            for k in range(4):
                a[k] += v[k]; tot[k] += v[k]
    print(f"{kernel}\n{rep}: static SASS {int(tot[0])} instructions ({int(tot[0]) * 16 // 1024} KB), executed {int(tot[1])} warp instructions, {int(tot[2])} stall samples "
          f"({100 * tot[3] / max(tot[2], 1):.1f} % of them no_instruction)\n")
    print(f"{'phase':48s} {'static':>7s} {'KB':>5s} {'instr %':>8s} {'samples %':>10s} {'no_inst % of phase':>19s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:48s} {int(a[0]):7d} {a[0] * 16 / 1024:5.1f} {100 * a[1] / tot[1]:8.2f} {100 * a[2] / max(tot[2], 1):10.2f} {100 * a[3] / max(a[2], 1):19.1f}")


if __name__ == "__main__":
    main()
