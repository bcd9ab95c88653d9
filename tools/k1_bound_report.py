#!/usr/bin/env python
"""What bounds K1: combines the fp64 operand-port probe (tools/fp64_probe.cu output) with the executed-instruction
mix of an ncu capture of sgp4_grid_kernel and a SASS excerpt of the shipped library (CPU tool).

    python tools/k1_bound_report.py gpurun_out/prof_k1_X.ncu-rep gpurun_out/fp64_probe_X.jsonl profiles/rXX_what_bounds_k1.md
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, probe, out = sys.argv[1:4]
CELLS = 13478 * 1440

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
warp_cells = CELLS / 32
# ncu's SASS text drops the .reuse flags; the same kernel disassembled from the shipped library (cuobjdump) keeps them.
# When the capture was taken with this very build the two listings match instruction for instruction, and the flagged
# text is used with ncu's execution counts.
sass_all = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "astroz_b200", "libastroz_b200.so")], capture_output=True, text=True).stdout
flagged = []
on = False
for ln in sass_all.splitlines():
    if "Function :" in ln:
        on = re.search(r"sgp4_grid_kernelILi0ELi0ELb1ELi4ELi384ELi[23]ELi3ELi0E", ln) is not None
        continue
    if on and re.search(r"/\*[0-9a-f]{4}\*/", ln):
        flagged.append(ln.split("*/", 1)[1].split(";")[0].strip())
body = rows[2:]
def _ops(t):
    t = t.split()
    return (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
same_build = len(flagged) == len(body) and all(_ops(a) == _ops(r[ix["Source"]]) for a, r in zip(flagged, body))
cnt = collections.Counter()
total = 0
held = {}   # operand slot -> register the previous instruction asked the reuse cache to keep
for k, r in enumerate(body):
    ex = int(r[ix["Instructions Executed"]])
    total += ex
    toks = (flagged[k] if same_build else r[ix["Source"]]).split()
    if not toks:
        continue
    if toks[0].startswith("@"):
        toks = toks[1:]
    op = toks[0].split(".")[0]
    srcs = [o.strip() for o in " ".join(toks[1:]).rstrip(";").split(",")[1:]]
    fresh = 0
    keep = {}
    for slot, o in enumerate(srcs):
        m = re.match(r"^[-|]*(R\d+)(\.reuse)?", o)
        if not m or m.group(1) == "RZ":
            continue
        if held.get(slot) != m.group(1):   # not served by the reuse cache: a register-file read
            fresh += 1
        if m.group(2):
            keep[slot] = m.group(1)
    held = keep
    if op in ("DFMA", "DMUL", "DADD") and ex:
        cnt[(op, fresh)] += ex
fp64 = sum(cnt.values()) / warp_cells
three = cnt[("DFMA", 3)] / warp_cells

rawm = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
mrows = list(csv.reader(io.StringIO(rawm)))
mh = mrows[0]
mv = mrows[-1]
met = {h: v for h, v in zip(mh, mv)}
dur_us = float(met["gpu__time_duration.sum"])
pipe = float(met["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"])
sms = 148
clk = 1.965e9
cycles = 2 * fp64 + three
floor_ms = cycles * warp_cells / (sms * 4) / clk * 1e3

pr = [json.loads(l) for l in open(probe) if l.startswith("{")]
lines = []
lines.append("# What bounds K1 (`sgp4_grid_kernel`, headline grid) — probe + ncu + SASS\n")
lines.append(f"Inputs: `{os.path.basename(rep)}` (ncu `--set full --clock-control none --import-source on`), `{os.path.basename(probe)}` "
             "(`tools/fp64_probe.cu` run on the same pool's B200), SASS of the shipped `libastroz_b200.so` (`cuobjdump -sass`).\n")
lines.append("## 1. The fp64 pipe's cost model (probe)\n")
lines.append("| operand pattern (SASS shape) | cycles per warp-instruction per scheduler | dependent-issue latency |\n|---|---|---|")
for p in pr:
    if "cycles_per_warp_instr_full" in p:
        lines.append(f"| {p['pattern']} | {p['cycles_per_warp_instr_full']:.3f} | {p['dependent_issue_latency_cycles']:.2f} |")
    elif "cycles_per_dfma" in p:
        c = p["cycles_per_dfma"]
        lines.append(f"| DFMA alone / + one independent IMAD per DFMA | {c[0]:.3f} / {c[1]:.3f} | |")
lines.append("\nA DFMA/DMUL/DADD holds the pipe 2 cycles per warp (16 lanes per scheduler); the register file delivers one fresh 64-bit pair "
             "per cycle to it, so an instruction reading **three** fresh pairs holds it 3 cycles. Immediates, uniform registers and "
             "`.reuse` hits cost nothing; the costs add (the 3:2:3 mix row); an independent integer instruction hides in a DFMA's shadow.\n")
lines.append("## 2. K1's executed fp64 mix (ncu source page, per warp-cell = 32 cells)\n")
lines.append("| opcode | register pairs read from the register file (operands served by the reuse cache excluded) | executed per warp-cell |\n|---|---|---|")
for k in sorted(cnt):
    lines.append(f"| {k[0]} | {k[1]} | {cnt[k] / warp_cells:.1f} |")
lines.append("\n(" + ("reuse-cache hits resolved from the shipped library's SASS, which matches the capture instruction for instruction"
                     if same_build else "the capture predates the current build: .reuse flags unavailable, every register operand counted as a read") + ")")
lines.append(f"\nTotal {fp64:.1f} fp64 instructions of {total / warp_cells:.1f} executed per warp-cell; {three:.1f} are three-pair DFMAs.\n")
lines.append(f"**Floor for this instruction mix** = (2 × {fp64:.1f} + {three:.1f}) = {cycles:.0f} pipe cycles per warp-cell × {warp_cells:.0f} warp-cells ÷ "
             f"({sms} SMs × 4 schedulers) ÷ 1.965 GHz = **{floor_ms:.3f} ms**. ncu measured {dur_us:.1f} µs for this launch (cold, serialised) ⇒ the pipe "
             f"is busy {100 * floor_ms / (dur_us * 1e-3):.0f} % of the time by this model; ncu's own `sm__pipe_fp64_cycles_active` reads {pipe:.1f} % "
             "(it counts 2 cycles per instruction, not the third cycle of a three-pair DFMA).\n")
lines.append("Against the pipe's arithmetic peak the same launch is "
             f"{578.0 * CELLS / (dur_us * 1e-6) / 1e12:.2f} TFLOP/s algorithmic of 37.22 = {578.0 * CELLS / (dur_us * 1e-6) / 37.22e12:.2f} "
             "(bench.py's `roofline.frac` uses the CUDA-event time of the warm back-to-back loop, a few % shorter).\n")
lines.append("## 3. Where the three-pair DFMAs are, and what was done about them\n")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "astroz_b200", "libastroz_b200.so")], capture_output=True, text=True).stdout
fn = []
on = False
for ln in sass.splitlines():
    if "Function :" in ln:
        on = re.search(r"sgp4_grid_kernelILi0ELi0ELb1ELi4ELi384ELi[23]ELi3ELi0E", ln) is not None
        continue
    if on and re.search(r"/\*[0-9a-f]{4}\*/", ln):
        fn.append(ln.split("*/", 1)[1].split(";")[0].strip())
# first Horner run with immediates
start = next((i for i, l in enumerate(fn) if "DFMA" in l and "0.0083333" in l), None)   # the 1/120 immediate of sincos_full
excerpt = fn[start - 4:start + 26] if start else fn[:30]
lines.append("Round 1's kernel kept every polynomial coefficient in a register (`LDC.64` then `DFMA p, z, p, Rc`): each Horner step of the four "
             "range-reduced sincos was a three-pair instruction. Round 2 first re-fitted those kernels with immediate coefficients "
             "(`tools/fit_sincos_imm.py`; an fp64 immediate must have a zero low word), then replaced the reduction itself: `sincos_full` "
             "now reduces to a 1024-point lattice of the circle, reads (sin, cos) of the lattice point from a 16 KB table and finishes with a "
             "two-term sine / two-term cosine of the remainder (|r| <= 3.1e-3; 1/120, 1/24, -1/2 and the head of 2π/1024 are immediates) and "
             "one angle addition: 14 fp64 instructions instead of 18, no quadrant selects, 12 fewer registers -- which let a third CTA onto "
             "the SM (`profiles/r02w_sincos_table.jsonl`). Excerpt of the shipped SASS around one such evaluation (three epochs per thread "
             "interleaved; note the immediates, the `.reuse` flags and the `LDG.E.128` of the table entry):\n")
lines.append("```\n" + "\n".join(excerpt) + "\n```\n")
lines.append("What remains are products of three live per-cell quantities — the angle additions `fma(S, cr, C*sr)` and rotations "
             "`fma(s0, cd, c0*sd)`, `fma(axnl, s, -(aynl*c))`, `fma(rate, t, angle0)` with per-satellite operands — not constants.\n")
lines.append("## 4. Why not 100 % of the floor\n")
lines.append("Three CTAs of four warps per SM (three resident warps per scheduler at 165 registers) × three epochs per thread give ≤ 9 independent chains against a dependent-issue "
             "latency of 8.1 cycles and a 2–3-cycle issue interval: enough inside the long polynomial blocks (18–20 stall samples per "
             "instruction) but not inside short dependent phases — reciprocal seeds (MUFU then two dependent FMAs), the quadrant selects "
             "between a sincos' polynomials and its consumers, the Newton steps. The launch-shape sweep (`profiles/r02m_k1_shapes.jsonl`: "
             "2 or 3 epochs per thread × 2–4 CTAs/SM × stripes 256–768) is flat within 2 %, i.e. trading chains for warps does not help; "
             "round 2 instead removed control flow from the hot path (speculative two-step Kepler solve and small rotations with cold "
             "fall-backs), which merged the short blocks and moved the kernel from 0.437 to 0.40 ms; the shorter sincos kernels and the Kepler hand-off "
             "(`profiles/r02r_kepler_handoff.jsonl`) brought it to 0.374 ms, the lattice-reduced sincos with a third resident CTA to "
             "0.347 ms warm (bench.py) = 0.868 of the arithmetic peak on the algorithmic FLOP count; by this section's model the pipe "
             "is then busy ~80 % of the time.\n")
open(out, "w").write("\n".join(lines))
print("wrote", out)
