#!/usr/bin/env python
"""Static statistics of the march loop of render_rays_kernel<V2, no clouds> from the SASS of the built library: opcode mix,
pipe classes, register-file operand reads per instruction (uniform-register, constant and immediate operands do not use
register-file ports; `.reuse` operands come from the operand-reuse cache). Evidence for DESIGN.md §5.1 (no GPU needed).
usage: python profiles/sass_loop_stats.py > profiles/r01/sass_loop_stats.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "godot_atmosphere_shader_b200", "libb200atmo.so")
KERNEL = "_ZN8b200atmo18render_rays_kernelILi0ELi0ENS_5RayIOELb0EEEvNS_9DevConstsET1_"

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
ins, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = KERNEL in line
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = [(a, int(re.search(r"0x([0-9a-f]+)", b).group(1), 16)) for a, b in ins
        if re.search(r"\bBRA", b) and re.search(r"0x([0-9a-f]+)", b) and int(re.search(r"0x([0-9a-f]+)", b).group(1), 16) < a]
end, start = back[0]   # first back edge = the 8x unrolled main loop
body = [b for a, b in ins if start <= a <= end]
print(f"kernel {KERNEL}")
print(f"{len(ins)} SASS instructions; main march loop {start:#x}..{end:#x}: {len(body)} instructions for 8 steps = {len(body) / 8:.2f} per step")

PIPE = {"FFMA": "fma", "FADD": "fma", "FMUL": "fma", "IMAD": "fma", "MUFU": "xu", "LDG": "lsu", "VIADDMNMX": "alu", "MOV": "alu",
        "LDC": "lsu/const", "UIADD3": "uniform", "UISETP": "uniform", "BRA": "branch"}
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", b).split()[0] for b in body)
pipes = collections.Counter()
for op, n in ops.items():
    pipes[PIPE.get(op.split(".")[0], "other")] += n
print("opcodes:", ", ".join(f"{k} x{v}" for k, v in ops.most_common()))
print("pipes  :", ", ".join(f"{k} {v} ({100 * v / len(body):.0f} %)" for k, v in pipes.most_common()))

reads = reuse = uni = const = imm = 0
hist = collections.Counter()
for b in body:
    m = re.match(r"(?:@!?U?P\d+\s+)?(\S+)\s+(.*)", b)
    operands = [o.strip() for o in m.group(2).split(",")][1:]   # drop the destination
    r = 0
    for o in operands:
        if re.search(r"\bUR\d+", o):
            uni += 1
        elif re.search(r"c\[0x", o):
            const += 1
        elif re.search(r"\bR\d+", o):
            if ".reuse" in o:
                reuse += 1
            else:
                r += 1
        elif re.search(r"[0-9]", o):
            imm += 1
    reads += r
    hist[r] += 1
print(f"register-file source operands: {reads} (+{reuse} served by .reuse) = {reads / len(body):.2f} per instruction; "
      f"uniform-register {uni}, constant-bank {const}, immediate {imm}")
print("instructions by number of register-file reads:", ", ".join(f"{k}: {v}" for k, v in sorted(hist.items())))
