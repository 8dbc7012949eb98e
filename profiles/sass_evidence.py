#!/usr/bin/env python
"""Static SASS mnemonic counts per kernel of the built library (no GPU needed): which instructions carry the stores, the cell
fetches, the MUFU approximations, the half packing, the TMA bulk copies and the peer hand-shake.
usage: python profiles/sass_evidence.py > profiles/r02/sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "godot_atmosphere_shader_b200", "libb200atmo.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
WANT = ["UBLKCP", "STG.E.128.STRONG.SYS", "STG.E.64.STRONG.SYS", "STG.E.STRONG.SYS", "STG.E.EF.128", "STG.E.EF.64", "STG.E.128", "STG.E.64", "LDG.E.128.CONSTANT",
        "LDG.E.EF.128", "LDG.E.STRONG.SYS", "ATOMG", "MEMBAR", "MUFU.EX2", "MUFU.RSQ", "MUFU.RCP", "FRND", "F2FP", "FFMA", "FADD", "FMUL", "LDCU.128", "LDCU.64", "LDCU", "LDC.64", "LDC",
        "BAR.SYNC", "NANOSLEEP", "SHFL", "VOTE", "LDS", "STS"]
print("# cuobjdump -sass godot_atmosphere_shader_b200/libb200atmo.so (sm_100a), selected mnemonics per kernel (static counts)")
print("# UBLKCP.G.S = TMA bulk copy shared->global (cp.async.bulk); STG.E.*.STRONG.SYS = multimem.st / st.release.sys (peer flags);")
print("# STG.E.EF.* = streaming result store (st.global.cs: .128 float4, .64 half4); LDG.E.128.CONSTANT = one 16-byte cell fetch (LUT / cube /")
print("# shape); MUFU.* = ex2 / rsqrt / rcp approximations; F2FP = fp32 -> packed half (RTN-even); ATOMG + MEMBAR + LDG.E.STRONG.SYS + NANOSLEEP =")
print("# the hand-shake fused into the peers kernels (B200AtmoPeerSync); LDCU.128 / .64 = packed constant loads (CloudHot)\n")
blocks = re.split(r"\s+Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    ops = collections.Counter()
    n = 0
    for line in blk.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + ".") or (w in ("UBLKCP", "ATOMG", "MEMBAR", "FRND", "F2FP", "SHFL", "VOTE") and op.startswith(w)):
                ops[w] += 1
                break
    short = name.replace("b200atmo::", "").replace("(int)", "").replace("(bool)", "")
    print(f"{short}   [{n} instr]")
    print("    " + ", ".join(f"{w} x{ops[w]}" for w in WANT if ops[w]))
