"""Instruction mix of the integrator's step loops from the shipped library (cuobjdump -sass): finds the backward branches
of a kernel, takes the two largest loop bodies (the 24-node and the 18-node copy of the Euler step) and counts opcodes by class.

    python tools/sass_mix.py [mangled-kernel-substring] > profiles/rNN_k1_sass_mix.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("NREM_LIB_PATH") or os.path.join(ROOT, "nremmodfc_b200", "csrc", "libnremfc.so")
KERNEL = sys.argv[1] if len(sys.argv) > 1 else "wc_batch_tc_kernelILi3ELi24ELb1ELb1ELi0EE"

CLASSES = [
    ("MUFU (XU pipe)", r"^MUFU"),
    ("FP32 FFMA/FMUL/FADD (FMA pipe)", r"^(FFMA|FMUL|FADD|FFMA2|FMUL2|FADD2)"),
    ("integer multiply IMAD* (FMA pipe)", r"^IMAD"),
    ("integer/logic ALU (IADD3, LOP3, SHF, LEA, PRMT, SEL, MOV ...)", r"^(IADD3|IADD|LOP3|SHF|LEA|PRMT|SEL|MOV|VIADD|ISETP|FSETP|PLOP3|CS2R|S2R|FSEL|FMNMX|I2F|F2I|IABS|UMOV|R2UR|FLO|POPC|P2R|R2P)"),
    ("shared memory LDS/STS", r"^(LDS|STS)"),
    ("tensor memory LDTM/STTM", r"^(LDTM|STTM)"),
    ("tcgen05.mma UTCHMMA / commit UTCBAR", r"^(UTCHMMA|UTCBAR|UTCQMMA)"),
    ("global LDG/STG", r"^(LDG|STG|LD\.|ST\.)"),
    ("local (spill) LDL/STL", r"^(LDL|STL)"),
    ("barriers / fences / mbarrier (BAR, SYNCS, FENCE, MEMBAR, WARPSYNC, ELECT)", r"^(BAR|SYNCS|FENCE|MEMBAR|WARPSYNC|ELECT|NANOSLEEP|DEPBAR|ERRBAR|CCTL)"),
    ("uniform datapath (U*)", r"^U[A-Z]"),
    ("branches", r"^(BRA|BSSY|BSYNC|EXIT|RET|CALL|BREAK|WARPSYNC)"),
]


def main():
    names = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    # split per function
    funcs = re.split(r"\n\s*Function : ", sass)
    body = next((f for f in funcs if KERNEL in f.split("\n", 1)[0]), None)
    if body is None:
        sys.exit(f"kernel containing {KERNEL!r} not found")
    fname = body.split("\n", 1)[0].strip()
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            text = m.group(2).strip()
            text = re.sub(r"^@!?U?P\d+\s+", "", text)
            ins.append((int(m.group(1), 16), text))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    loops = []
    for k, (a, t) in enumerate(ins):
        m = re.match(r"BRA(?:\.[A-Z.]+)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_index:
                loops.append((k - addr_index[tgt] + 1, addr_index[tgt], k))
    loops.sort(reverse=True)
    # outermost distinct bodies only
    picked = []
    for n, lo, hi in loops:
        if all(not (lo >= plo and hi <= phi) for _, plo, phi in picked):
            picked.append((n, lo, hi))
    print(f"# SASS instruction mix of the step loops of `{fname}`")
    print(f"\n`cuobjdump -sass {os.path.relpath(LIB, ROOT)}`, loop bodies = backward-branch ranges; counts are static instructions per warp and Euler step.\n")
    whole = collections.Counter()
    for n, lo, hi in picked[:2]:
        cnt, other = collections.Counter(), collections.Counter()
        for _, t in ins[lo:hi + 1]:
            op = t.split()[0]
            for label, rx in CLASSES:
                if re.match(rx, op):
                    cnt[label] += 1
                    break
            else:
                other[op.split(".")[0]] += 1
            whole[op.split(".")[0]] += 1
        print(f"## loop at 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {n} instructions\n")
        print("| class | count | share |\n|---|---:|---:|")
        for label, _ in CLASSES:
            if cnt[label]:
                print(f"| {label} | {cnt[label]} | {100.0 * cnt[label] / n:.1f} % |")
        if other:
            print(f"| other ({', '.join(f'{k} {v}' for k, v in other.most_common(8))}) | {sum(other.values())} | {100.0 * sum(other.values()) / n:.1f} % |")
        ops = collections.Counter(t.split()[0].split(".")[0] for _, t in ins[lo:hi + 1])
        print("\ntop opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)) + "\n")


if __name__ == "__main__":
    main()
