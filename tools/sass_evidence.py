"""SASS evidence for profiles/: mnemonic histogram and excerpts of one kernel of the built library.
usage: python tools/sass_evidence.py <substring of the mangled kernel name> > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "beat_this_b200", "libbeatthis_sm100.so")
want = sys.argv[1]
dump = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in dump.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1).strip())
name = [k for k in funcs if want in k][0]
ins = funcs[name]
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
keep = ("MUFU", "FADD2", "FFMA2", "FMUL2", "F2FP", "FMNMX3", "FMNMX", "IMAD", "SYNCS", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "STG", "LDG",
        "UTCBAR", "LDL", "STL")
print(f"# cuobjdump -sass beat_this_b200/libbeatthis_sm100.so (built from this tree): {name}")
print(f"# {len(ins)} instructions; " + ", ".join(f"{k} {ops[k]}" for k in keep if ops[k]))
print()


def excerpt(title, pred, before=6, after=34):
    for i, t in enumerate(ins):
        if pred(t):
            print(f"## {title}")
            for t2 in ins[max(0, i - before): i + after]:
                print("    " + t2)
            print()
            return


excerpt("softmax: packed subtract of the reference maximum (FADD2), MUFU.EX2 pairs, polynomial pairs (Cody-Waite split on FADD2, "
        "degree 3 on FFMA2, exponent insertion IMAD), pack to fp16 (F2FP), packed row sums (FADD2)", lambda t: t.startswith("FFMA2"), 14, 60)
excerpt("row maximum on 3-input FMNMX3", lambda t: t.startswith("FMNMX3"), 2, 12)
excerpt("issuer warp: TMA loads (UTMALDG), tcgen05.mma with P read from tensor memory (UTCHMMA ... tmem[...] A operand), commits (UTCBAR)",
        lambda t: "UTCHMMA" in t and "tmem[UR" in t.split(",")[0] + t.split(",")[1], 10, 16)
excerpt("softmax warps: tcgen05.ld of S (LDTM), tcgen05.st of P (STTM)", lambda t: "LDTM" in t, 2, 6)
excerpt("", lambda t: "STTM" in t, 2, 6)
