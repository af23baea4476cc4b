"""Static look at one kernel's SASS: opcode histogram plus the sum of the issue-stall fields of its control codes
(bits 105..108 of every instruction), i.e. the cycles ONE warp needs to issue the code once if nothing else waits.

    python tools/sass_stalls.py gpu_matrix_inversion_b200/csrc/gj_batched.o batched_pk_kernelILi64ELi4 [--dump]
"""
import collections
import re
import subprocess
import sys


def load(obj, needle):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, keep = [], False
    for line in names.split("\n"):
        if "Function :" in line:
            keep = needle in line
        if keep:
            out.append(line)
    return out


def decode(lines):
    pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/")
    pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]+) \*/")
    ins, i = [], 0
    while i < len(lines):
        m = pat.match(lines[i])
        if m and i + 1 < len(lines):
            m2 = pat2.match(lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 46) & 7, (hi >> 49) & 7,
                            (hi >> 52) & 0x3F))
                i += 2
                continue
        i += 1
    return ins


def opcode(text):
    t = text.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


if __name__ == "__main__":
    ins = decode(load(sys.argv[1], sys.argv[2]))
    lo = int(sys.argv[sys.argv.index("--from") + 1], 16) if "--from" in sys.argv else 0
    hi = int(sys.argv[sys.argv.index("--to") + 1], 16) if "--to" in sys.argv else 1 << 30
    ins = [x for x in ins if lo <= x[0] <= hi]
    cnt, st = collections.Counter(), collections.Counter()
    for a, text, stall, wb, rb, wm in ins:
        cnt[opcode(text)] += 1
        st[opcode(text)] += stall
    print(f"{len(ins)} instructions, stall sum {sum(st.values())}")
    for op, c in cnt.most_common(24):
        print(f"  {op:14s} {c:5d}  stall {st[op]:5d}")
    if "--dump" in sys.argv:
        for a, text, stall, wb, rb, wm in ins:
            print(f"{a:05x} {stall:2d} {'w%d' % wb if wb < 7 else '  '} {'r%d' % rb if rb < 7 else '  '} {wm:06b} {text[:110]}")
