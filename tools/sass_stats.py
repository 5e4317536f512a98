"""Static SASS statistics of one kernel of libpml.so: opcode mix of the hottest loop (the largest
backward-branch span) -- a no-GPU estimate of issue slots / pipe cycles per row step.

    python tools/sass_stats.py <lib.so> <substring of mangled kernel name> [--all]
"""
import collections
import re
import subprocess
import sys

FMA_PIPE = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2", "FMUL2", "FADD2", "I2FP", "IMAD.WIDE"}
ALU_PIPE = {"IADD3", "LOP3", "SHF", "ISETP", "FSETP", "FSEL", "SEL", "FMNMX", "MOV", "LEA", "VIADD", "VIADDMNMX",
            "VIMNMX", "FSET", "IABS", "PRMT", "CS2R", "PLOP3", "P2R", "R2P", "IMNMX", "FCHK"}
LSU = {"LDG", "STG", "LDS", "STS", "SHFL", "ATOMG", "RED", "ATOMS", "LDL", "STL", "REDG", "LDSM"}
XU = {"MUFU", "F2I", "I2F", "FRND", "F2F"}


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, res = None, {}
    for l in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", l)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2)))
    return res


def opcode(text):
    t = [x for x in text.split() if not x.startswith("@")]
    return t[0].split(".")[0]


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    ks = kernels(lib)
    for name, ins in ks.items():
        if pat not in name:
            continue
        # largest backward branch span
        best = (0, 0, 0)
        for addr, text in ins:
            if opcode(text) == "BRA" and (text.lstrip().startswith("@") or "BRA.U" in text):   # loop back-edges
                m = re.search(r"0x([0-9a-f]+)", text)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt < addr and addr - tgt > best[0]:
                        best = (addr - tgt, tgt, addr)
        _, lo, hi = best
        if "--all" in sys.argv:
            lo, hi = 0, 1 << 30
        body = [(a, t) for a, t in ins if lo <= a <= hi]
        ops = collections.Counter(opcode(t) for _, t in body)
        n = len(body)
        packed = ops["FFMA2"] + ops["FMUL2"] + ops["FADD2"]
        fma = sum(c for o, c in ops.items() if o in FMA_PIPE)
        alu = sum(c for o, c in ops.items() if o in ALU_PIPE)
        lsu = sum(c for o, c in ops.items() if o in LSU)
        xu = sum(c for o, c in ops.items() if o in XU)
        print("%s\n  total %d instrs, loop [%#x, %#x] = %d instrs" % (name, len(ins), lo, hi, n))
        print("  fma-pipe %d (+%d packed second cycles = %d cycles)  alu-pipe %d (x2 = %d cycles)  lsu %d  xu %d  other %d"
              % (fma, packed, fma + packed, alu, 2 * alu, lsu, xu, n - fma - alu - lsu - xu))
        print("  " + "  ".join("%s %d" % oc for oc in ops.most_common(40)))


if __name__ == "__main__":
    main()
