"""Per-source-line SASS instruction counts of one kernel inside a hot address range.
    python tools/sass_lines.py <cubin> <kernel substring> [lo hi]   (needs nvcc -lineinfo)"""
import collections, re, subprocess, sys
cubin, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
infn, line, cnt, ops = False, None, collections.Counter(), collections.defaultdict(collections.Counter)
for l in out.splitlines():
    if l.startswith(".text."):
        infn = pat in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        a = int(m.group(1), 16)
        if lo <= a <= hi:
            t = [x for x in m.group(2).split() if not x.startswith("@")]
            cnt[line] += 1
            ops[line][t[0].split(".")[0]] += 1
for k in sorted(cnt, key=lambda k: (k is None, k)):
    print(k, cnt[k], dict(ops[k].most_common(6)))
print("total", sum(cnt.values()))
