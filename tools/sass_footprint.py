"""Static code-size attribution of a kernel: bytes of SASS per source line inside an address range.
    nvdisasm -g <cubin> > all.dis ; cut the kernel's section into a file ; python tools/sass_footprint.py file lo hi [top]
Used to see which source constructs make the per-substep loop larger than the 32 KB L1.5 instruction cache."""
import collections
import re
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur = None
by_line = collections.Counter()
total = 0
for l in open(path):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", l)
    if m:
        a = int(m.group(1), 16)
        if lo <= a < hi:
            by_line[cur] += 16
            total += 16
print("range", hex(lo), hex(hi), "bytes", total)
src = {}
for (f, n), b in by_line.most_common(top):
    if f not in src:
        try:
            src[f] = open("/root/repo/highwayenv_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ""
    print(f"{b:6d} B {100.0 * b / total:5.1f}%  {f}:{n} | {text}")
# coarse: by file and 25-line bucket
buckets = collections.Counter()
for (f, n), b in by_line.items():
    buckets[(f, n // 25 * 25)] += b
print("== 25-line buckets")
for (f, n), b in sorted(buckets.items()):
    if b >= 256:
        print(f"{b:6d} B  {f}:{n}-{n + 24}")
