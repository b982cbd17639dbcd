"""Hot call sites of the __noinline__ device helpers, from an `ncu --page source --print-source cuda,sass --csv` dump:
every executed CALL instruction with the source line it sits on, the source line of its target, how often it ran and with
how many threads.    python tools/ncu_calls.py dump.csv [top]"""
import csv
import sys

top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file, cur_line, cur_src = None, None, ""
addr_line, calls = {}, []
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0].isdigit():
        cur_line, cur_src = int(r[0]), r[1].strip()[:70]
    elif r[0] == "" and len(r) > 8 and r[2].startswith("0x"):
        addr_line.setdefault(int(r[2], 16), (cur_file, cur_line, cur_src))
        if "CALL" in r[3]:
            tgt = r[3].split()[-1]
            try:
                calls.append((int(r[7]), int(r[8]), cur_file, cur_line, cur_src, int(tgt, 16) if tgt.startswith("0x") else tgt))
            except ValueError:
                pass
total = sum(c[0] for c in calls)
print("executed CALLs", total)
for n, tn, f, l, src, tgt in sorted(calls, reverse=True)[:top]:
    t = addr_line.get(tgt) if isinstance(tgt, int) else None
    tname = f"{t[0]}:{t[1]} {t[2][:40]}" if t else str(tgt)
    print(f"{n:9d} calls thr {tn / max(n, 1):5.1f}  {f}:{l} | {src[:60]}  ->  {tname}")
