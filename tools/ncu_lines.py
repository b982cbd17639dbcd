"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
    python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = csv.reader(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, fn, hdr, out = None, None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1][:60]
    elif r[0] == "Line No":
        if hdr is None and __import__("os").environ.get("HWY_LINES_HEADER"):
            print("columns:", r)
        hdr = r
    elif r[0].isdigit() and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((int(d["Instructions Executed"]), int(d["# Samples"]), cur_file, int(r[0]), r[1].strip()[:90],
                        (int(d.get("Thread Instructions Executed") or 0) / max(int(d["Instructions Executed"]), 1)
                         if d.get("Thread Instructions Executed") else float(d["Avg. Threads Executed"] or 0)), fn))
        except (ValueError, KeyError):
            pass
ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
tt = sum(o[0] * o[5] for o in out)
print("total inst", ti, "samples", ts, "thread inst", int(tt), "threads per warp inst %.2f" % (tt / max(ti, 1)))
# the instructions a warp would need if every line ran with all of the warp's busy lanes converged:
# sum over lines of thread instructions / (threads per warp instruction at the busiest line)
print("(thr = threads executing per warp instruction of the line)")
print("== by instructions")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]/ti*100:5.1f}% i {o[1]/ts*100:5.1f}% s thr {o[5]:4.0f} {o[2]}:{o[3]} | {o[4]}")
print("== by stall samples")
for o in sorted(out, key=lambda o: -o[1])[:top]:
    print(f"{o[0]/ti*100:5.1f}% i {o[1]/ts*100:5.1f}% s thr {o[5]:4.0f} {o[2]}:{o[3]} | {o[4]}")
