"""Per-source-line instruction / stall-sample shares of one launch in an .ncu-rep
(ncu -i REP --page source --csv --print-source cuda,sass).  usage: ncu_source_hotspots.py REP LAUNCH_INDEX [N]"""
import csv, io, subprocess, sys
rep, launch = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", launch, "--launch-count", "1"], capture_output=True, text=True).stdout
cur, hdr, out, fn = None, None, [], ""
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1][:80]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            out.append((cur, int(r[0]), int(r[7]), int(r[8]), int(r[6]), r[1].strip()[:96]))
        except ValueError:
            pass
ti, tt, ts = (sum(o[k] for o in out) or 1 for k in (2, 3, 4))
print(f"{fn}\nwarp instructions {ti}, thread instructions {tt} ({tt / ti:.1f} active lanes/instruction), samples {ts}")
by = {}
for o in out:
    by[o[0]] = by.get(o[0], 0) + o[2]
print("by file:", {k: f"{100 * v / ti:.1f}%" for k, v in by.items()})
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print(f"{o[0]:16s}:{o[1]:4d} inst {100 * o[2] / ti:5.1f}%  samples {100 * o[4] / ts:5.1f}%  lanes {o[3] / max(o[2], 1):4.1f}  {o[5]}")
