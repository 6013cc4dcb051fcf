"""Per-source-line instruction counts and stall samples from an ncu report:
   python scripts/ncu_lines.py report.ncu-rep <units per launch> [top]"""
import csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr, out = None, None, []
for r in csv.reader(txt.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and cur and len(r) == len(hdr) and r[0] not in ("", "-"):
        try:
            ie, smp = float(r[hdr.index("Instructions Executed")]), float(r[hdr.index("# Samples")])
        except ValueError:
            continue
        stalls = {h: float(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v not in ("", "-") and float(v) > 0}
        out.append((cur, int(r[0]), ie / units, smp, r[1].strip()[:80], stalls))
ti, ts = sum(o[2] for o in out), sum(o[3] for o in out)
print("warp instructions per unit: %.1f   samples: %d" % (ti, ts))
for o in sorted(out, key=lambda x: -x[3])[:top]:
    st = ",".join("%s %.0f" % (k[6:], v) for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3])
    print("%-16s %5d %6.1f i/u %5.1f%%  %-80s %s" % (o[0], o[1], o[2], 100 * o[3] / ts, o[4], st))
