"""The SASS of the hot loop out of an `ncu --page source --csv` dump: every instruction whose executed count is within
a factor of the most-executed one, in address order, with its share of warp-level instructions and of stall samples.
    python profiles/extract_node_step.py gpurun_out/<name>_source.csv [min share of max, default 0.6] > profiles/<name>.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
kernel = rows[0][1]
hdr = rows[1]
ia, isrc, ismp, iex, ithr = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed"))
body = [r for r in rows[2:] if len(r) > iex and r[iex].replace(",", "").isdigit()]
ex = [int(r[iex].replace(",", "")) for r in body]
smp = [int(r[ismp].replace(",", "") or 0) for r in body]
tot_ex, tot_s, mx = sum(ex), sum(smp), max(ex)
print("kernel:", kernel)
print("total warp instructions %d, stall samples %d; listed: instructions executed >= %.0f %% of the most executed one" % (tot_ex, tot_s, 100 * frac))
n = 0
share = 0.0
for r, e, s in zip(body, ex, smp):
    if e >= frac * mx:
        n += 1
        share += e / tot_ex
        print(" %s  %-58s exec %5.2f%%  samples %5.2f%%  thr %s" % (r[ia][-5:], r[isrc].strip()[:58], 100 * e / tot_ex, 100 * s / max(tot_s, 1), r[ithr]))
print("%d instructions in the listed loop body = %.1f %% of all executed warp instructions" % (n, 100 * share))
