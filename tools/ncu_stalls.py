"""Top stall locations of one kernel from an ncu report:  python tools/ncu_stalls.py <report.ncu-rep> <kernel regex> [N]
Reads `ncu --page source --csv` (SASS view) and prints the N instructions with the most stall samples, plus totals per stall reason."""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for blk in blocks[1:2]:
    lines = blk.splitlines()
    print("kernel:", lines[0][:120])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    tot = {h: 0 for h in stall_cols}
    total_samples = 0
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        try:
            n = int(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        total_samples += n
        for h in stall_cols:
            try:
                tot[h] += int(r[col[h]] or 0)
            except ValueError:
                pass
        data.append((n, r))
    print("total samples", total_samples)
    print("by reason:", ", ".join(f"{h[6:]}={v} ({100*v/max(total_samples,1):.0f}%)" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
    data.sort(key=lambda t: -t[0])
    for n, r in data[:topn]:
        reasons = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
        print(f"{n:7d} {100*n/max(total_samples,1):5.1f}%  {r[col['Address']][-5:]}  {r[col['Source']][:70]:70s} ex={r[col['Instructions Executed']]:>8s} " +
              " ".join(f"{k}:{v}" for v, k in reasons if v))
