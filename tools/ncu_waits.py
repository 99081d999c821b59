"""mbarrier wait sites of one kernel in an ncu report: python tools/ncu_waits.py <report> <kernel regex>
For every SYNCS...TRYWAIT instruction: barrier operand, times executed, and the stall samples collected in the
spin loop around it (the instructions between it and the next TRYWAIT/arrive site are attributed to it)."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
blk = out.split('"Kernel Name",')[1]
lines = blk.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
recs = [r for r in rows[1:] if len(r) >= len(hdr)]
total = sum(int(r[col["# Samples"]] or 0) for r in recs)
sites = []
cur = None
for r in recs:
    src = r[col["Source"]]
    n = int(r[col["# Samples"]] or 0)
    if "TRYWAIT" in src:
        if cur and cur["first"] and src.split("[")[1].split("]")[0] == cur["op"] and len(cur["span"]) < 40:
            cur["span"].append((src, n)); cur["samples"] += n; cur["ex2"] = r[col["Instructions Executed"]]
            continue
        cur = {"addr": r[col["Address"]][-5:], "op": src.split("[")[1].split("]")[0], "ex": r[col["Instructions Executed"]], "samples": n, "span": [], "first": True}
        sites.append(cur)
    elif cur is not None and len(cur["span"]) < 34:
        cur["span"].append((src, n)); cur["samples"] += n
print("total samples", total)
for s in sites:
    print(f"{s['addr']}  [{s['op']:28s}] executed {s['ex']:>9s}  samples~{s['samples']:6d} ({100*s['samples']/total:4.1f}%)")
