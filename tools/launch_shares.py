"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of one step.

    python tools/launch_shares.py gpurun_out/launches.csv [step_index_from_end]
A step is delimited by consecutive launches of the forward loss kernel (mae_loss_kernel<false>)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    recs = list(csv.DictReader(lines))

    def dur_us(x):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        return v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)

    names = [x["Kernel Name"] for x in recs]
    marks = [i for i, n in enumerate(names) if "mae_loss_kernel<0" in n or "mae_loss_kernel<false" in n or
             ("mae_loss_kernel" in n and "(bool)0" in n)]
    if len(marks) < back + 1:
        marks = [i for i, n in enumerate(names) if "loss_reduce_kernel" in n]
    lo, hi = marks[-back - 1], marks[-back]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for x in recs[lo:hi]:
        n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("(anonymous namespace)::", "")
        agg[n][0] += 1
        agg[n][1] += dur_us(x)
    tot = sum(v[1] for v in agg.values())
    print(f"# one step = launches [{lo}, {hi}) of {len(recs)}: {hi - lo} launches, {tot / 1e3:.2f} ms summed kernel time "
          f"(ncu-serialised, cold caches: compare shares, not absolutes)")
    print(f"{'ms':>9} {'share':>6} {'count':>6}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:9.3f} {100 * t / tot:5.1f}% {c:6d}  {n[:120]}")


if __name__ == "__main__":
    main()
