"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown).
   python scripts/launch_shares.py gpurun_out/r29_launches.csv [title]"""
import collections, csv, sys

def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1000, "us": v, "ms": v * 1000}.get(row["Metric Unit"], v)
        k = (row["Kernel Name"].split("(")[0][:64], row["Grid Size"])
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# Launch list shares: {title}\n")
    print("Cold-cache serialised per-launch times (ncu): compare SHARES with the CUDA-event numbers, not absolutes.\n")
    print("| kernel | grid | launches | avg us | share % |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print(f"| {k[0]} | {k[1]} | {a[0]} | {a[1]/a[0]:.1f} | {100*a[1]/tot:.1f} |")
    print(f"\ntotal {tot/1000:.2f} ms over {sum(a[0] for a in agg.values())} launches")

main()
