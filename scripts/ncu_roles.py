"""Per-role warp-stall breakdown of mlp_tc_kernel from an ncu report: splits the SASS of the kernel at its
setmaxnreg instructions (one per warp role) and sums the stall samples / executed instructions per region.

  python scripts/ncu_roles.py gpurun_out/r87_fwd.ncu-rep [launch_index]
"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:mlp_tc",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[h]
si, ai, ii, so = hdr.index("# Samples"), hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
stalls = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
data, seen, base = [], set(), None
for r in rows[h + 1:]:
    try:
        ad = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    except (ValueError, IndexError):
        continue
    if ad in seen:
        continue
    seen.add(ad)
    base = ad if base is None else base
    data.append((ad - base, int(r[si] or 0), int(r[ii] or 0), r[so], {c: int(r[i] or 0) for i, c in stalls}))
marks = [0] + [d[0] for d in data if "USETMAXREG" in d[3]] + [1 << 40]
print("kernel:", rows[0][1][:90] if rows and len(rows[0]) > 1 else "")
print("total samples", sum(d[1] for d in data), "instructions", sum(d[2] for d in data))
for a, b in zip(marks[:-1], marks[1:]):
    d = [x for x in data if a <= x[0] < b]
    c = collections.Counter()
    for x in d:
        for k, v in x[4].items():
            c[k] += v
    print(f"== [{a:#x}, {b:#x}) samples {sum(x[1] for x in d)} instructions {sum(x[2] for x in d)}")
    print("   " + ", ".join(f"{k}:{v}" for k, v in c.most_common(8) if v))
    for x in sorted(d, key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 8]:
        st = max(x[4].items(), key=lambda kv: kv[1])
        print(f"   {x[0]:#7x} {x[1]:5d} inst={x[2]:8d} {x[3][:64]:64s} {st[0]}:{st[1]}")
