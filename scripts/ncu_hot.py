"""Executed-instruction histogram of a kernel from an ncu report, in consecutive SASS blocks of `step` instructions:
   python scripts/ncu_hot.py report.ncu-rep [kernel-regex] [launch_skip] [step]
Shows where the issue slots go (instructions executed, stall samples, the dominant opcodes of each block)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "mlp_tc"
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
step = int(sys.argv[4]) if len(sys.argv) > 4 else 64
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[h]
si, ai, ii, so = hdr.index("# Samples"), hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
data, seen, base = [], set(), None
for r in rows[h + 1:]:
    try:
        ad = int(r[ai], 16)
    except (ValueError, IndexError):
        continue
    if ad in seen:
        continue
    seen.add(ad)
    base = ad if base is None else base
    data.append((ad - base, int(r[si] or 0), int(r[ii] or 0), r[so].strip()))
tot_i, tot_s = sum(d[2] for d in data), sum(d[1] for d in data)
print(f"total instructions {tot_i}, samples {tot_s}, SASS instructions {len(data)}")
for b in range(0, len(data), step):
    blk = data[b:b + step]
    ni, ns = sum(d[2] for d in blk), sum(d[1] for d in blk)
    if ni < tot_i * 0.004 and ns < tot_s * 0.004:
        continue
    ops = collections.Counter()
    for d in blk:
        op = d[3].split()[0] if not d[3].startswith("@") else d[3].split()[1]
        ops[op.split(".")[0]] += d[2]
    top = ", ".join(f"{k}:{100*v/max(ni,1):.0f}%" for k, v in ops.most_common(6))
    print(f"{blk[0][0]:#7x}-{blk[-1][0]:#7x} inst {ni:9d} ({100*ni/tot_i:4.1f}%) samples {ns:5d} ({100*ns/max(tot_s,1):4.1f}%)  {top}")
