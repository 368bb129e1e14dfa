"""Summarise an ncu report (.ncu-rep) of the hot-path kernels into profiles/<tag>_ncu_summary.{md,json}.

  python scripts/summarize_ncu.py gpurun_out/r12_prof.ncu-rep r01_edge_node

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  One row per captured launch:
duration, SM clock, DRAM bytes read/written (= roofline.traffic), DRAM %, tensor-pipe %, issue-slot %,
registers, plus the warp-stall mix of the first launch from the source page.
"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def run(args):
    return subprocess.run(args, capture_output=True, text=True, check=True).stdout


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]]}
        for k, name in KEYS:
            if k in col:
                try:
                    d[name] = float(r[col[k]].replace(",", ""))
                except ValueError:
                    d[name] = r[col[k]]
                if name.endswith("_MB") and units[col[k]].lower().startswith("gbyte"):
                    d[name] *= 1e3
                if name.endswith("_MB") and units[col[k]].lower().startswith("kbyte"):
                    d[name] /= 1e3
        d["dram_traffic_bytes"] = int((d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0)) * 1e6)
        launches.append(d)
    # stall mix of the first launch
    src = list(csv.reader(run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
    stall = collections.Counter()
    h = None
    nk = 0
    for r in src:
        if r and r[0] == "Kernel Name":
            nk += 1
            if nk == 2:
                break
            continue
        if r and r[0] == "Address":
            h = {x: i for i, x in enumerate(r)}
            continue
        if h and len(r) == len(h):
            for k, i in h.items():
                if k.startswith("stall_") and "Not Issued" not in k and r[i]:
                    stall[k] += int(r[i])
    tot = sum(stall.values()) or 1
    mix = {k: round(100.0 * v / tot, 1) for k, v in stall.most_common(8)}
    out = {"report": os.path.basename(rep), "launches": launches, "stall_mix_pct_first_launch": mix}
    json.dump(out, open(os.path.join(root, "profiles", tag + "_ncu_summary.json"), "w"), indent=1)
    with open(os.path.join(root, "profiles", tag + "_ncu_summary.md"), "w") as f:
        f.write(f"# ncu --set full summary ({os.path.basename(rep)})\n\n")
        f.write("| kernel | grid x block | regs | duration us | SM GHz | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | issue active % | L2 hit % |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for d in launches:
            f.write(f"| {d['kernel'][:48]} | {int(d.get('grid', 0))} x {int(d.get('block', 0))} | {int(d.get('registers', 0))} | "
                    f"{d.get('duration_us', 0):.1f} | {d.get('sm_ghz', 0):.2f} | {d.get('dram_read_MB', 0):.1f} | "
                    f"{d.get('dram_write_MB', 0):.1f} | {d.get('dram_pct', 0):.1f} | {d.get('tensor_pipe_pct', 0):.1f} | "
                    f"{d.get('issue_active_pct', 0):.1f} | {d.get('l2_hit_pct', 0):.1f} |\n")
        f.write("\nWarp-stall mix of the first launch (all samples, %): " + ", ".join(f"{k[6:]} {v}" for k, v in mix.items()) + "\n")
    print(open(os.path.join(root, "profiles", tag + "_ncu_summary.md")).read())


if __name__ == "__main__":
    main()
