"""Print the key numbers of a bench.py JSON line: python scripts/print_bench.py gpurun_out/xx_bench.log"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"] / 1e6, 2), "M; e2e", round(d["e2e"]["value"] / 1e6, 2), "M")
if d.get("roofline"):
    print("roofline:", d["roofline"]["kernel"][:50], round(d["roofline"]["kernel_ms"] * 1e3, 1), "us frac", round(d["roofline"]["frac"], 3))
for o in d.get("roofline_other_kernels") or []:
    print("   other:", o["kernel"][:60], round(o["kernel_ms"] * 1e3, 1), "us frac", round(o["frac"], 3))
