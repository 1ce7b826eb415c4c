"""Turn an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of ONE
forward step (tools/ncu_step.py) into profiles/traffic.json, which bench.py reports as roofline.traffic.

Usage: python tools/summarise_ncu_launches.py gpurun_out/<launches>.csv profiles/<copy>.csv [batch]"""
import csv
import json
import os
import shutil
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 512
lines = [l for l in open(src) if not l.startswith("==")]
with open(dst, "w") as f:
    f.writelines(lines)
rows = list(csv.DictReader(lines))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
k = OrderedDict()
for r in rows:
    e = k.setdefault(r["ID"], {"name": r["Kernel Name"]})
    e[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]]
conv = [e for e in k.values() if any(t in e["name"] for t in ("conv_gemm", "chain_gemm", "pair_chain", "conv3x3_tap3", "l1_block", "stem_rows", "stem_fused"))]
tot = lambda es, m: sum(e[m] for e in es)  # noqa: E731
out = {
    "source": os.path.relpath(dst, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
    "batch": batch,
    "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
           "one forward of tools/ncu_step.py (second forward; cold-cache, serialised launches)",
    "conv_launches": len(conv),
    "conv_dram_bytes_per_step": tot(conv, "dram__bytes_read.sum") + tot(conv, "dram__bytes_write.sum"),
    "conv_dram_read_bytes_per_step": tot(conv, "dram__bytes_read.sum"),
    "conv_dram_write_bytes_per_step": tot(conv, "dram__bytes_write.sum"),
    "conv_ms_under_ncu": tot(conv, "gpu__time_duration.sum"),
    "all_launches": len(k),
    "all_dram_bytes_per_step": tot(k.values(), "dram__bytes_read.sum") + tot(k.values(), "dram__bytes_write.sum"),
    "all_ms_under_ncu": tot(k.values(), "gpu__time_duration.sum"),
}
with open(os.path.join(os.path.dirname(dst), "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
