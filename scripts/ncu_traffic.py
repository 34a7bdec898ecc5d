"""Writes profiles/ncu_traffic.json from an ncu CSV of the dominant kernel, stamped with the hash of the kernel's sources
(bench.py refuses a capture whose hash differs from the build it runs).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:stencil3d_tma_kernel --csv --log-file gpurun_out/traffic.csv python scripts/profile_step.py 512 12
    python scripts/ncu_traffic.py gpurun_out/traffic.csv sdf3d_512_1M f32 [out.json]
"""
import collections
import csv
import datetime
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path, workload, precision = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = collections.OrderedDict()
for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
    m = row.get("Metric Name")
    if m not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    a = per.setdefault((row["ID"], row["Kernel Name"]), {"r": 0.0, "w": 0.0})
    a["r" if m.endswith("read.sum") else "w"] += float(row["Metric Value"].replace(",", "")) * scale_b.get(row["Metric Unit"], 1.0)
# the fused PCG kernel (template arguments <T, R, S, Fused = 1, MINB, Epi, GS, TY>): skip the first launches (cold L2, beta = 0)
import re
fused = [(k, v) for k, v in per.items() if re.search(r"stencil3d_tma_kernel<\w+, \d+, \d+, (1|true|\(bool\)1),", k[1])]
if not fused:
    fused = [(k, v) for k, v in per.items() if "stencil3d_tma_kernel" in k[1]]
use = fused[2:] if len(fused) > 4 else fused
r = sum(v["r"] for _, v in use) / len(use)
w = sum(v["w"] for _, v in use) / len(use)
doc = json.load(open(out)) if os.path.exists(out) else {}
doc["_doc"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (mean over the launches of one ncu pass of "
               "scripts/profile_step.py); bench.py copies it into roofline.traffic when kernel_source_sha matches the build")
n = {"sdf3d_512_1M": 512, "sdf3d_256_1M": 256}.get(workload, 512)
B = 4 if precision == "f32" else 8
doc[f"{workload}:{precision}"] = {"kernel": use[0][0][1][:80], "launches_averaged": len(use), "dram_bytes_per_launch": r + w, "dram_read_bytes": r, "dram_write_bytes": w,
                                  "algorithmic_bytes_per_launch": 5 * B * n ** 3, "kernel_source_sha": bench.kernel_source_sha(),
                                  "captured": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"), "source": os.path.basename(path)}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc[f"{workload}:{precision}"], indent=1))
