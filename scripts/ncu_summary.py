"""Turns ncu output brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv          > profiles/rNN_launches.md
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep [regex]       > profiles/rNN_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("fi::<unnamed>::", "").replace("unnamed>::", "").strip()


def launches(path):
    """Launch list of a `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` pass: time per
    kernel and, when the DRAM byte counters were collected too, the DRAM GB/s each kernel sustained."""
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    have_bytes = False
    for row in csv.DictReader(lines):
        m = row.get("Metric Name")
        if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v = float(row["Metric Value"].replace(",", ""))
        key = (short(row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0, 0.0])
        if m == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += v * scale_t.get(row["Metric Unit"], 1.0)
        else:
            have_bytes = True
            a[2] += v * scale_b.get(row["Metric Unit"], 1.0)
    tot = sum(v[1] for v in agg.values())
    extra_h = " DRAM MB / launch | DRAM GB/s |" if have_bytes else ""
    print(f"| kernel | grid | block | launches | total us | share | avg us |{extra_h}\n|---|---|---|---:|---:|---:|---:|" + ("---:|---:|" if have_bytes else ""))
    for (k, g, b), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f" {v[2] / v[0] / 1e6:.2f} | {v[2] / (v[1] * 1e-6) / 1e9:.0f} |" if have_bytes else ""
        print(f"| `{k}` | {g} | {b} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |{extra}")
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (ncu per-launch times: cold-cache, serialised)")


def full(path, pattern=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def volume(r):
        return -eval(r[ix["Grid Size"]].replace("(", "").replace(")", "").replace(",", "*").replace(" ", "") or "0")

    # one launch per kernel name: the one with the largest grid (the first of them), i.e. the finest level of a hierarchy
    best = {}
    for n, r in enumerate(rows[2:]):
        name = short(r[ix["Kernel Name"]])
        if name not in best or volume(r) < volume(rows[2 + best[name]]):
            best[name] = n
    for n, r in enumerate(rows[2:]):
        name = short(r[ix["Kernel Name"]])
        if pattern and not re.search(pattern, name):
            continue
        if best[name] != n:
            continue
        print(f"\n### `{name}`  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in ix:
                print(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            cur = {"name": short(r[1]), "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None:
            cur["rows"].append(r)
    for n, s in enumerate(secs):
        if (pattern and not re.search(pattern, s["name"])) or best.get(s["name"]) != n or not s["hdr"]:
            continue
        h = s["hdr"]
        ix = {k: i for i, k in enumerate(h)}
        stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        tot = {k: 0 for k in stalls}
        ns = 0
        good = [r for r in s["rows"] if len(r) >= len(h)]
        for r in good:
            for k in stalls:
                tot[k] += int(r[ix[k]] or 0)
            ns += int(r[ix["# Samples"]] or 0)
        print(f"\n### warp-stall samples, `{s['name']}` ({ns} samples)\n")
        print(", ".join(f"{k[6:]} {100 * v / max(ns, 1):.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        print("\n| samples | SASS | dominant stall |\n|---:|---|---|")
        for r in sorted(good, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
            st = {k: int(r[ix[k]] or 0) for k in stalls}
            m = max(st, key=st.get)
            print(f"| {r[ix['# Samples']]} | `{r[ix['Source']].strip()[:80]}` | {m[6:]} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
