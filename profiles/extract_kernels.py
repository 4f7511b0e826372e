"""Summarise an `ncu --set full` report per kernel into JSON (dram traffic per launch, duration, issue utilisation, ...):

    python profiles/extract_kernels.py gpurun_out/prof.ncu-rep profiles/r01_kernels.json

bench.py reads the JSON to fill `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel)."""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
WANT = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    agg = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("gsr::<unnamed>::", "").strip()
        d = agg.setdefault(name, {"launches": 0})
        d["launches"] += 1
        for m, key in WANT.items():
            if m in hdr:
                j = hdr.index(m)
                try:
                    v = float(r[j].replace(",", "")) * UNIT.get(units[j], 1.0)
                except ValueError:
                    continue
                d[key] = d.get(key, 0.0) + v
    for d in agg.values():
        n = d["launches"]
        for k in list(d):
            if k != "launches":
                d[k] = round(d[k] / n, 4)
        d["dram_traffic_bytes"] = round(d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0))
    json.dump(agg, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(agg, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
