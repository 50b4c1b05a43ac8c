"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON committed under profiles/.

    ncu -i gpurun_out/r1_kernels.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv > profiles/r1_kernels_ncu.json

One row per distinct (kernel, grid, block) -- the first launch of each -- with the counters the
roofline discussion needs (DESIGN.md "Measured").
"""
import csv
import json
import re
import sys

COLS = {
    "duration_us": ("gpu__time_duration.sum", 1e-3),            # ns -> us
    "grid": ("launch__grid_size", 1),
    "block": ("launch__block_size", 1),
    "cluster": ("launch__cluster_size", 1),
    "regs": ("launch__registers_per_thread", 1),
    "dram_read_bytes": ("dram__bytes_read.sum", 1),
    "dram_write_bytes": ("dram__bytes_write.sum", 1),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "dram_throughput_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "issue_active_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "tensor_pipe_active_pct": ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    "inst_executed": ("smsp__inst_executed.sum", 1),
    "l2_read_bytes": ("lts__t_bytes_op_read.sum", 1),
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6,
              "ns": 1, "us": 1e3, "ms": 1e6}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        name = re.sub(r"\(anonymous namespace\)::|<?unnamed>::|bqa::|void |\(int\)", "", name)
        name = re.sub(r"\(.*$", "", name).strip()
        d = {"kernel": name}
        for key, (metric, scale) in COLS.items():
            if metric not in col:
                continue
            raw = r[col[metric]].replace(",", "")
            try:
                v = float(raw)
            except ValueError:
                continue
            u = units[col[metric]]
            if key.endswith("_bytes") or key == "duration_us":
                v *= UNIT_SCALE.get(u, 1)
            d[key] = round(v * scale, 3) if key == "duration_us" else v
        k = (name, d.get("grid"), d.get("block"))
        if k in seen:
            continue
        seen.add(k)
        out.append(d)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
