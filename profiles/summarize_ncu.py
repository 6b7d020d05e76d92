"""Summarise an `ncu --set full` report into the per-kernel CSV committed under profiles/ (+ the DRAM traffic JSON that
bench.py's `roofline.traffic` reads).

    python profiles/summarize_ncu.py gpurun_out/prof_r2_bench.ncu-rep profiles/ncu_full_r2.csv profiles/traffic_r2.json
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_op_read_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def main(rep, out_csv, out_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in ix]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + cols)
        w.writerow([""] + [units[ix[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[ix["Kernel Name"]]] + [r[ix[c]] for c in cols])
    if out_json:
        def to_bytes(v, unit):
            v = float(v)
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

        agg = defaultdict(list)
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            if "gather_kernel" in name:
                agg["gather"].append((to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]),
                                      to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]),
                                      float(r[ix["gpu__time_duration.sum"]])))
        g = agg["gather"]
        if g:
            n = len(g)
            json.dump({"gather_kernel_B256": {
                "dram_read_bytes": round(sum(x[0] for x in g) / n), "dram_write_bytes": round(sum(x[1] for x in g) / n),
                "launches": n, "ncu_us_per_launch": round(sum(x[2] for x in g) / n, 3),
                "note": "mean over the gather launches of one `ncu --set full --graph-profiling node --cache-control none "
                        "--clock-control none` capture of profiles/prof_bench_gather.py (8-leaf 1M-row Atari storage, "
                        "B=256): algorithmic READ bytes 14.46 MB; the writes are still in L2 (write-back) when the "
                        "kernel ends"}}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
