#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel JSON bench.py reads (`traffic`) and the judge can diff.

usage: python profiles/extract_ncu_full.py report.ncu-rep > profiles/r02_ncu_full_kernels.json

Every captured launch is one entry; entries of the streaming step's own six launches also carry "kernel" = the name bench.py uses
for that launch (its `roofline_kernels` keys), so that bench.py can fill `traffic` (DRAM bytes per launch) from the FIRST such entry.
"""
import csv
import json
import subprocess
import sys

STAGE = [("dft_filter_bulk_kernel", "dft_filter"), ("dft_filter_kernel", "dft_filter(generic)"), ("gather_ab_kernel", "gather_ab"), ("nbr_aggregate_kernel", "nbr_lookup_aggregate"),
         ("pe_mlp_cluster_kernel<(int)8>", "pe_mlp(nbr)"), ("pe_mlp_cluster_kernel<(int)4>", "pe_mlp(update A)"),
         ("pe_mlp_cluster_kernel<(int)12>", "pe_mlp(update B)"), ("pe_mlp_cluster_kernel<8>", "pe_mlp(nbr)"),
         ("pe_mlp_cluster_kernel<4>", "pe_mlp(update A)"), ("pe_mlp_cluster_kernel<12>", "pe_mlp(update B)"),
         ("phaseB_push_kernel", "phaseB_push"),
         ("edge_aggregate_kernel", "edge_aggregate"), ("ring_append_kernel", "ring_append"), ("sample_recent_kernel", "sample_recent")]
# kernel of the Reddit-shaped step -> bench.py's KERNELS name (the paired MLP runs 32-row tiles there, the phase-B MLP 40-row tiles)
BENCH = [("dft_filter_bulk_kernel", "dft_filter"), ("gather_ab_kernel", "gather_ab (a6 lookup+aggregate || a7 edge aggregate)"),
         ("pe_mlp_cluster_kernel<8>", "pe_mlp pair (neighbourhood MLP || phase-A MLP)"), ("phaseB_push_kernel", "phaseB_push"),
         ("pe_mlp_cluster_kernel<10>", "pe_mlp (phase B)"), ("ring_append_kernel", "ring_append")]
KEYS = {
    "gpu__time_duration.sum": "dur",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
    "launch__cluster_size": "cluster", "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_bytes.sum": "l2_bytes", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "smsp__inst_executed.sum": "warp_insts",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(ln for ln in out.splitlines() if ln.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    res = []
    for r in data:
        e = {"name": r[ki].split("(")[0].replace("void ", "").replace("lstep::", "").replace("<unnamed>::", "")[:60]}
        full = r[ki]
        for pat, st in STAGE:
            if pat in full:
                e["stage"] = st
                break
        for pat, kn in BENCH:
            if pat in full.replace("(int)", ""):
                e["kernel"] = kn
                break
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    e[KEYS[h]] = v
                    continue
                e[KEYS[h]] = x * UNIT.get(u, 1.0) if u in UNIT else x
        if "dram_read" in e and "dram_write" in e:
            e["dram_bytes_per_launch"] = e["dram_read"] + e["dram_write"]
        if "dur" in e:
            e["dur_us"] = e.pop("dur")
        res.append(e)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
