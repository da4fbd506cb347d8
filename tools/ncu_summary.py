"""Condense an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into the table kept under profiles/
and into profiles/roofline_traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py).

    ncu -i gpurun_out/r01_fourier_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r01_fourier_ncu_raw.csv [--traffic tcgen05_fourier]

The Fourier path's roofline object aggregates the five launches of PartI layers 2+3 per fragment (forward transform, layer-2
per-irrep GEMMs, inverse/activation/forward transform, layer-3 GEMMs, inverse/shortcut/activation transform), so its
`traffic` is the mean DRAM bytes per launch over those launches of one captured step.
"""
import csv
import json
import os
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in body:
            w.writerow([r[c] for c in cols])
    if "--traffic" in sys.argv:
        impl = sys.argv[sys.argv.index("--traffic") + 1]
        kn, rd, wr, gs = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size"))
        sel = []
        for r in body:
            name = r[kn]
            tot = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
            # layers 2+3 of PartI: the transform launches and the grouped per-irrep GEMM launches (the longest tensor-core launches)
            sel.append((name, tot, float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")), r))
        xf = [s for s in sel if "group_transform" in s[0]]
        gm = sorted([s for s in sel if "gconv_tc_kernel" in s[0]], key=lambda s: -s[2])[: 2 * len(xf) // 3]
        use = xf + gm
        out_path = os.path.join(os.path.dirname(os.path.abspath(dst)), "roofline_traffic.json")
        try:
            doc = json.load(open(out_path))
        except Exception:
            doc = {}
        doc[impl] = {
            "kernel": "PartI layers 2+3 in the group-Fourier domain: %d transform launches + %d grouped per-irrep GEMM launches of one step" % (len(xf), len(gm)),
            "dram_bytes_per_launch": sum(s[1] for s in use) / max(1, len(use)),
            "dram_bytes_transform_launches": [s[1] for s in xf],
            "dram_bytes_gemm_launches": [s[1] for s in gm],
            "source": os.path.basename(dst),
        }
        json.dump(doc, open(out_path, "w"), indent=1)
        print(json.dumps(doc[impl], indent=1))


if __name__ == "__main__":
    main()
