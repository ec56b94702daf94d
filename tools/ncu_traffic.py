"""Turn an `ncu --set full` capture of the dominant kernel's launches of ONE pass (tools/profile_pass.py) into
  * profiles/<tag>.csv — one row per captured launch with the metrics quoted anywhere in this repository, and
  * the entry `<workload>:<stage>:<variant>` of profiles/ncu_traffic.json that bench.py reads for `roofline.traffic`,
    `roofline.dram_frac`, `roofline.l2_frac` and the `roofline.ncu` counters.
Read here (no GPU needed): python tools/ncu_traffic.py gpurun_out/r2_trace_full_rungholt.ncu-rep rungholt trace wavefront r2_ncu_trace_rungholt
A whole-pass capture exported as csv on the box:   python tools/ncu_traffic.py gpurun_out/r2_pass_full_cornell.raw.csv cornell trace wavefront r2_ncu_trace_cornell wfTrace
Every number bench.py derives from the entry can be recomputed from the csv: per-launch sums / duration-weighted means."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main(rep, workload, stage, variant, tag, kernel=None):
    """rep: a .ncu-rep, or the csv of its raw page exported on the GPU box (ncu -i x.ncu-rep --page raw --csv) when the report is too
    large to bring back; kernel: regular expression selecting the launches of the stage out of a whole-pass capture; stage "-" writes
    the csv only (a whole-pass table, no ncu_traffic.json entry)"""
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if kernel:
        import re
        data = [r for r in data if re.search(kernel, r[hdr.index("Kernel Name")])]
    cols = [h for h in KEEP if h in hdr]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for h in cols:
            i = hdr.index(h)
            try:
                d[h] = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)      # bytes and milliseconds
            except ValueError:
                d[h] = None
        launches.append(d)
    out_csv = os.path.join(ROOT, "profiles", tag + ".csv")
    with open(out_csv, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["kernel"] + cols)
        wr.writerow(["(units: bytes, ms, sectors, %)"] + [units[hdr.index(h)] for h in cols])
        for d in launches:
            wr.writerow([d["kernel"]] + [d[h] for h in cols])
    if stage == "-":
        print("wrote", out_csv, len(launches), "launches")
        return
    n = len(launches)
    ms = [d["gpu__time_duration.sum"] for d in launches]
    tot_ms = sum(ms)

    def wmean(h):
        return sum(d[h] * t for d, t in zip(launches, ms)) / tot_ms

    def total(h):
        return sum(d[h] for d in launches)
    dram = total("dram__bytes_read.sum") + total("dram__bytes_write.sum")
    entry = {
        "launches": n, "ncu_ms_per_step": tot_ms,
        "dram_bytes_per_launch": dram / n, "dram_bytes_read_per_step": total("dram__bytes_read.sum"), "dram_bytes_write_per_step": total("dram__bytes_write.sum"),
        "l2_bytes_per_step": 32.0 * total("lts__t_sectors.sum"), "l2_to_l1_read_bytes_per_step": total("l1tex__m_xbar2l1tex_read_bytes.sum"),
        "l1_load_sectors_per_step": total("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
        "l1_load_requests_per_step": total("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
        "source": f"profiles/{tag}.csv: ncu --set full --clock-control none, the {n} launches of the kernel in one steady-state pass (tools/profile_pass.py, "
                  "tools/ncu_traffic.py); byte counts summed over the launches, percentages duration-weighted",
        "ncu": {
            "issue_slot_utilisation_pct": wmean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "active_lanes_per_instruction": wmean("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "l1_sector_hit_rate_pct": wmean("l1tex__t_sector_hit_rate.pct"), "l2_sector_hit_rate_pct": wmean("lts__t_sector_hit_rate.pct"),
            "dram_throughput_pct_of_peak": wmean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_throughput_pct_of_peak": wmean("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1_data_stage_wavefronts_pct_of_peak": wmean("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": wmean("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": launches[0].get("launch__registers_per_thread"),
            "weighting": f"duration-weighted over the {n} launches of the pass",
        },
    }
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allv = json.load(open(path)) if os.path.exists(path) else {}
    allv[f"{workload}:{stage}:{variant}"] = entry
    json.dump(allv, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:7])
