"""Summarise an .ncu-rep (read here, no GPU): key metrics per captured launch."""
import csv, subprocess, sys
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__maximum_warps_per_active_cycle_pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio', 'smsp__pcsamp_sample_count']

def main(path, out=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        lines.append(f"## {name}")
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                lines.append(f"{h},{u},{v}")
            if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued'):
                try: stalls.append((float(v.replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                except ValueError: pass
        tot = sum(s for s, _ in stalls) or 1
        lines.append("stall_samples," + " ".join(f"{n}={100*s/tot:.1f}%" for s, n in sorted(stalls, reverse=True)[:8]))
    text = "\n".join(lines)
    print(text)
    if out: open(out, 'w').write(text + "\n")

if __name__ == '__main__':
    main(*sys.argv[1:3])
