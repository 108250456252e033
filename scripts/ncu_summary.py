"""Summarise an `ncu --page raw --csv` dump: one block of key counters per profiled launch."""
import csv
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum',
        'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum', 'sm__icc_requests.sum', 'sm__icc_request_hit_rate.pct',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=====")
        for k in KEYS:
            if k in d:
                print(f"{k:70s} {d[k]} {units[hdr.index(k)]}")
        stalls = [(float(d[k].replace(',', '')), k) for k in hdr
                  if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and d[k]]
        for v, k in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.2f} warps/issue")


if __name__ == "__main__":
    main(sys.argv[1])
