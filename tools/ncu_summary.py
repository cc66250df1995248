"""Summarise an .ncu-rep (read here on the CPU box): key throughput, traffic and stall metrics per kernel."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg', 'sm__cycles_active.avg',
        'derived__smsp__sass_thread_inst_executed_op_ffma_pred_on_x2', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'local_load_requests', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')][:70], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for w in WANT:
            if w in hdr:
                print(f'   {w:72s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 3.0:
                    print(f'   stall {h[len("smsp__warp_issue_stalled_"):-len("_per_warp_active.pct")]:40s} {v:8.1f} %')


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
