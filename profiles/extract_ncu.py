"""Turn an .ncu-rep brought back in gpurun_out/ into the small text summaries committed here.
usage: python profiles/extract_ncu.py gpurun_out/prof_XXX.ncu-rep profiles/NAME"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu.sum']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(out + '_kernels.csv', 'w') as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:110] for _, i in idx])
    print('wrote', out + '_kernels.csv', len(rows) - 2, 'kernels')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
