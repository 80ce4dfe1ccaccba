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


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] != '--traffic':
    main(sys.argv[1], sys.argv[2])


def traffic(kernels_csv, out_json, note):
    """Per-stage DRAM traffic table for bench.py's roofline.traffic, from a *_kernels.csv of ONE forward (19 launches in
    order: stem, conv1..13, decode_conv1..5).  usage: python profiles/extract_ncu.py --traffic profiles/X_kernels.csv out.json"""
    import json
    rows = list(csv.reader(open(kernels_csv)))
    hdr = rows[0]
    names = ['conv%d' % i for i in range(14)] + ['decode_conv%d' % j for j in range(1, 6)]
    body = rows[2:]
    assert len(body) == len(names), (len(body), len(names))
    g = lambda r, k: float(r[hdr.index(k)])
    unit = dict(zip(hdr, rows[1]))
    scale = {'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6, 'Gbyte': 1e3}
    stages = {}
    for n, r in zip(names, body):
        stages[n] = {
            'kernel': r[hdr.index('Kernel Name')][:60],
            'dram_read_mb': g(r, 'dram__bytes_read.sum') * scale[unit['dram__bytes_read.sum']],
            'dram_write_mb': g(r, 'dram__bytes_write.sum') * scale[unit['dram__bytes_write.sum']],
            'ncu_time_us': g(r, 'gpu__time_duration.sum'),
            'issue_active_pct': g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            'tensor_active_pct': g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        }
    json.dump({'source': note, 'stages': stages}, open(out_json, 'w'), indent=1)
    print('wrote', out_json)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == '--traffic':
    traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
