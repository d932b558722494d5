"""Summarise an `ncu --set full` report into a small markdown table.

    python profiles/summarize_ncu.py gpurun_out/r01_prof.ncu-rep > profiles/r01_ncu_full.md

Reads the report with `ncu -i <rep> --page raw --csv` (works on the CPU-only dev
container) and keeps the metrics the roofline discussion needs."""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    name_col = header.index('Kernel Name')
    cols = [(header.index(m), label) for m, label in METRICS if m in header]
    print('| # | kernel | ' + ' | '.join(label for _, label in cols) + ' |')
    print('|---|---|' + '---|' * len(cols))
    for n, row in enumerate(rows[2:]):
        name = row[name_col].split('(')[0].replace('void ', '').replace('ppgs::', '')
        cells = []
        for i, _ in cols:
            value, unit = row[i], units[i]
            try:
                value = f'{float(value.replace(",", "")):.4g}'
            except ValueError:
                pass
            cells.append(f'{value} {unit}'.strip())
        print(f'| {n} | `{name}` | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
