"""profiles/ncu_traffic.json from an `ncu --set full` report of profiles/run_forward.py: the LAST
forward in the report (launch order: mel, conv_in, then qkv / attention / out-projection / FFN per
layer, conv_out) -> DRAM bytes, tensor-pipe activity and duration per engine kernel label.

    python profiles/make_ncu_traffic.py gpurun_out/r02_full_final2.ncu-rep profiles/r02_ncu_full_batch64.md
"""
import csv
import io
import json
import os
import subprocess
import sys

LAYERS = 5
ORDER = (['mel_stft_fbank_rows', 'tc_conv_in'] +
         ['tc_qkv', 'tc_attention', 'tc_out_proj_ln', 'tc_ffn_fused_ln'] * LAYERS + ['tc_conv_out_softmax'])


def main(report, source):
    raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, data = rows[0], rows[2:][-len(ORDER):]
    col = {name: header.index(name) for name in (
        'Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}
    units = dict(zip(header, rows[1]))
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    time_scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
    assert 'mel_kernel' in data[0][col['Kernel Name']], data[0][col['Kernel Name']]
    kernels = {}
    for label, row in zip(ORDER, data):
        bytes_ = sum(float(row[col[m]]) * scale[units[m]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        entry = kernels.setdefault(label, {'dram_bytes': 0.0, 'tensor_pipe_active_pct': 0.0,
                                           'launches_per_forward': 0, 'ncu_duration_us': 0.0})
        entry['dram_bytes'] += bytes_
        entry['tensor_pipe_active_pct'] += float(row[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])
        entry['ncu_duration_us'] += float(row[col['gpu__time_duration.sum']]) * time_scale[units['gpu__time_duration.sum']]
        entry['launches_per_forward'] += 1
    total = sum(k['dram_bytes'] for k in kernels.values())
    for entry in kernels.values():
        n = entry['launches_per_forward']
        entry.update(dram_bytes=entry['dram_bytes'] / n, tensor_pipe_active_pct=round(entry['tensor_pipe_active_pct'] / n, 2),
                     ncu_duration_us=round(entry['ncu_duration_us'] / n, 1), source=source)
    out = {'source': f'{source} (ncu --set full --clock-control none, one forward at 64 x 10 s: the last '
                     f'{len(ORDER)} launches of profiles/run_forward.py)',
           'dram_bytes_per_forward': total, 'kernels': kernels}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_traffic.json')
    json.dump(out, open(path, 'w'), indent=1)
    print(json.dumps({k: (round(v['dram_bytes'] / 1e6, 1), v['ncu_duration_us']) for k, v in kernels.items()}), total / 1e9)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
