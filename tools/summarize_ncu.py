"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_tc_r16.md "title"
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram %peak'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %peak'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor inst'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem'),
    ('launch__occupancy_limit_shared_mem', 'occ lim smem'),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else 'ncu summary')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_col = idx.get('Kernel Name', 4)
    seen_labels = []
    cols = []
    for k, label in KEYS:
        if k in idx and label not in seen_labels:
            cols.append((idx[k], label, units[idx[k]]))
            seen_labels.append(label)
    lines = [f'# {title}', '', f'source: `{rep}` (ncu --set full --clock-control none; per-launch, cold-cache, serialised)', '',
             '| kernel | ' + ' | '.join(f'{l} [{u}]' if u else l for _, l, u in cols) + ' |',
             '|---|' + '---|' * len(cols)]
    for r in rows[2:]:
        if len(r) <= name_col:
            continue
        name = r[name_col].split('(')[0].replace('moka::', '')
        lines.append('| ' + name + ' | ' + ' | '.join(r[i] for i, _, _ in cols) + ' |')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
