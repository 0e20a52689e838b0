"""One table row per kernel launch of an `ncu --page raw --csv` export (profiles/*_ncu_raw_*.csv):
duration, DRAM bytes, GB/s against the measured HBM peak, tensor-pipe %, issue active %, warps
active %, SMs active (sm__cycles_active / sm__cycles_elapsed), L2 throughput %.

    python profiles/ncu_summary.py profiles/r02ab_ncu_raw_pipeline_kernel_interpenetration.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f).get('hbm_gbs', 6553.3))
    except Exception:
        return 6553.3


def to_float(x):
    try:
        return float(str(x).replace(',', ''))
    except ValueError:
        return float('nan')


def scale(value, unit, base):
    """value in `unit` -> bytes or seconds (base 'byte' / 'second')."""
    pre = {'': 1.0, 'K': 1e3, 'M': 1e6, 'G': 1e9, 'T': 1e12, 'k': 1e3, 'm': 1e-3, 'u': 1e-6, 'n': 1e-9}
    u = unit.strip()
    if base == 'byte':
        u = u.replace('byte', '').replace('B', '')
    else:
        u = u.replace('second', '').replace('s', '')
    return value * pre.get(u, 1.0)


def rows_of(path):
    rows = list(csv.reader(open(path, errors='ignore')))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, units = r, rows[i + 1]
            return hdr, units, [x for x in rows[i + 2:] if len(x) == len(hdr)]
    raise SystemExit('no "Kernel Name" header in ' + path)


def main(path):
    hdr, units, rows = rows_of(path)
    col = {h: j for j, h in enumerate(hdr)}

    def get(r, name, base=None):
        j = col.get(name)
        if j is None:
            return float('nan')
        v = to_float(r[j])
        return scale(v, units[j], base) if base else v

    peak = hbm_peak()
    print('| kernel | grid x block | duration | DRAM read + written (MB) | GB/s | frac of HBM | tensor pipe % '
          '| issue active % | warps active % | SMs active (cycles active / elapsed) | L2 throughput % |')
    print('|---|---|---|---|---|---|---|---|---|---|---|')
    for r in rows:
        dur = get(r, 'gpu__time_duration.sum', 'second')
        rd = get(r, 'dram__bytes_read.sum', 'byte')
        wr = get(r, 'dram__bytes_write.sum', 'byte')
        gbs = (rd + wr) / dur / 1e9
        tens = max([get(r, h) for h in hdr if h.startswith('sm__pipe_tensor_cycles_active') and 'pct' in h] or [0.0])
        issue = get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')
        warps = get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')
        act = get(r, 'sm__cycles_active.avg')
        ela = get(r, 'sm__cycles_elapsed.avg')
        l2 = get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
        name = r[col['Kernel Name']].split('(')[0]
        grid = r[col['Grid Size']] if 'Grid Size' in col else '?'
        block = r[col['Block Size']] if 'Block Size' in col else '?'
        print('| `{}` | {} x {} | {:.1f} us | {:.2f} + {:.2f} | {:.3g} | {:.5f} | {:.1f} | {:.1f} | {:.1f} | '
              '{:.3f} | {:.1f} |'.format(name, grid, block, dur * 1e6, rd / 1e6, wr / 1e6, gbs, gbs / peak,
                                         tens, issue, warps, act / ela if ela == ela and ela else float('nan'), l2))


if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
