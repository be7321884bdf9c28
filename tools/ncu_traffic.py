"""Summarise an `ncu --csv` metric log of the persistent loop kernels: per kernel launch DRAM bytes, duration, L2 traffic, tensor-pipe
share -> a table (stdout / --table) and profiles/r2/ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_traffic.py gpurun_out/ncu_loops_app_replay.csv --B 60 --L 180 --T 900 --json profiles/r2/ncu_traffic.json --table profiles/r2/ncu_loops_metrics.txt
"""
import argparse
import collections
import csv
import json
import re

SHORT = [('lstm_loop_tc_kernel<1', 'lstm_loop_tc_kernel<att>'), ('lstm_loop_tc_kernel<0', 'lstm_loop_tc_kernel<gen>'),
         ('lstm_loop_tc_kernel<true', 'lstm_loop_tc_kernel<att>'), ('lstm_loop_tc_kernel<false', 'lstm_loop_tc_kernel<gen>'),
         ('lstm_loop_tc_kernel<true>', 'lstm_loop_tc_kernel<att>'), ('lstm_loop_tc_kernel<false>', 'lstm_loop_tc_kernel<gen>'),
         ('att_bwd_loop_kernel', 'att_bwd_loop_kernel'), ('lstm_bwd_loop_tc_kernel', 'lstm_bwd_loop_tc_kernel'),
         ('lstm_bwd_loop_kernel', 'lstm_bwd_loop_kernel'), ('att_post_kernel', 'att_post_kernel')]


def short(name):
    for pat, s in SHORT:
        if pat in name:
            return s
    return re.sub(r'\(.*', '', name)[-60:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--B', type=int, required=True); ap.add_argument('--L', type=int, required=True); ap.add_argument('--T', type=int, required=True)
    ap.add_argument('--json', default=''); ap.add_argument('--table', default=''); ap.add_argument('--note', default='')
    a = ap.parse_args()
    rows = collections.OrderedDict()
    with open(a.csv) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        key = (r['ID'], short(r['Kernel Name']))
        rows.setdefault(key, {})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    # first launch of every kernel name (later launches of the same command repeat it)
    first = collections.OrderedDict()
    for (kid, name), m in rows.items():
        first.setdefault(name, m)
    out = {'B': a.B, 'L': a.L, 'T': a.T, 'note': a.note, 'kernels': {}}
    lines = [f'{"kernel":28s} {"ms":>8s} {"DRAM rd GB":>11s} {"DRAM wr GB":>11s} {"DRAM %pk":>9s} {"L2 GB":>9s} {"tensor inst":>12s} {"tensor %act":>11s} {"warps %":>8s} {"SM thr %":>9s}']
    for name, m in first.items():
        g = lambda k: m.get(k, float('nan'))   # noqa: E731
        out['kernels'][name] = {'dram_read': g('dram__bytes_read.sum'), 'dram_write': g('dram__bytes_write.sum'),
                                'duration_ms': g('gpu__time_duration.sum') / 1e6, 'l2_bytes': g('lts__t_bytes.sum'),
                                'tensor_inst': g('sm__inst_executed_pipe_tensor.sum'),
                                'tensor_active_pct': g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                                'dram_pct_of_peak': g('dram__throughput.avg.pct_of_peak_sustained_elapsed')}
        lines.append(f'{name:28s} {g("gpu__time_duration.sum") / 1e6:8.2f} {g("dram__bytes_read.sum") / 1e9:11.3f} {g("dram__bytes_write.sum") / 1e9:11.3f} '
                     f'{g("dram__throughput.avg.pct_of_peak_sustained_elapsed"):9.2f} {g("lts__t_bytes.sum") / 1e9:9.2f} {g("sm__inst_executed_pipe_tensor.sum"):12.3e} '
                     f'{g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):11.2f} {g("sm__warps_active.avg.pct_of_peak_sustained_active"):8.2f} '
                     f'{g("sm__throughput.avg.pct_of_peak_sustained_elapsed"):9.2f}')
    text = '\n'.join(lines)
    print(text)
    if a.table:
        with open(a.table, 'w') as f:
            f.write((a.note + '\n' if a.note else '') + text + '\n')
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
