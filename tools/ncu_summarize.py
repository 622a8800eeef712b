#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summarize.py launches gpurun_out/launches.csv           -> markdown table on stdout (per-kernel shares)
    python tools/ncu_summarize.py metrics  gpurun_out/prof.ncu-rep           -> markdown table of the roofline metrics
    python tools/ncu_summarize.py stalls   gpurun_out/prof.ncu-rep [N]       -> top-N SASS lines by warp-stall samples
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum',
           'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem']


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    if name.startswith('at::') or 'at::native' in name:
        m = re.search(r'(\w+Functor\w*|\w+Ops\w*|launch_\w+|\w+Impl)', name)
        return 'torch: ' + (m.group(1) if m else name[:40])
    return name


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[row['Metric Unit']]
        rows.append((short(row['Kernel Name']), v))
    tot = sum(v for _, v in rows)
    agg = collections.OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    print(f'launches: {len(rows)}, total device time {tot / 1e3:.2f} ms\n')
    print('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|')
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% | {v / c:.1f} |')
    big = [i for i, (n, v) in enumerate(rows) if 'sweep_sm100' in n and v > 300]
    if len(big) >= 3:
        step = rows[big[-2] - (big[-1] - big[-2] - (len(rows) - big[-1])):]     # the last full step (two passes)
        st = sum(v for _, v in step)
        sw = sum(v for n, v in step if 'sweep_sm100' in n and v > 300)
        print(f'\nLast timed step (2 head passes): {len(step)} launches, {st / 1e3:.3f} ms of device time, main sweep share {100 * sw / st:.1f}%')


def raw(path, page):
    return subprocess.run(['ncu', '-i', path, '--page', page, '--csv'], capture_output=True, text=True).stdout


def metrics(path):
    r = list(csv.reader(io.StringIO(raw(path, 'raw'))))
    hdr, units, vals = r[0], r[1], r[2]
    print('| metric | unit | value |\n|---|---|---|')
    for want in ['Kernel Name'] + METRICS:
        if want in hdr:
            i = hdr.index(want)
            print(f'| {want} | {units[i]} | {vals[i]} |')


def stalls(path, top=24):
    r = list(csv.reader(io.StringIO(raw(path, 'source'))))
    while r and 'Source' not in r[0]:      # first line names the kernel
        r = r[1:]
    hdr = r[0]
    col = {h: i for i, h in enumerate(hdr)}
    src = col.get('Source')
    samp = col.get('Warp Stall Sampling (All Samples)', col.get('# Samples'))
    ex = col.get('Instructions Executed')
    stall_cols = [(h[len('stall_'):], i) for h, i in col.items() if h.startswith('stall_') and 'Not Issued' not in h]
    rows = []
    for row in r[1:]:
        try:
            n = int(row[samp])
        except Exception:
            continue
        reasons = sorted(((int(row[i]) if row[i].isdigit() else 0, nm) for nm, i in stall_cols), reverse=True)[:2]
        rows.append((n, row[src].strip(), row[ex] if ex is not None else '', reasons))
    tot = sum(n for n, *_ in rows) or 1
    print(f'total samples {tot}')
    for n, s, e, reasons in sorted(rows, key=lambda t: -t[0])[:top]:
        print(f'{n:7d} {100 * n / tot:5.1f}%  {s[:58]:58s} exec={e:>10s} {[(c, nm) for c, nm in reasons if c]}')


if __name__ == '__main__':
    cmd, path = sys.argv[1], sys.argv[2]
    {'launches': launches, 'metrics': metrics, 'stalls': lambda p: stalls(p, int(sys.argv[3]) if len(sys.argv) > 3 else 24)}[cmd](path)
