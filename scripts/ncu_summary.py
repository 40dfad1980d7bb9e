"""Turns the ncu captures of scripts/gpu_profile.sh (gpurun_out/) into the committed summaries under profiles/:
  python scripts/ncu_summary.py TAG
    gpurun_out/prof_train_TAG.ncu-rep, prof_render_TAG.ncu-rep -> profiles/TAG_ncu_full_summary.json
    gpurun_out/launches_TAG.csv                                -> profiles/TAG_launches_bench_graph.{csv,md}
Reads the reports with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum']


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    res = []
    for r in body:
        d = OrderedDict()
        d['Kernel Name'] = r[head.index('Kernel Name')]
        for k in KEEP:
            if k in head:
                i = head.index(k)
                d[k] = f'{r[i]} {units[i]}'.strip()
        res.append(d)
    return res


def launches_md(tag, src_csv):
    rows = []
    with open(src_csv) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(io.StringIO(''.join(lines)))
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1000.0 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1000.0)
        rows.append((r['Kernel Name'], us))
    # one steady-state step = between the last two adam_kernel launches
    adam = [i for i, (k, _) in enumerate(rows) if 'adam_kernel' in k]
    if len(adam) < 2:
        return None
    step = rows[adam[-2] + 1:adam[-1] + 1]
    agg = OrderedDict()
    for k, us in step:
        k = k.split('(')[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(us for _, us in step)
    ours = sum(v[1] for k, v in agg.items() if k.startswith('hn::') or 'hn::' in k)
    md = [f'# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-render --no-skip-leg` ({tag})', '',
          f'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 3500 --csv` (full CSV: `{tag}_launches_bench_graph.csv`).',
          'Per-launch times under ncu are cold-cache and serialised; what must agree with the bench is each kernel\'s SHARE of the step.', '',
          f'One steady-state train step (between two `adam_kernel` launches, replayed from the CUDA graph): **{len(step)} launches, '
          f'{total / 1000:.2f} ms** summed, of which **{ours / 1000:.2f} ms ({100 * ours / total:.0f} %)** in this repo\'s kernels (`hn::*`).', '',
          '| kernel | launches / step | us / step | share |', '|---|---:|---:|---:|']
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        md.append(f'| `{k[:100]}` | {n} | {us:.1f} | {100 * us / total:.1f} % |')
    return '\n'.join(md) + '\n'


def main():
    tag = sys.argv[1]
    go = os.path.join(ROOT, 'gpurun_out')
    summary = OrderedDict()
    for name, title in (('prof_train_%s.ncu-rep', 'train step kernels, bf16 path (ncu --set full --clock-control none --import-source on, '
                                                 'scripts/profile_render.py --rays 6144 --train)'),
                        ('prof_render_%s.ncu-rep', 'render kernels (ncu --set full --clock-control none --import-source on, '
                                                  'scripts/profile_render.py --rays 32768)')):
        rep = os.path.join(go, name % tag)
        if os.path.exists(rep):
            summary[title] = raw_rows(rep)
    with open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_summary.json'), 'w') as f:
        json.dump(summary, f, indent=1)
    src = os.path.join(go, f'launches_{tag}.csv')
    if os.path.exists(src):
        shutil.copy(src, os.path.join(ROOT, 'profiles', f'{tag}_launches_bench_graph.csv'))
        md = launches_md(tag, src)
        if md:
            with open(os.path.join(ROOT, 'profiles', f'{tag}_launches_bench_graph.md'), 'w') as f:
                f.write(md)
            print(md)
    for title, rows in summary.items():
        print(title)
        for r in rows:
            print('  ', r['Kernel Name'][:60], r.get('gpu__time_duration.sum'), 'tensor', r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
                  'dram r/w', r.get('dram__bytes_read.sum'), r.get('dram__bytes_write.sum'))


if __name__ == '__main__':
    main()
