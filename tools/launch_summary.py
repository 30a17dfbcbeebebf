"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table of ONE train step.
    python tools/launch_summary.py launches.csv out.md [title]
A step is delimited by consecutive `k_march_warp<0>` launches (the first kernel of render_train after near_far); the
LAST complete step in the list that contains no occupancy update is summarised."""
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else 'ncu launch list of one train step'
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1e3 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1e3
        rows.append((r['Kernel Name'], us))
    marks = [i for i, (n, _) in enumerate(rows) if 'k_march_warp<0>' in n or 'k_march_warp<(bool)0>' in n]
    steps = [(a, b) for a, b in zip(marks[:-1], marks[1:])]
    steps = [s for s in steps if not any('k_packbits' in rows[i][0] for i in range(*s))] or steps
    a, b = steps[-1]
    agg = {}
    for n, us in rows[a:b]:
        n = re.sub(r'\(.*$', '', n) if n.startswith('void at::') or 'at::' in n else n
        n = n[:110]
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += us
    total = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# %s\n\n' % title)
        f.write('(per-launch times are cold-cache and serialised under ncu: compare SHARES.)  One step (no occupancy update) = '
                '%d launches, %.2f ms of kernel time.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n' % (b - a, total / 1e3))
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.1f | %.1f%% |\n' % (n, c, us, 100 * us / total))
    print('step of %d launches, %.2f ms' % (b - a, total / 1e3))


if __name__ == '__main__':
    main()
