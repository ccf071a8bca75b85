#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (one row per metric per launch) for ONE bench step:

    python scripts/launch_summary.py gpurun_out/rNN_launches_batch512.csv > profiles/rNN_launch_summary_batch512.txt

Finds one whole step (from the first space-to-depth / pad kernel of an encoder pass to the next pass), aggregates
duration, DRAM bytes, L2->SM read sectors and tensor-pipe activity per kernel, lists the encoder's GEMM launches, and
prints the average DRAM bytes per conv launch (bench.py's `roofline.traffic`, profiles/traffic.json)."""
import collections
import csv
import json
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[start]
    ik, iv, im, iid = (h.index(n) for n in ('Kernel Name', 'Metric Value', 'Metric Name', 'ID'))
    L, last = [], None
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        if r[iid] != last:
            L.append({'name': r[ik]})
            last = r[iid]
        try:
            L[-1][r[im]] = float(r[iv].replace(',', ''))
        except ValueError:
            pass
    first = [i for i, d in enumerate(L) if ('s2d_split' in d['name'] or 'pad_c3' in d['name'])]
    starts = [i for k, i in enumerate(first) if k == 0 or i - first[k - 1] > 50]
    s0, s1 = starts[0], starts[1]
    step = L[s0:s1]
    dur = lambda d: d.get('gpu__time_duration.sum', 0.0) / 1e3
    dram = lambda d: (d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)) / 1e6
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for d in step:
        n = d['name'].split('(')[0][-52:]
        a = agg[n]
        a[0] += 1; a[1] += dur(d); a[2] += dram(d)
        a[3] += d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0) * dur(d)
        a[4] += d.get('lts__t_sectors_srcunit_tex_op_read.sum', 0.0) * 32 / 1e6
    tot = sum(a[1] for a in agg.values())
    print('# one bench step = launches %d..%d of %s (ncu: serialised, cold caches -> compare shares)' % (s0, s1 - 1, path))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-54s n=%4d %9.1f us %5.1f%%  dram %8.1f MB  L2->SM %9.1f MB (%4.2f TB/s)  tensor-active %4.1f%%' % (
            n, a[0], a[1], 100 * a[1] / tot, a[2], a[4], a[4] / max(a[1], 1e-9) / 1e6 * 1e6 / 1e6, a[3] / max(a[1], 1e-9)))
    print('total %.1f us in %d launches' % (tot, len(step)))
    print('# encoder GEMM launches of the step')
    enc = [d for d in step if 'gemm_bf16x3' in d['name']]
    n_dec = sum(1 for d in step if 'attn_fused' in d['name'])
    enc = enc[:len(enc) - 2 * n_dec - 1] if n_dec else enc          # drop project + per-decode-step gates / [logits|q]
    for k, d in enumerate(enc):
        print('%3d %-34s %8.1f us  tensor %4.1f%%  dram %6.1f MB  L2->SM %7.1f MB' % (
            k, d['name'].split('(')[0][-34:], dur(d), d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0),
            dram(d), d.get('lts__t_sectors_srcunit_tex_op_read.sum', 0.0) * 32 / 1e6))
    tr = sum(dram(d) for d in enc) * 1e6 / max(len(enc), 1)
    print('# average DRAM bytes per conv launch: %.0f over %d launches' % (tr, len(enc)))
    att = [d for d in step if 'attn_fused' in d['name']]
    ta = sum(dram(d) for d in att) * 1e6 / max(len(att), 1)
    print('# average DRAM bytes per fused-attention launch: %.0f over %d launches' % (ta, len(att)))
    if len(sys.argv) > 2:
        json.dump({'conv': tr, 'scores': ta, 'note': 'dram__bytes_read.sum + dram__bytes_write.sum averaged over the %d conv launches '
                                       'of one 512-image step (%s)' % (len(enc), path)}, open(sys.argv[2], 'w'))


if __name__ == '__main__':
    main()
