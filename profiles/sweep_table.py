#!/usr/bin/env python
"""Render a bench.py sweep (one JSON line per point, tools/sweep.sh) as a markdown table.

  python profiles/sweep_table.py gpurun_out/sweep.jsonl r1 > profiles/r1_sweep.md
"""
import json
import sys


def cell(k):
  if k is None:
    return '-'
  return '%.1f (%d, %d%%)' % (k['ms'] * 1e3, round(k['gbs']), round(100 * k['frac_of_measured']))


def main():
  path, tag = sys.argv[1], sys.argv[2]
  rows = [json.loads(l) for l in open(path) if l.strip().startswith('{')]
  peak = rows[0]['roofline']['peak']
  print(f'# Kernel sweep {tag} (1xB200, `tools/sweep.sh`, CUDA events, per-kernel back-to-back launches)\n')
  print(f'GB/s = algorithmic bytes / event time; % = of measured HBM copy peak {peak} GB/s '
        '(MEASURED_PEAKS.json). Train step = fwd_pre + post_vg (fused loss+n_bar) + bpd_reduce + '
        'bwd_pre. `form` = the loss formula the post / bwd_pre kernels run (mulan_kernel_param: '
        'velocity_from_epsilon evaluates the epsilon form unless MULAN_VFE_LITERAL=1). Rows <= 2048 '
        'are (partly) L2-resident and host-launch bound in the per-kernel columns: read the '
        'graph-replayed step time there.\n')
  print('| workload | param | form | note | rows | step ms | Msamples/s | step GB/s (%) | B/sub-pixel | '
        'fwd_pre us (GB/s, %) | post us (GB/s, %) | bwd_pre us (GB/s, %) |')
  print('|---|---|---|---|---|---|---|---|---|---|---|---|')
  for d in rows:
    c, k = d['config'], d['kernels']
    wl = 'dense_vlb' if 'dense' in d['metric'] else 'train'
    post = k.get('post_vg') or k.get('fwd_post')
    bps = d['step_hbm']['algo_bytes_per_step'] / (c['rows_per_gpu'] * c['dim'])
    print('| %s | %s | %s | %s | %d | %.4f | %.2f | %d (%d%%) | %d | %s | %s | %s |' % (
        wl, c['param'], c.get('loss_form', '?'), d.get('note', ''), c['rows_per_gpu'],
        d['ms_per_step'], d['value'] / 1e6, round(d['step_hbm']['gbs']),
        round(100 * d['step_hbm']['frac_of_measured']), round(bps), cell(k.get('fwd_pre')),
        cell(post), cell(k.get('bwd_pre'))))


if __name__ == '__main__':
  main()
