#!/usr/bin/env python
"""Turn an `ncu --set full` report (+ the per-launch duration list) into the committed
summaries under profiles/:

  python profiles/summarize.py gpurun_out/prof_r1.ncu-rep gpurun_out/launches_r1.csv r1 16384

writes profiles/<tag>_ncu_summary.md, profiles/<tag>_launches.csv (kernel launches of the
bench step with their device time) and profiles/traffic.json (dram bytes per launch, read by
bench.py for roofline.traffic).
"""
from __future__ import annotations

import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
try:
  COMMIT = subprocess.run(['git', '-C', HERE, 'rev-parse', '--short', 'HEAD'], capture_output=True,
                          text=True).stdout.strip()
except OSError:
  COMMIT = None
KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'CTAs/SM (register limit)'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
     'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
     'stall not_selected / issue'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'L1 global load sectors'),
    ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'L1 global load requests'),
]
ALGO = {'fwd_pre': 29, 'fwd_post': 12, 'bwd_post': 16, 'post_vg': 16, 'bwd_pre': 37}


def kernel_key(name):
  """fwd_pre | bwd_pre | fwd_post | bwd_post | post_vg from a demangled kernel name
  (post_kernel<PARAM, HAVEW, MODE, CRAW, REDUCE, NT>: MODE 0 value, 1 gradient, 2 both)."""
  if 'fwd_pre' in name:
    return 'fwd_pre'
  if 'bwd_pre' in name:
    return 'bwd_pre'
  m = re.search(r'post_kernel<\s*\(?[^,]*,\s*\(?[^,]*,\s*\(?(?:int\))?(\d)', name)
  mode = int(m.group(1)) if m else 2
  return ('fwd_post', 'bwd_post', 'post_vg')[mode]


def ncu_csv(rep, page, extra=()):
  out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra],
                       capture_output=True, text=True).stdout
  return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
  mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  return float(v) * mult.get(unit, 1)


def main():
  rep, launches, tag, rows = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
  raw = ncu_csv(rep, 'raw')
  hdr, units = raw[0], raw[1]
  col = {h: i for i, h in enumerate(hdr)}
  nsub = rows * 3072
  md = [f'# ncu summary `{tag}` ({os.path.basename(rep)}; bench.py --rows {rows}, eps model)\n',
        'One `ncu --set full --clock-control none --import-source on` capture per kernel of the '
        'bench step (cold-cache, serialised: compare shares, not absolutes). Algorithmic bytes = '
        'bytes/sub-pixel x rows x 3072.\n']
  traffic = {}
  for r in raw[2:]:
    name = r[col['Kernel Name']]
    short = re.sub(r'\(.*', '', name).replace('void ', '').strip()
    md.append(f'\n## `{short}`\n\n| metric | value |\n|---|---|')
    vals = {}
    for k, label in KEYS:
      if k in col:
        vals[k] = (r[col[k]], units[col[k]])
        md.append(f'| {label} (`{k}`) | {r[col[k]]} {units[col[k]]} |')
    rd = to_bytes(*vals['dram__bytes_read.sum'])
    wr = to_bytes(*vals['dram__bytes_write.sum'])
    algo = ALGO.get(kernel_key(name))
    inst = float(vals['smsp__inst_executed.sum'][0])
    md.append(f'| dram traffic per launch | {(rd + wr) / 1e9:.4f} GB |')
    if algo:
      md.append(f'| algorithmic bytes per launch | {algo * nsub / 1e9:.4f} GB ({algo} B/sub-pixel) |')
      md.append(f'| traffic / algorithmic | {(rd + wr) / (algo * nsub):.3f} |')
    md.append(f'| thread instructions per sub-pixel | {inst * 32 / nsub:.1f} |')
    key = kernel_key(name)
    traffic[key] = {'rows': rows, 'dram_bytes_per_launch': rd + wr, 'kernel': short,
                    'capture': os.path.basename(rep), 'commit': COMMIT}
    # instruction mix from the source page
    src = ncu_csv(rep, 'source', ['--kernel-name', 'regex:' + re.escape(short.split('<')[0].split('::')[-1])])
    try:
      h2 = src[1]
      i_s, i_e = h2.index('Source'), h2.index('Instructions Executed')
      cnt = collections.Counter()
      for row in src[2:]:
        if len(row) <= i_e:
          continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', row[i_s].strip())
        if not m:
          continue
        op = m.group(2)
        op = op if op.startswith('MUFU') else op.split('.')[0]
        cnt[op] += int(row[i_e])
      tot = sum(cnt.values())
      if abs(tot - inst) / inst < 0.05:
        top = ', '.join(f'{op} {n * 32 / nsub:.1f}' for op, n in cnt.most_common(14))
        md.append(f'| instruction mix (thread instr / sub-pixel) | {top} |')
    except (ValueError, IndexError):
      pass
  with open(os.path.join(HERE, f'{tag}_ncu_summary.md'), 'w') as f:
    f.write('\n'.join(md) + '\n')
  with open(os.path.join(HERE, 'traffic.json'), 'w') as f:
    json.dump(traffic, f, indent=1)
  # launch list: keep our kernels' rows only
  keep = []
  with open(launches) as f:
    for ln in f:
      if ln.startswith('"ID"') or 'mulan::' in ln or 'Kernel Name' in ln:
        keep.append(ln)
  with open(os.path.join(HERE, f'{tag}_launches.csv'), 'w') as f:
    f.writelines(keep)
  print('\n'.join(md))


if __name__ == '__main__':
  main()
