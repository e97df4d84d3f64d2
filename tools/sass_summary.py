#!/usr/bin/env python
"""Static evidence from the built library (no GPU): per kernel family, registers / stack (spills)
/ shared memory from `cuobjdump -res-usage` and SASS instruction counts from `cuobjdump -sass`
(128-bit global accesses, MUFU, barriers, shuffles, multimem, PDL).
    python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mulan_b200',
                   'libmulan_b200.so')


def demangle(names):
  out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout
  return dict(zip(names, out.splitlines()))


def short(name):
  name = re.sub(r'^void ', '', name)
  name = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', name)        # the parameter list
  return name.replace('mulan::', '').replace('(anonymous namespace)::', '')


res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)', res):
  usage[m.group(1)] = tuple(int(m.group(i)) for i in (2, 3, 4))

sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
counts, cur = {}, None
PAT = collections.OrderedDict([
    ('LDG.128', r'\bLDG\.E(\.\w+)*\.128'), ('STG.128', r'\bSTG\.E(\.\w+)*\.128'),
    ('LDG other', r'\bLDG\.'), ('STG other', r'\bSTG\.'), ('MUFU', r'\bMUFU\.'),
    ('BAR', r'\bBAR\.'), ('SHFL', r'\bSHFL\.'), ('LDGMC', r'\bLDGMC\.'),
    ('local LD/ST', r'\b(LDL|STL)\b')])
for line in sass.splitlines():
  m = re.match(r'\s*Function : (\S+)', line)
  if m:
    cur = m.group(1)
    counts[cur] = collections.Counter()
    continue
  if cur is None or not re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
    continue
  c = counts[cur]
  c['instr'] += 1
  body = line.split('*/', 1)[1]
  hit128 = False
  for k, pat in PAT.items():
    if re.search(pat, body):
      if k in ('LDG other', 'STG other') and hit128:
        continue
      c[k] += 1
      if k in ('LDG.128', 'STG.128'):
        hit128 = True

names = demangle(sorted(usage))
print('# Static kernel summary of `libmulan_b200.so` (sm_100a; `tools/sass_summary.py`, no GPU needed)\n')
print('Registers / stack bytes (0 = no spills) / static shared memory from `cuobjdump -res-usage`;')
print('instruction counts from `cuobjdump -sass`.  One line per kernel instantiation the launchers')
print('can select; template arguments are the ones in the source (`csrc/*.cu`).\n')
rows = []
for mangled, (reg, stack, sh) in usage.items():
  c = counts.get(mangled, collections.Counter())
  rows.append((short(names[mangled]), reg, stack, sh, c))
HOT = [  # what bench.py's default line and its configs block launch at 16384 / 128 rows
    'fwd_pre_kernel<0, true, 2, false, 128, 8>', 'fwd_pre_kernel<0, false, 2, false, 128, 8>',
    'fwd_pre_kernel<0, true, 2, false, 768, 1>',
    'post_kernel<0, true, 2, false, true, 256>', 'post_kernel<1, false, 2, false, true, 256>',
    'post_kernel<0, true, 2, false, true, 768>', 'post_kernel<1, false, 0, false, true, 256>',
    'bwd_pre_kernel<0, 0, false, true, false, 256, 5>', 'bwd_pre_kernel<1, 0, false, true, false, 256, 5>',
    'bwd_pre_kernel<0, 0, false, true, false, 768, 1>', 'adamw_ema_kernel<false>',
    'adamw_ema_peer_kernel<2, 8, false, false>', 'adamw_ema_peer_kernel<4, 2, false, false>',
    'adamw_ema_peer_kernel<8, 4, true, false>']
want = sys.argv[1:] or None
HDR = ('| kernel | regs | stack | smem B | SASS instr | LDG.128 | STG.128 | other LDG / STG | MUFU '
       '| BAR | SHFL | LDGMC | local LD/ST |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|')


def line(name, reg, stack, sh, c):
  return (f"| `{name}` | {reg} | {stack} | {sh} | {c['instr']} | {c['LDG.128']} | {c['STG.128']} | "
          f"{c['LDG other']} / {c['STG other']} | {c['MUFU']} | {c['BAR']} | {c['SHFL']} | "
          f"{c['LDGMC']} | {c['local LD/ST']} |")


by_name = {r[0]: r for r in rows}
print('The instantiations `bench.py` launches (ε / velocity at 16384 rows: 128- and 256-thread CTAs;')
print('128 rows: 768 threads per row; optimizer; fused exchange at 2 / 4 / 8 ranks):\n')
print(HDR)
for h in HOT:
  if h in by_name:
    print(line(*by_name[h]))
print("""
Every per-element global access of the hot kernels is 128 bits wide (`LDG.E.128` / `STG.E.128`;
the remaining narrow ones are x as one `uchar4` per four sub-pixels, the row's t and the per-row
outputs).  The 8 / 16-byte stack frames of `fwd_pre` are one to three accumulators the compiler
parks around the per-pixel slow path: the straight-line fast path pays one `LDL` + one `STL` per
column (ε) or three `STL`s (velocity), i.e. <= 0.75 instructions per sub-pixel of 107.
`LDGMC.E.ADD.F32x4` is `multimem.ld_reduce.add.v4.f32` (the in-switch reduction of
`adamw_ema_peer_kernel<.., true>`); its `multimem.st` is a `STG.E.128.STRONG.SYS` to the
multicast address.  No tensor-core or TMA instructions: the path is elementwise + row reductions
(the opt-in `fwd_pre_tma_kernel` instantiations are the only `UBLKCP` users).

All kernels:
""")
print(HDR)
for name, reg, stack, sh, c in sorted(rows):
  if want and not any(w in name for w in want):
    continue
  print(line(name, reg, stack, sh, c))
spilled = [r[0] for r in rows if r[2] > 0]
print(f'\n{len(rows)} kernels; kernels with a stack frame (spills or local arrays): '
      f"{', '.join('`' + s + '`' for s in spilled) if spilled else 'none'}.")
