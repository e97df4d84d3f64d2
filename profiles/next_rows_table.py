#!/usr/bin/env python
"""Render tools/bench_next_rows.py output as markdown.

  python profiles/next_rows_table.py gpurun_out/next_rows.jsonl r1 > profiles/r1_next_rows.md
"""
import json
import sys

NOTES = '''
Notes
* `mulan_sample_step`: the first version (the reference's statements op for op: IEEE division, `sqrtf`, `expf`, `expm1f`; kept as `MULAN_SAMPLER_IEEE=1`) measured 318.5 us broadcast / 337.2 us per-example (39 % / 64 %). The default now forms the step as z_s = m1 z_t + m2 net + m3 eps with the factors from e^{gamma/2} and MUFU reciprocal square roots (161.5 us / 209.1 us), and with one coefficient row broadcast over the batch (the unconditional sampler, where all rows also share t) a persistent CTA caches the factors in shared memory while consecutive rows share (t, s): three FMAs per sub-pixel on 16 B of traffic. Rows with individual times under broadcast coefficients (not something the reference's samplers do) recompute the factors per row and are issue-bound.
* `mulan_rk45_stage` / `mulan_rk45_norm`: scalar 4-byte loads measured 529.2 us / 426.2 us (64 % / 79 %); four elements per thread (LDG.128 per stage row) reach the roofline. Grid: 2048 CTAs for the stage kernel (102 % with a resident-only grid), resident-only for the norm kernel (91 % with 2048 CTAs).
* `mulan_adamw_ema`: a resident grid-stride loop measured 437.3 us (89 %); one float4 column per thread over a full grid is what is shown. `mulan_grad_sumsq`: 88 % -> 91 % with 2048 instead of resident-only CTAs.
* `mulan_rng_normal` is instruction-bound by construction (threefry2x32: 20 rounds of add / rotate / xor per two words on the half-rate integer pipe, then erfinv): 4 B written per ~70 instructions; torch's Philox `normal_` is shown for scale. This is why the draws are not fused into `mulan_fwd_pre` (DESIGN.md section 7).
* `mulan_generate_x` runs once per generated batch (after 1000 sampler steps); `mulan_aux_topk_*` work on [B, 50] logits (25 KB at the shipped batch) and are launch-latency bound there: neither was tuned.
'''


def main():
  rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip().startswith('{')]
  tag = sys.argv[2]
  pk = rows[0]['peak_gbs']
  print(f'# "Next" rows (SURVEY.md 8f) and auxiliary kernels, per-kernel roofline {tag} (1xB200)\n')
  print(f'`python tools/bench_next_rows.py`: CUDA events around 20 back-to-back launches, '
        f'{rows[0]["rows"]} rows x 3072 sub-pixels (every operand set is larger than the 126 MB L2). '
        f'GB/s = algorithmic bytes (each declared input read once, each output written once) / time; '
        f'% = of the measured HBM copy peak {pk} GB/s (MEASURED_PEAKS.json).\n')
  print('| kernel | us | algorithmic MB | GB/s | % of measured | bytes counted |')
  print('|---|---|---|---|---|---|')
  for d in rows:
    print('| `%s` | %.1f | %.0f | %d | %.0f%% | %s |' % (
        d['kernel'], d['us'], d['algo_bytes'] / 1e6, round(d['gbs']),
        100 * d['frac_of_measured'], d.get('note', '')))
  print(NOTES)


if __name__ == '__main__':
  main()
