#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, both bench arms, the ncu launch list of the
# bench command and one `--set full` capture of each kernel of the step.
#   bash tools/profile_round.sh r1        -> gpurun_out/{bench,ref,launches,prof}_r1.*
# then here:  python profiles/summarize.py gpurun_out/prof_r1.ncu-rep gpurun_out/launches_r1.csv r1 16384
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$tag.txt
python bench.py --impl reference > gpurun_out/ref_$tag.json 2> gpurun_out/ref_$tag.err
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cut -c1-300 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-sustained > gpurun_out/ncu_launch_$tag.log 2>&1
# second eager warm-up step: one instance each of fwd_pre, post (value-and-grad), bwd_pre
ncu --set full --clock-control none --import-source on \
    -k regex:'fwd_pre_kernel|post_kernel|bwd_pre_kernel' -s 3 -c 3 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-sustained > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/launches_$tag.csv
# kernel sweep and the "next" rows' per-kernel rooflines
bash tools/sweep.sh
python tools/bench_next_rows.py > gpurun_out/next_rows.jsonl 2> gpurun_out/next_rows.err; tail -2 gpurun_out/next_rows.err
