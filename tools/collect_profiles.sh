#!/bin/bash
# Everything under profiles/r02_* that one B200 produces, in one pass (run through gpurun from the repo root; the
# outputs land in gpurun_out/ and are copied / summarised into profiles/ by hand afterwards: tools/ncu_summary.py,
# tools/ncu_traffic.py).  Every ncu command runs AFTER the same program has exited 0 without ncu.
#   gpurun --timeout 1500 -- 'bash tools/collect_profiles.sh'
set -u
O=gpurun_out
mkdir -p $O
run() { echo "== $*" >&2; timeout 600 "$@"; }

# the bench line of the final build, default flags
run python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
tail -c 400 $O/r02_bench.err >&2

# secondary tools
run python tools/bench_match.py > $O/r02_bench_match.json 2>/dev/null
run python tools/bench_match.py --G 200 --classes 2 > $O/r02_bench_match_sarship.json 2>/dev/null
run python tools/bench_select.py > $O/r02_bench_select.json 2>/dev/null
run python tools/density_sweep.py > $O/r02_density_sweep.txt 2>/dev/null
run python tools/stream_bw.py > $O/r02_stream_bw.json 2>/dev/null
{ run python tools/bench_loss_kernels.py 81 20 | tail -1; RD_BWD=tile run python tools/bench_loss_kernels.py 81 20 | tail -1;
  RD_BWD=regs run python tools/bench_loss_kernels.py 81 20 | tail -1; run python tools/bench_loss_kernels.py 21 20 | tail -1;
  run python tools/bench_loss_kernels.py 2 20 | tail -1; } > $O/r02_loss_kernels.json 2>/dev/null
run python tools/train_step_time.py 30 2>/dev/null | tail -1 > $O/r02_train_step_time.json
run python tools/bench_a3.py sparse 2>/dev/null | tail -1 > $O/r02_bench_a3.json
[ -x scratch/zero_bw ] && run ./scratch/zero_bw > $O/r02_zero_bw.txt
[ -x scratch/read_bw ] && run ./scratch/read_bw > $O/r02_read_bw.txt

# ncu: launch list of a short bench run, then --set full captures
BENCH_SHORT="python bench.py --steps 3 --warmup 3 --regions 1 --no-e2e --no-cpu-baseline --no-secondary"
run $BENCH_SHORT > $O/short.json 2> $O/short.err && \
  run ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches.csv $BENCH_SHORT > $O/ncu_l.log 2>&1
run $BENCH_SHORT --streams 1 > /dev/null 2>&1 && \
  run ncu --set full --clock-control none --import-source on -k regex:'collect_kernel|graph_kernel|nms_small_kernel|nms_large_kernel' \
      -s 8 -c 4 -f -o $O/prof_r02_stage $BENCH_SHORT --streams 1 > $O/prof_r02_stage.log 2>&1
run python tools/train_once.py > /dev/null 2>&1 && \
  run ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'match_pass|hnm_|conf_loss|loss_reduce|loss_final|loss_backward' -f -o $O/prof_r02_train python tools/train_once.py > $O/prof_r02_train.log 2>&1
run python tools/stage_once.py dense > /dev/null 2>&1 && \
  run ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'collect_kernel|nms_large_kernel' \
      -f -o $O/prof_r02_dense python tools/stage_once.py dense > $O/prof_r02_dense.log 2>&1
echo done >&2
