#!/bin/bash
# Round-2 measurement artefacts (copied to profiles/ by hand afterwards): stage breakdowns, ncu --set full of
# every kernel of one sub-batch per message type, the launch list of the bench command.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv,noheader > $O/gpu.txt 2>&1
for m in 5 4 3; do
  timeout 300 python tools/prof_l7.py --msg $m --records 625000 > $O/stage_breakdown_msg$m.txt 2>&1
done
timeout 300 python tools/prof_l7.py --lz4 --runs 4 > $O/stage_breakdown_lz4.txt 2>&1
# one full sub-batch of each type: skip the 4 warm-up passes (3 + the fetch re-run is absent in steady state)
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"^k_|k_l7|k_frame|k_slot|k_finish|k_commit|k_prepare" --launch-skip 66 -c 22 -f -o $O/ncu_l7 python tools/prof_l7.py --runs 1 > $O/ncu_l7.log 2>&1; echo "ncu l7 rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_l4|k_frame_walk" --launch-skip 15 -c 5 -f -o $O/ncu_l4 python tools/prof_l7.py --msg 4 --runs 1 > $O/ncu_l4.log 2>&1; echo "ncu l4 rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"k_m_decode|k_m_tag|k_m_emit" --launch-skip 9 -c 3 -f -o $O/ncu_metrics python tools/prof_l7.py --msg 3 --runs 1 > $O/ncu_metrics.log 2>&1; echo "ncu m rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_lz4 --launch-skip 3 -c 3 -f -o $O/ncu_lz4 python tools/prof_l7.py --lz4 --runs 1 > $O/ncu_lz4.log 2>&1; echo "ncu lz4 rc=$?"
# launch lists: durations of every launch (cold cache, serialised): one sub-batch, and the bench command itself
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_l7_subbatch.csv python tools/prof_l7.py --runs 2 > $O/launches_l7.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --legs c2 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "bench launch list rc=$?"
ls -la $O
