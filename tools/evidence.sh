#!/bin/bash
# Round-end evidence on ONE B200 box (run under gpurun from the repo root):  bash tools/evidence.sh <tag>
# Writes small text / JSON files only into gpurun_out/<tag>/ (ncu reports are summarised on the box and deleted: the
# merge-back limit of gpurun_out/ is 64 MiB).  Copy what should be judged into profiles/.
set -u
O=gpurun_out/${1:-evidence}
mkdir -p "$O"
OWN='regex:^(cost_volume|warp_(fwd|bwd)_n|photo_|smooth_|consis_|pyramid_|upsample_(fwd|bwd)|bias_lrelu|splat)'
( timeout 900 python -m pytest tests -m gpu -q > "$O/pytest_gpu.log" 2>&1; echo "rc=$?" >> "$O/pytest_gpu.log" )
timeout 500 python bench.py > "$O/bench.json" 2> "$O/bench.err"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > "$O/bench_reference.json" 2> "$O/bench_reference.err"
timeout 300 python -m unopticalflow_b200.kernel_bench --json "$O/kernel_bench.json" > "$O/kernel_bench.txt" 2>&1
# launch list of ONE training step (same command as the bench, --profile-step), with DRAM bytes per launch
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file "$O/launches.csv" python bench.py --profile-step > "$O/profile_step.log" 2>&1
cp gpurun_out/profile_step_calls.json "$O/" 2>/dev/null
python profiles/make_traffic_json.py "$O/launches.csv" "$O/profile_step_calls.json" "$O/ncu_traffic.json" >> "$O/profile_step.log" 2>&1
python profiles/summarize_launches.py "$O/launches.csv" > "$O/launches_summary.txt" 2>&1
# ncu --set full of every hand-written kernel at the benchmark shapes (kernel_bench --once), summarised here
timeout 900 ncu --set full --clock-control none -k "$OWN" -o "$O/full" -f python -m unopticalflow_b200.kernel_bench --once > "$O/ncu_full.log" 2>&1
python profiles/ncu_summary.py "$O/full.ncu-rep" > "$O/ncu_full_summary.txt" 2>&1
rm -f "$O/full.ncu-rep"
tail -3 "$O/pytest_gpu.log"; head -c 400 "$O/bench.json"; echo; ls -la "$O"
