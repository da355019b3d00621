#!/bin/bash
# GPU-side (run under gpurun): the round's evidence — one bench line per workload, ncu launch lists of the bench command and
# one `ncu --set full` capture per dominant kernel.  Reports land in gpurun_out/; tools/collect_profiles.py reads them on the CPU box.
R=${1:-r02}
REP=${AW_REP_DIR:-/tmp/ncu_$R}          # the .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): they are summarised here
mkdir -p gpurun_out $REP
[ -n "$SKIP_BENCH" ] || bash tools/bench_all.sh C2 C1 C3 C4 C5-64 C5-128 C5-512 C5-1024 C5-2048 C5-4096 F3 | tee gpurun_out/${R}_bench_all.txt
# launch list of the default bench command: skip the 2 x 40 warm-up launches (+ set-up kernels), list 40 launches of the timed region
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv --log-file gpurun_out/${R}_launches_C2.csv python bench.py --steps 100 --warmup 40 --no-cpu --e2e-steps 3 --no-single-block > gpurun_out/ncu_launches_C2.log 2>&1; echo "launch list C2 rc=$?"
# C4 launches two kernels per step (KP + EQ): warm-up = 2 x 80 launches + set-up
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 40 --csv --log-file gpurun_out/${R}_launches_C4.csv python bench.py --workload C4 --steps 100 --warmup 40 --no-cpu --e2e-steps 3 --no-single-block > gpurun_out/ncu_launches_C4.log 2>&1; echo "launch list C4 rc=$?"
for w in C2 C3 C4 C5-512 C5-64 C5-2048; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 45 -c 1 -o $REP/${R}_full_$w -f python bench.py --workload $w --steps 4 --warmup 41 --no-cpu --e2e-steps 3 --no-single-block > gpurun_out/ncu_$w.log 2>&1; echo "$w rc=$?"
done
# the same kernel with one block per launch (the latency-oriented mode)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 45 -c 1 -o $REP/${R}_full_C2k1 -f python bench.py --workload C2 --blocks-per-call 1 --steps 4 --warmup 41 --no-cpu --e2e-steps 3 > gpurun_out/ncu_C2k1.log 2>&1; echo "C2k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_eq_systolic -s 45 -c 1 -o $REP/${R}_full_C4eq -f python bench.py --workload C4 --steps 4 --warmup 41 --no-cpu --e2e-steps 3 --no-single-block > gpurun_out/ncu_C4eq.log 2>&1; echo "C4eq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_input_rfft -s 45 -c 1 -o $REP/${R}_full_C5-4096_k2 -f python bench.py --workload C5-4096 --steps 4 --warmup 41 --no-cpu --e2e-steps 3 > gpurun_out/ncu_k2.log 2>&1; echo "K2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_irfft_out -s 45 -c 1 -o $REP/${R}_full_C5-4096_k4 -f python bench.py --workload C5-4096 --steps 4 --warmup 41 --no-cpu --e2e-steps 3 > gpurun_out/ncu_k4.log 2>&1; echo "K4 rc=$?"
# summarise on the box (ncu reads its own reports), keep the text + the C2 report
AW_REP_DIR=$REP python tools/collect_profiles.py $R && cp profiles/${R}_* gpurun_out/ && cp $REP/${R}_full_C2.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -20
