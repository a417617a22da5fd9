#!/bin/bash
# round-2 end-of-round evidence: mid-kernel ncu capture, bench line (with mid_operator), launch list
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
echo "=== ncu full, mid-size series kernel (N=4096, one 24-term launch)"
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:mid_series -s 4 -c 1 -f -o gpurun_out/mid_ncu \
    python tools/gpu_mid.py --sizes 4096 --l2mb 0 --reps 4 2>&1 | tail -2
echo "=== bench"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r2d.csv &
SMI=$!
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_r2d.json | cut -c1-400
kill $SMI
echo "=== ncu launch list of the bench command"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -k regex:'dual_matvec|epilogue_kernel|series_init|resident_series|mid_series' -c 600 \
    --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-65k > gpurun_out/bench_under_ncu_r2d.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu_r2d.log
ls -la gpurun_out | tail -8
