#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the dominant kernel
R=${1:-r1}
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
echo "=== launch list (bench.py --steps 2 --warmup 1, our kernels only)"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -k regex:'dual_matvec|epilogue_kernel|series_init' -c 400 \
    --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/bench_under_ncu_$R.log 2>&1
tail -3 gpurun_out/launches_$R.csv
echo "=== full capture of the dominant kernel"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:dual_matvec_tma -s 30 -c 2 -f -o gpurun_out/prof_$R \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/ncu_full_$R.log 2>&1
ls -la gpurun_out/
echo "=== LDG variant for comparison"
timeout 300 python bench.py --steps 20 --warmup 3 --skip-cpu --skip-e2e --skip-65k --kernel ldg 2>&1 | tail -1 | tee gpurun_out/bench_ldg_$R.json
echo "=== clocks during a normal run"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
timeout 600 python bench.py --steps 100 --warmup 5 --e2e-steps 2 2>&1 | tail -1 | tee gpurun_out/bench_$R.json
kill $SMI
