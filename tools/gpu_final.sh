#!/bin/bash
# what the driver does at round end, plus the ncu evidence for profiles/
R=${1:-r1e}
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref_$R.json
echo "=== bench"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_$R.json
kill $SMI
NCU=/usr/local/cuda/bin/ncu
echo "=== ncu launch list"
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -k regex:'dual_matvec|epilogue_kernel|series_init|resident_series' -c 400 \
    --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/bench_under_ncu_$R.log 2>&1
echo "=== ncu full"
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:dual_matvec_tma -s 30 -c 2 -f -o gpurun_out/prof_$R \
    python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/ncu_full_$R.log 2>&1
echo "=== ncu full, resident series kernel (N=900)"
timeout 400 $NCU --set full --clock-control none --import-source on -k regex:resident_series -s 3 -c 1 -f -o gpurun_out/prof_${R}_res \
    python bench.py --basis 900 --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-65k > gpurun_out/ncu_full_${R}_res.log 2>&1
ls -la gpurun_out | tail -8
