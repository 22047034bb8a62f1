# Round-2 evidence run (one gpurun call): headline bench line, Zara-shaped / dense-crowd lines, ncu launch list, sanitizers.
set -x
python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err
python bench.py --agents-per-scene 32 --scenes 4096 --no-train --no-cpu-baseline > gpurun_out/r2_bench_zara_32x20.json 2>/dev/null
python bench.py --agents-per-scene 256 --k 128 --scenes 64 --no-train --no-cpu-baseline > gpurun_out/r2_bench_dense_256x128.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/r2_memcheck.log 2>&1; tail -3 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > gpurun_out/r2_racecheck.log 2>&1; tail -3 gpurun_out/r2_racecheck.log
