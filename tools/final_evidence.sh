set -x
timeout 200 python bench.py --cpu-all-cores > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err
timeout 200 python bench.py --config 3 --steps 5 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 300 python bench.py --config 5 --steps 10 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
timeout 400 python bench.py --config 4 --steps 3 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --use-graph 0 --loop-mode 0 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:spmv_staged -s 300 -c 90 --csv --log-file gpurun_out/r2_warm_traffic.csv python bench.py --steps 1 --warmup 3 --use-graph 0 --loop-mode 0 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_staged -s 300 -c 10 -o gpurun_out/r2_prof_fine python bench.py --steps 1 --warmup 3 --use-graph 0 --loop-mode 0 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r2_prof_fine.ncu-rep --page raw --csv > gpurun_out/r2_spmv_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_prof_fine.ncu-rep
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo memcheck rc=$? >> gpurun_out/r2_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo racecheck rc=$? >> gpurun_out/r2_sanitizer_racecheck.log
tail -5 gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
