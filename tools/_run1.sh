set -x
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json; echo
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 11000 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_run.py cfg3 2048 > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.csv; cat gpurun_out/r2_launches_summary.csv
ncu --set full --clock-control none --import-source on -k regex:score_conf -s 600 -c 3 -o gpurun_out/r2_score_conf python tools/prof_run.py cfg3 2048 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:grow_cluster -s 2 -c 1 -o gpurun_out/r2_grow_cluster python tools/time_drivers.py cfg2 128 3 > gpurun_out/ncu_full2.log 2>&1; tail -3 gpurun_out/ncu_full2.log
ls -la gpurun_out/
