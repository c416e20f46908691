# round-end measurement set (run on a GPU box from the repo root); everything lands in gpurun_out/
set -x
T=$1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/${T}_BENCH_c5.json 2> gpurun_out/${T}_BENCH_c5.err
timeout 400 python bench.py --impl reference > gpurun_out/${T}_BENCH_reference.json 2> gpurun_out/${T}_BENCH_reference.err
for w in c1 c2 c4; do timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err; done
timeout 300 python bench.py --e2e-u32 --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_c5_e2e_u32.json 2> /dev/null
timeout 300 env GDS_BUNDLE=sort python bench.py --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_c5_sortpath.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_ncu_launches_c5.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_maxflow|k_direct_hist|k_direct_mark|k_scan_lookback" -c 6 -o gpurun_out/${T}_prof_c5 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
tail -2 gpurun_out/${T}_pytest_gpu.txt
python tools/show_bench.py gpurun_out/${T}_BENCH_c5.json gpurun_out/${T}_bench_c1.json gpurun_out/${T}_bench_c2.json gpurun_out/${T}_bench_c4.json gpurun_out/${T}_bench_c5_sortpath.json
cat gpurun_out/${T}_BENCH_reference.json | cut -c1-600
timeout 300 python tools/gpu_fuzz.py ${FUZZ_CASES:-1500} ${FUZZ_SEED:-77} 2>&1 | tail -2 > gpurun_out/${T}_gpu_fuzz.txt; cat gpurun_out/${T}_gpu_fuzz.txt
