# per-phase SM clocks of K3 (GDS_DUMP_COMP): first relabel (thread 0: loop body vs barrier wait) and rounds
for w in c1; do
  GDS_DUMP_COMP=gpurun_out/comp_$w.txt timeout 200 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
  cat gpurun_out/comp_$w.txt
done
