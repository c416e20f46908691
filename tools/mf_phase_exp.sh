# per-phase SM clocks of K3 (GDS_DUMP_COMP) for 1, 148 and 512 concurrent components
for cfg in "c1 0" "c5 148" "c5 512"; do set -- $cfg
  GDS_DUMP_COMP=gpurun_out/comp_$1_$2.txt timeout 200 python bench.py --workload $1 $( [ $2 != 0 ] && echo --samples $2 ) --steps 1 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
  awk -v tag="$1 $2" 'NR>1{n++; c+=$9; a+=$10; b+=$11; s+=$12; f+=$13; if($9>m)m=$9} END{printf "%s comps %d mean cycles %.0f (max %.0f) gr_init %.0f gr_bfs %.0f gr_snap %.0f front %.0f rounds %.0f\n", tag, n, c/n, m, a/n, b/n, s/n, f/n, (c-a-b-s-f)/n}' gpurun_out/comp_$1_$2.txt
done
