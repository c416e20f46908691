#!/bin/bash
# One GPU call: tests, smoke, the bench lines of every config, the sweep, K3 probes.
# usage (through gpurun): bash tools/r02_final_measure.sh <tag>
T=${1:-r02z3}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/${T}_box.txt; nproc >> $O/${T}_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.txt
tail -3 $O/${T}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.txt 2>&1; tail -1 $O/${T}_smoke.txt
for w in c4 c1 c2 ref_hole; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/${T}_bench_$w.json 2> $O/${T}_bench_$w.err; echo "$w rc=$?"
done
timeout 900 python bench.py > $O/${T}_BENCH_c5.json 2> $O/${T}_BENCH_c5.err; echo "c5 rc=$?"
for w in c1 c5 c4 c2; do
  timeout 600 python bench.py --workload $w --algorithm mcp --no-cpu-baseline > $O/${T}_bench_mcp_$w.json 2> $O/${T}_bench_mcp_$w.err; echo "mcp $w rc=$?"
done
timeout 300 python tools/k3_probe.py c4 1 > $O/${T}_k3_probe_c4.txt 2>&1
for e in 4 8 12 16; do
  timeout 600 python bench.py --no-cpu-baseline --steps 5 --encode-threads $e > $O/${T}_bench_c5_enc$e.json 2>/dev/null; echo "enc$e rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/${T}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e",{})
        print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e %.4g"%e.get("value",0), {k:v for k,v in e.items() if k.startswith("ms") or "encode" in k or "u32" in k})
    except Exception as ex:
        print(f, "ERR", ex)
PY
