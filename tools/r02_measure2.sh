#!/bin/bash
O=gpurun_out
python tools/bench_attnmap.py > $O/r02_attnmap_bench_v3.json 2> $O/r02_attnmap_bench_v3.err; cut -c1-600 $O/r02_attnmap_bench_v3.json
python tools/bench_xattn_one.py > $O/r02_xattn_bench2.log 2>&1; cat $O/r02_xattn_bench2.log
: > $O/r02_config5_sweep_graphed.jsonl
for L in 64 96 128; do for B in 1 2 4; do
  timeout 300 python bench.py --config 5 --batch $B --latent $L --steps 3 --warmup 3 >> $O/r02_config5_sweep_graphed.jsonl 2>> $O/r02_config5_sweep_graphed.err
done; done
python - <<'PY'
import json
for l in open('gpurun_out/r02_config5_sweep_graphed.jsonl'):
    d=json.loads(l); print(d['config']['workload'][-58:], round(d['ms_per_step'],1), d['roofline']['achieved'] and round(d['roofline']['achieved'],1), d['config'].get('cuda_graphs_taped_calls'))
PY
tail -3 $O/r02_config5_sweep_graphed.err
