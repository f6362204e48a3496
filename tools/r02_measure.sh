#!/bin/bash
# round-2 measurement batch (ONE GPU): attention-map loss API / kernel rates, BASELINE configs[4] sweep, ncu --set full captures
O=gpurun_out
python tools/bench_attnmap.py > $O/r02_attnmap_bench_v2.json 2> $O/r02_attnmap_bench_v2.err; cat $O/r02_attnmap_bench_v2.json
python tools/bench_xattn_one.py > $O/r02_xattn_bench.log 2>&1; cat $O/r02_xattn_bench.log
: > $O/r02_config5_sweep.jsonl
for L in 64 96 128; do for B in 1 2 4; do
  timeout 300 python bench.py --config 5 --batch $B --latent $L --steps 3 --warmup 2 >> $O/r02_config5_sweep.jsonl 2>> $O/r02_config5_sweep.err
done; done
cut -c1-160 $O/r02_config5_sweep.jsonl
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd_long -s 3 -c 1 -f -o $O/r02_attn_fwd_long_v3 python tools/bench_attn_one.py > $O/r02_ncu_a1.log 2>&1
$NCU -k regex:attn_bwd_kernel -s 2 -c 2 -f -o $O/r02_attn_bwd_v3 python tools/bench_attn_one.py > $O/r02_ncu_a2.log 2>&1
$NCU -k regex:attn_fwd_kernel -s 6 -c 1 -f -o $O/r02_xattn_fwd_pexport_v3 python tools/bench_xattn_one.py > $O/r02_ncu_a3.log 2>&1
KERNEL=pair BN=160 $NCU -k regex:gemm_tc_pair -s 20 -c 1 -f -o $O/r02_conv_pair160_1920_1280 python tools/bench_gemm_one.py conv 8 16 1920 1280 > $O/r02_ncu_a4.log 2>&1
ls -la $O/*_v3.ncu-rep $O/r02_conv_pair160_1920_1280.ncu-rep
