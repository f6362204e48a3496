#!/bin/bash
# A/B of the attention forward variants on the 64x64-latent self-attention shape (one process per variant: the env is read once)
for v in "COMAT_ATTN_LONG=0" "COMAT_ATTN_LONG=1 COMAT_ATTN_POLY=0" "COMAT_ATTN_LONG=1 COMAT_ATTN_POLY=3"; do
  echo "== $v"; env $v python tools/bench_attn_one.py 2>&1 | grep "attn"
done
