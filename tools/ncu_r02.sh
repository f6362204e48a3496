#!/bin/bash
# round-2 ncu captures (run under gpurun on ONE GPU): one launch of each named kernel, --set full, clocks untouched
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:attn_fwd_kernel -s 6 -c 1 -f -o $O/r02_xattn_fwd_pexport python tools/bench_xattn_one.py > $O/r02_ncu_xattn_fwd.log 2>&1
$NCU -k regex:attn_bwd_kernel -s 4 -c 2 -f -o $O/r02_xattn_bwd_dp python tools/bench_xattn_one.py > $O/r02_ncu_xattn_bwd.log 2>&1
$NCU -k regex:gemm_tc -s 20 -c 1 -f -o $O/r02_conv320_64 python tools/bench_gemm_one.py conv 8 64 320 320 > $O/r02_ncu_conv.log 2>&1
$NCU -k regex:gemm_tc -s 20 -c 1 -f -o $O/r02_lin_32768_320_320 python tools/bench_gemm_one.py 32768 320 320 > $O/r02_ncu_lin.log 2>&1
$NCU -k regex:gemm_tc -s 20 -c 1 -f -o $O/r02_lin_2048_1280_1280 python tools/bench_gemm_one.py 2048 1280 1280 > $O/r02_ncu_lin2.log 2>&1
$NCU -k regex:gn_fused_kernel -s 4 -c 2 -f -o $O/r02_gn_fused python -c "
import sys; sys.path.insert(0, '.')
import torch
from comat_b200 import ops
x = torch.randn(8, 4096, 320, device='cuda').half(); dy = torch.randn_like(x)
g, b = torch.ones(320, device='cuda'), torch.zeros(320, device='cuda')
for _ in range(3):
    y, mr = ops.groupnorm_fwd(x, g, b, 32, 1e-5, True); ops.groupnorm_bwd(x, dy, g, b, mr, 32, True)
torch.cuda.synchronize()" > $O/r02_ncu_gn.log 2>&1
ls -la $O/*.ncu-rep
