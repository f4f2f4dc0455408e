#!/bin/bash
# Ring-depth sweep of the GEMV kernel on the whole-model bench (run on a B200: gpurun -- 'bash tools/sweep_smem.sh').
# Prints tokens/s and the GEMV family's us per launch for each setting. Round-1 result (Llama-2-7B Q4, ctx 2048):
#   IFB_GEMV_SMEM_KB=110: 512   150: 559   200: 586   216: 586   216 + IFB_SMEM_QKV=150: 601
# Other knobs read by the library (all default off): IFB_L2_DIST_MB, IFB_PDL_LATE, IFB_GEMV_BATCH, IFB_GEMV_R,
# IFB_GEMV_NOCOMPUTE (streaming only), IFB_ATTN_NO_CLUSTER, IFB_NO_PDL, IFB_NO_GRAPH, IFB_TRACE.
run() { echo "$1: $(env $1 timeout 100 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['us_per_launch_avg'],2))")"; }
run "IFB_GEMV_SMEM_KB=110 IFB_SMEM_QKV=110"
run "IFB_GEMV_SMEM_KB=150 IFB_SMEM_QKV=150"
run "IFB_GEMV_SMEM_KB=216 IFB_SMEM_QKV=216"
run "IFB_GEMV_SMEM_KB=216 IFB_SMEM_QKV=150"
run "IFB_GEMV_SMEM_KB=216 IFB_SMEM_QKV=150 IFB_GEMV_NOCOMPUTE=1"
