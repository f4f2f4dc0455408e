run() { echo "$1: $(env $1 python bench.py --steps 200 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['us_per_launch_avg'],2))")"; }
B="IFB_GEMV_SMEM_KB=216 IFB_SMEM_QKV=150"
run "$B IFB_L2_DIST_MB=4"
run "$B IFB_L2_DIST_MB=16"
run "$B IFB_L2_DIST_MB=48"
run "IFB_GEMV_SMEM_KB=110 IFB_L2_DIST_MB=16"
run "$B IFB_SMEM_QKV=140"
run "$B IFB_SMEM_QKV=156"
run "$B IFB_SMEM_HEAD=110"
