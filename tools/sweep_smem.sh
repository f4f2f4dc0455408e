run() { echo "$1: $(env $1 timeout 100 python bench.py --steps 200 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['us_per_launch_avg'],2))")"; }
run "IFB_GEMV_BATCH=0"
run "IFB_GEMV_BATCH=1"
run "IFB_GEMV_BATCH=0"
run "IFB_GEMV_BATCH=1"
