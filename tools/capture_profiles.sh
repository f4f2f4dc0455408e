set -x
mkdir -p gpurun_out/final
# 1. launch list of eager decode steps (every launch visible), filtered later
IFB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/final/launches_bench.log 2>&1
# 2. --set full of the GEMV kernel (QKV, wo, GLU, w2 of one layer) and the attention kernel
IFB_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemv_stream -s 129 -c 4 -o gpurun_out/final/prof_gemv python bench.py --steps 1 --warmup 3 > gpurun_out/final/ncu_gemv.log 2>&1
IFB_NO_GRAPH=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 33 -c 2 -o gpurun_out/final/prof_attn python bench.py --steps 1 --warmup 3 > gpurun_out/final/ncu_attn.log 2>&1
# 3. timeline + bench (both arms)
python tools/trace_step.py --graph --attn > gpurun_out/final/trace.log 2>&1
python bench.py --steps 300 --warmup 5 > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err
ls -la gpurun_out/final
