timeout 600 python -m pytest tests/test_nets_gpu.py tests/test_conv_slab_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --workload overlap --steps 10 --warmup 3 > gpurun_out/bench22_overlap.json 2> gpurun_out/bench22_overlap.err; tail -c 300 gpurun_out/bench22_overlap.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench22_overlap.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], round(d["roofline"]["frac"],4))
print([(k["kernel"],k["launches_per_step"],round(k["ms_per_step"],3)) for k in d["extra"]["kernels"]])
P
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_overlap_v22.csv python bench.py --workload overlap --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_v22.log 2>&1; tail -1 gpurun_out/ncu_v22.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -c 29 -o /tmp/overlap_step_v5 -f python scripts/trace_overlap.py > gpurun_out/ncu_overlap_step_v5.log 2>&1; tail -2 gpurun_out/ncu_overlap_step_v5.log
python scripts/ncu_summary.py /tmp/overlap_step_v5.ncu-rep gpurun_out/overlap_step_v5_ncu_summary.txt > /dev/null 2>&1; ls -la /tmp/overlap_step_v5.ncu-rep gpurun_out/overlap_step_v5_ncu_summary.txt
du -sh gpurun_out
