timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; tail -3 gpurun_out/r02l_pytest.log
python bench.py --steps 300 --warmup 20 --no-extra --no-cpu-baseline --trace gpurun_out/r02l_trace --e2e-steps 8 > gpurun_out/r02l_tr.json 2> gpurun_out/r02l_tr.err; grep -h "passes\|host enq\|engine stats" gpurun_out/r02l_tr.err
python tools/trace_report.py gpurun_out/r02l_trace.rank0.json
python tools/hot_trace.py gpurun_out/r02l_trace.rank0.json
python bench.py --steps 400 --warmup 20 > gpurun_out/r02l_bench_c2.json 2> gpurun_out/r02l_bench_c2.err; echo "bench rc=$?"; grep -h "passes\|host enq\|extra" gpurun_out/r02l_bench_c2.err | cut -c1-1500; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_c2.json'))
print("C2 value %.3f G/s  %.1f us/step  e2e %.3f G  k_main+hot %.1f us frac %.4f" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e9, d["roofline"]["avg_launch_ms"]*1e3, d["roofline"]["frac"]))
PY
