python bench.py --steps 400 --warmup 20 --no-extra --no-cpu-baseline --trace gpurun_out/r02p_trace > gpurun_out/r02p_c2.json 2> gpurun_out/r02p_c2.err; grep -h "passes\|host enq" gpurun_out/r02p_c2.err
python tools/trace_report.py gpurun_out/r02p_trace.rank0.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_c2.json'))
print("C2 value %.3f G/s  %.1f us/step  e2e %.3f G  replay %.1f us frac %.4f" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e9, d["roofline"]["avg_launch_ms"]*1e3, d["roofline"]["frac"]))
PY
python bench.py --workload C3 --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02p_c3.json 2> gpurun_out/r02p_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_c3.json'))
print("C3 value %.3f G/s  %.1f us/step  e2e %.3f G  replay %.1f us frac %.4f" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e9, d["roofline"]["avg_launch_ms"]*1e3, d["roofline"]["frac"]))
PY
