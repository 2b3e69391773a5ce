timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; tail -3 gpurun_out/r02k_pytest.log
python bench.py --steps 300 --warmup 10 --no-extra --no-cpu-baseline --trace gpurun_out/r02k_trace --e2e-steps 8 > gpurun_out/r02k_tr.json 2> gpurun_out/r02k_tr.err; grep -h "passes\|host enq\|engine stats" gpurun_out/r02k_tr.err
python tools/trace_report.py gpurun_out/r02k_trace.rank0.json
python tools/hot_trace.py gpurun_out/r02k_trace.rank0.json
