out=gpurun_out; tag=r02w
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
timeout 200 python bench.py --steps 400 --warmup 20 --no-extra --no-cpu-baseline --e2e-steps 64 > $out/${tag}_c2.json 2> $out/${tag}_c2.err; python tools/bench_summary.py $out/${tag}_c2.json | head -3
