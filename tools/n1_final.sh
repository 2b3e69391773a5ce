out=gpurun_out; tag=r02x
timeout 400 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "rc=$?"; grep -h "bench +" $out/${tag}_bench_default.err | cut -c1-160 | tail -14; python tools/bench_summary.py $out/${tag}_bench_default.json
