out=gpurun_out; tag=r02t
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s 8 -c 8 --csv --log-file $out/${tag}_launches_c2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra --device-pass-only > $out/${tag}_ncu_c2.log 2>&1
python tools/ncu_summary.py $out/${tag}_launches_c2.csv
for f in 64 32; do RL_L2_FETCH=$f timeout 200 python bench.py --workload C3 --steps 30 --warmup 5 --no-extra --no-cpu-baseline --device-pass-only --e2e-steps 4 > $out/${tag}_c3_l2f$f.json 2> $out/${tag}_c3_l2f$f.err; grep -h "passes" $out/${tag}_c3_l2f$f.err; done
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kats.py tests/test_shard_peer.py -m gpu -x -q -k "kat1 or kat7 or wraps or hot_owner or matches_global_oracle" > $out/${tag}_san_memcheck.log 2>&1; tail -4 $out/${tag}_san_memcheck.log
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "records_fast_path and 7" > $out/${tag}_san_racecheck.log 2>&1; tail -4 $out/${tag}_san_racecheck.log
timeout 300 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "records_fast_path and 7" > $out/${tag}_san_synccheck.log 2>&1; tail -4 $out/${tag}_san_synccheck.log
