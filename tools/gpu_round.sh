#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch lists.  Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_round.sh <tag> [what...]   what: tests newtests bench c3 ncu ncufull san sannew
tag=${1:-r02}; shift
what=${*:-tests bench c3 ncu}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
for w in $what; do
case $w in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log;;
newtests)  # the entry points added without GPU time: their tests alone, each file reported
  for f in tests/test_zz2_crdt_gpu.py tests/test_zz3_maint_gpu.py tests/test_zz1_rls_gpu.py; do
    timeout 600 python -m pytest $f -m gpu -q > $out/${tag}_$(basename $f .py).log 2>&1; echo "$f rc=$?"; tail -3 $out/${tag}_$(basename $f .py).log; done;;
sannew)
  timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_zz2_crdt_gpu.py tests/test_zz3_maint_gpu.py -m gpu -x -q -k "not large" > $out/${tag}_san_new_memcheck.log 2>&1; tail -5 $out/${tag}_san_new_memcheck.log
  timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_zz3_maint_gpu.py -m gpu -x -q -k "metrics" > $out/${tag}_san_new_racecheck.log 2>&1; tail -5 $out/${tag}_san_new_racecheck.log;;
bench)
  timeout 600 python bench.py --steps 1000 --warmup 20 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; echo "bench rc=$?"; grep -h "passes\|host enq" $out/${tag}_bench_c2.err; python tools/bench_summary.py $out/${tag}_bench_c2.json;;
c3)
  timeout 600 python bench.py --workload C3 --steps 60 --warmup 5 > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; echo "c3 rc=$?"; tail -c 600 $out/${tag}_bench_c3.json;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s 8 -c 8 --csv --log-file $out/${tag}_launches_c2.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra --device-pass-only > $out/${tag}_ncu_c2.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -s 8 -c 6 --csv --log-file $out/${tag}_launches_c3.csv python bench.py --workload C3 --steps 4 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra --device-pass-only > $out/${tag}_ncu_c3.log 2>&1
  python tools/ncu_summary.py $out/${tag}_launches_c2.csv $out/${tag}_launches_c3.csv;;
ncufull)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_main -s 8 -c 1 -o $out/${tag}_k_main_c2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra --device-pass-only > $out/${tag}_ncufull.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_front -s 8 -c 1 -o $out/${tag}_k_front_c2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra --device-pass-only >> $out/${tag}_ncufull.log 2>&1;;
kstats)
  timeout 300 python bench.py --steps 200 --warmup 10 --kstats --no-cpu-baseline --no-extra > $out/${tag}_kstats_c2.json 2> $out/${tag}_kstats_c2.err; grep "engine stats" $out/${tag}_kstats_c2.err
  timeout 300 python bench.py --steps 200 --warmup 10 --kstats --no-cpu-baseline --no-extra --no-pipeline > $out/${tag}_kstats_c2_np.json 2> $out/${tag}_kstats_c2_np.err; grep "engine stats" $out/${tag}_kstats_c2_np.err;;
ncuprobe)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_count -s 8 -c 1 -o $out/${tag}_k_probe_c2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-pipeline --no-extra > $out/${tag}_ncuprobe.log 2>&1;;
san)
  timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or chained or hot_key or csr_general" > $out/${tag}_san_memcheck.log 2>&1; tail -5 $out/${tag}_san_memcheck.log
  timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or chained or hot_key" > $out/${tag}_san_racecheck.log 2>&1; tail -5 $out/${tag}_san_racecheck.log
  timeout 1500 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path or chained or hot_key" > $out/${tag}_san_synccheck.log 2>&1; tail -5 $out/${tag}_san_synccheck.log;;
esac
done
