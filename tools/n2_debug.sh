timeout 200 python -m pytest tests/test_shard_peer.py -m gpu -x -q 2>&1 | tail -2
run() { tag=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 150 --warmup 6 --parity-steps 2 --no-extra --no-cpu-baseline --e2e-steps 16 "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; echo "$tag rc=$?"; grep -h "EngineError\|host enqueue\|passes\|parity:" gpurun_out/${tag}.err | sort | uniq | head -6; }
run r02o_tr --trace gpurun_out/r02o_trace
python tools/trace_report.py gpurun_out/r02o_trace.rank0.json gpurun_out/r02o_trace.rank1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o_tr.json')); print("N=2 value %.3f G/s %.1f us/step e2e %.3f" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["e2e"]["value"]/1e9))
PY
