run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 --no-extra "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err; echo "$tag rc=$?"; grep -h "EngineError\|host enqueue\|passes" gpurun_out/${tag}.err | sort | uniq | head -6; }
run r02i_tr --trace gpurun_out/r02i_trace --e2e-steps 8
python tools/trace_report.py gpurun_out/r02i_trace.rank0.json gpurun_out/r02i_trace.rank1.json
