"""One short GPU-box visit for the entry points added without GPU time: runs their -m gpu tests as plain
functions, most valuable first, WITHOUT importing torch until the very end (a fresh box can spend a minute on that import),
and rewrites gpurun_out/r02_new_gpu.json after every test so that a visit cut short still reports what ran.
    gpurun --timeout 100 -- 'timeout 95 python tools/gpu_quick_new.py'"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "r02_new_gpu.json")
T0 = time.time()
results = {}


def run(name, fn):
    t = time.time()
    try:
        fn()
        results[name] = f"ok ({time.time() - t:.1f} s, at +{time.time() - T0:.0f} s)"
    except BaseException as ex:  # noqa: BLE001 — report everything, keep going
        results[name] = {"failed": f"{type(ex).__name__}: {ex}"[:400], "trace": traceback.format_exc()[-1500:]}
    with open(OUT, "w") as f:
        json.dump(results, f, indent=1)
    print(name, "->", results[name] if isinstance(results[name], str) else "FAILED: " + results[name]["failed"], flush=True)


def main():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    import tests.test_zz2_crdt_gpu as C
    import tests.test_zz3_maint_gpu as M
    import tests.test_zz1_rls_gpu as R
    run("crdt_session_2_actors", lambda: C.test_gpu_sessions_match_the_oracle(11, 2, 0))
    run("crdt_errors_and_clear", C.test_errors_are_loud)
    run("compact_cells1_regions8", lambda: M.test_compact_after_sweep_is_invisible_and_the_hot_path_finds_every_row_again(1, 8))
    run("ns_metrics_flags0", lambda: M.test_ns_metrics_accumulate_behind_every_record_call(0))
    run("rls_serve", R.test_serve_through_the_engine_equals_the_cpu_stages_around_the_oracle)
    run("front_with_matcher", R.test_front_with_the_matcher_inside_concurrent_callers_linearise)
    run("compact_cells3_regions4", lambda: M.test_compact_after_sweep_is_invisible_and_the_hot_path_finds_every_row_again(3, 4))
    run("compact_cells7_regions16", lambda: M.test_compact_after_sweep_is_invisible_and_the_hot_path_finds_every_row_again(7, 16))
    run("compact_no_tombstones", M.test_compact_on_a_table_without_tombstones_does_nothing)
    run("ns_metrics_pipeline_flag", lambda: M.test_ns_metrics_accumulate_behind_every_record_call(2))
    run("crdt_large_batches", C.test_large_batches_match_the_oracle)
    run("crdt_two_replicas", C.test_two_gpu_replicas_converge_through_export_and_merge)
    run("crdt_session_16_actors", lambda: C.test_gpu_sessions_match_the_oracle(13, 16, 7))
    run("crdt_session_3_actors", lambda: C.test_gpu_sessions_match_the_oracle(12, 3, 2))
    run("crdt_session_1_actor", lambda: C.test_gpu_sessions_match_the_oracle(14, 1, 0))
    run("ns_metrics_pipelined_device_calls (imports torch)", M.test_ns_metrics_behind_pipelined_device_calls)


if __name__ == "__main__":
    main()
