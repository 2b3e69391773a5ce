"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/rl_engine.h declares; the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from limitador_b200 import build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="rl_engine.h"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build_engine()
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rl_engine.h but not exported"
    assert sorted(engine.ABI_SYMBOLS) == names
    # the CPU front's header (limits -> counters)
    from limitador_b200 import matcher
    mnames = declared_functions("rl_match.h")
    assert len(mnames) == 18 and sorted(matcher.MATCH_SYMBOLS) == mnames
    for n in mnames:
        assert hasattr(lib, n), f"{n} declared in include/rl_match.h but not exported"
    # the RLS wire surface
    from limitador_b200 import rls
    rnames = declared_functions("rl_rls.h")
    assert sorted(rls.RLS_SYMBOLS) == rnames
    for n in rnames:
        assert hasattr(lib, n), f"{n} declared in include/rl_rls.h but not exported"
    # the replicated counter value
    from limitador_b200 import crdt
    cnames = declared_functions("rl_crdt.h")
    assert sorted(crdt.CRDT_SYMBOLS) == cnames
    for n in cnames:
        assert hasattr(lib, n), f"{n} declared in include/rl_crdt.h but not exported"


def test_binary_targets_sm_100a_only():
    path = build.build_engine()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    assert engine.RECORD_DTYPE.itemsize == 32
    assert engine.COUNTER_DTYPE.itemsize == 24
    assert engine.LIMIT_DESC_DTYPE.itemsize == 32
    assert ctypes.sizeof(engine.RlConfig) == 40
    assert ctypes.sizeof(engine.RlStats) == 136
    # include/rl_rls.h, include/rl_crdt.h
    from limitador_b200 import crdt, rls
    assert rls.ENTRY_DTYPE.itemsize == 20 and ctypes.sizeof(rls.RlsRequest) == 20
    assert ctypes.sizeof(crdt.CrdtConfig) == 24 and crdt.KEY_DTYPE.itemsize == 16 and crdt.UPDATE_DTYPE.itemsize == 32
    hdr = open(os.path.join(ROOT, "include", "rl_engine.h")).read()
    assert len(re.findall(r"uint64_t \w+;", hdr[hdr.index("typedef struct rl_compact_stats"):hdr.index("} rl_compact_stats;")])) == 6


def test_owner_of_is_pure_host_function():
    lib = engine.load_library()
    owners = [lib.rl_owner_of(ns, 8) for ns in range(1000)]
    assert set(owners) == set(range(8))
    assert all(lib.rl_owner_of(ns, 1) == 0 for ns in range(10))


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.EngineError):
        engine.Engine(1024)
    from limitador_b200 import crdt
    with pytest.raises(crdt.CrdtError, match="no CPU implementation"):
        crdt.CrdtTable(1024, 2, 0)
