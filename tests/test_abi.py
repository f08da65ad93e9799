"""CPU tests of the drop-in boundary: libhpf_b200.so loads, exports every symbol
include/hpf_cuda.h declares, and refuses to run without a CUDA device (there is
no CPU fallback to silently fall into)."""
import ctypes
import os
import re

import pytest

import hgaprec_b200 as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hpf_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"HPF_API\s+[\w\s\*]+?\b(hpf_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("hpf_create", "hpf_destroy", "hpf_set_ratings_csr", "hpf_set_state", "hpf_get_state",
              "hpf_iterate", "hpf_heldout_loglik", "hpf_elbo", "hpf_topn", "hpf_item_ranks", "hpf_partition_users", "hpf_comm_init",
              "hpf_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    assert os.path.exists(H.LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(H.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "libhpf_b200.so does not export %s" % s


def test_config_struct_layout_matches_header():
    # the python mirror must have one field per header field, same order
    src = open(HEADER).read()
    body = src[src.index("typedef struct hpf_config {"):src.index("} hpf_config;")]
    names = []
    for line in body.splitlines()[1:]:
        line = line.split("/*")[0]
        m = re.match(r"\s*(uint32_t|int32_t|uint64_t|double)\s+([^;]+);", line)
        if m:
            names += [re.sub(r"\[.*\]", "", n).strip() for n in m.group(2).split(",")]
    from hgaprec_b200.capi import _Config, MAX_DEVICES
    assert [f[0] for f in _Config._fields_] == names
    assert "#define HPF_MAX_DEVICES %d" % MAX_DEVICES in src and "#define HPF_ABI_VERSION %d" % H.capi.ABI_VERSION in src


def test_stats_struct_layout_matches_header():
    src = open(HEADER).read()
    body = src[src.index("typedef struct hpf_stats {"):src.index("} hpf_stats;")]
    fields = []
    for line in body.splitlines()[1:]:
        m = re.match(r"\s*(uint32_t|uint64_t|float)\s+(\w+);", line.split("/*")[0])
        if m:
            fields.append((m.group(2), m.group(1)))
    from hgaprec_b200.capi import Stats
    ctype = {"uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64, "float": ctypes.c_float}
    assert [(f[0], f[1]) for f in Stats._fields_] == [(n, ctype[t]) for n, t in fields]


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(H.HpfError) as ei:
        H.Engine(8, 8, 4)
    assert ei.value.code == -5  # HPF_ENODEVICE
    assert "no CUDA device" in str(ei.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "hgaprec_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("hpf_oracle_not", ""), "%s mentions the oracle" % fn


def test_bench_and_entry_scripts_have_no_undefined_names():
    """bench.py only runs on the GPU box; catch typos (a name used but never bound) here."""
    import ast
    import builtins
    for script in ("bench.py", "__graft_entry__.py"):
        tree = ast.parse(open(os.path.join(ROOT, script)).read())
        bound = set(dir(builtins)) | {"__file__", "__name__"}
        for node in ast.walk(tree):
            if isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
                bound.add(node.id)
            elif isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                bound.add(node.name)
                if isinstance(node, ast.FunctionDef):
                    a = node.args
                    bound.update(x.arg for x in a.args + a.kwonlyargs + a.posonlyargs)
                    bound.update(x.arg for x in (a.vararg, a.kwarg) if x)
            elif isinstance(node, ast.Lambda):
                bound.update(x.arg for x in node.args.args)
            elif isinstance(node, (ast.Import, ast.ImportFrom)):
                bound.update((x.asname or x.name).split(".")[0] for x in node.names)
            elif isinstance(node, ast.ExceptHandler) and node.name:
                bound.add(node.name)
        used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
        assert not sorted(used - bound), (script, sorted(used - bound))
