"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/scgib.h declares,
and its host-only entry points (layout, workspace sizing, argument validation) behave.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from scgib_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("scgib_build", os.path.join(ROOT, "s-cgib_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _lib


def test_library_exports_every_declared_symbol():
    L = _lib()
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "scgib.h")).read()
    declared = set(re.findall(r"SCGIB_API\s+[\w\s\*]+?\b(scgib_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.scgib_version() == 100


def test_param_layout_matches_reference_parameter_count():
    L = _lib()
    lib = L.load()
    from scgib_b200.engine import param_names, param_shapes
    for F, layers in ((9, 4), (11, 4), (9, 5)):
        d = L.Dims(F, 32, 64, layers)
        n = lib.scgib_param_slots(ctypes.byref(d))
        off = (ctypes.c_int64 * n)()
        sz = (ctypes.c_int64 * n)()
        total = lib.scgib_param_layout(ctypes.byref(d), off, sz)
        shapes = param_shapes(F, layers)
        assert n == len(param_names(layers)) == len(shapes)
        for s, shp in zip(sz, shapes):
            k = 1
            for v in shp:
                k *= v
            assert s == k
        assert all(o % 4 == 0 for o in off) and total >= sum(sz)
    # used parameters of the reference's default configuration: 80 674
    d = L.Dims(9, 32, 64, 4)
    n = lib.scgib_param_slots(ctypes.byref(d))
    sz = (ctypes.c_int64 * n)()
    lib.scgib_param_layout(ctypes.byref(d), None, sz)
    assert sum(sz) == 80674


def test_argument_validation_without_gpu():
    L = _lib()
    lib = L.load()
    bad = L.Dims(9, 32, 48, 4)
    assert lib.scgib_param_slots(ctypes.byref(bad)) == -2
    assert lib.scgib_pretrain_workspace_bytes(ctypes.byref(bad), 1, 2, 2, 2, 2) == 0
    good = L.Dims(9, 32, 64, 4)
    small = lib.scgib_pretrain_workspace_bytes(ctypes.byref(good), 128, 1920, 4100, 6000, 8200)
    big = lib.scgib_pretrain_workspace_bytes(ctypes.byref(good), 4096, 61000, 131000, 193000, 263000)
    assert 0 < small < big < 4 << 30
    assert lib.scgib_pretrain_forward_f32(ctypes.byref(good), None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.scgib_ego_count(None, None, 4, 1, None, None, None, None, 0, None) == -1
    assert lib.scgib_adam_step_f32(None, None, None, None, 1, 1, 0.1, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert b"NULL" in lib.scgib_error_string(-1)


def test_batch_struct_matches_header_and_integration_doc():
    """The ctypes mirror of ScgibBatch (s-cgib_b200/_lib.py) and the binding INTEGRATION.md shows a maintainer have the
    library's struct size, and a struct of another size is refused (SCGIB_E_ABI) before anything is dereferenced."""
    L = _lib()
    lib = L.load()
    assert ctypes.sizeof(L.Batch) == lib.scgib_batch_abi_size()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class Dims\(ctypes.Structure\).*?\[\(\"eval_mode\".*?\)\]\n", doc, re.S)
    assert m, "INTEGRATION.md no longer shows the ctypes structs"
    ns = {"ctypes": ctypes}
    exec(m.group(0), ns)
    assert ctypes.sizeof(ns["Batch"]) == lib.scgib_batch_abi_size()
    assert [f[0] for f in ns["Batch"]._fields_] == [f[0] for f in L.Batch._fields_]
    assert [f[0] for f in ns["Dims"]._fields_] == [f[0] for f in L.Dims._fields_]
    good = L.Dims(9, 32, 64, 4)
    b = L.Batch()
    b.struct_size = ctypes.sizeof(L.Batch) - 8          # e.g. a binding without eval_mode / recon_logm_steps (round 1 doc)
    one = ctypes.c_float(0)
    rc = lib.scgib_pretrain_forward_f32(ctypes.byref(good), ctypes.byref(one), None, ctypes.byref(b), ctypes.byref(one), None, None,
                                        None, None, ctypes.byref(one), 0, None)
    assert rc == -6 and b"struct_size" in lib.scgib_error_string(-6)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    L = _lib()
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        L.load()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "s-cgib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
