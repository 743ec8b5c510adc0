"""The C-ABI library loads and exports every symbol include/mlgi.h declares; without a GPU
the product refuses to run (no CPU fallback). No compute calls here."""
import ctypes as C
import os
import re

import pytest

import mlgi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mlgi.h")).read()
    return sorted(set(re.findall(r"MLGI_API\s+[\w\s\*]+?\b(mlgi_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    syms = _declared_symbols()
    assert len(syms) >= 20
    lib = mlgi.load()
    for s in syms:
        assert hasattr(lib, s), "libmlgi.so does not export %s" % s
    assert sorted(mlgi.EXPORTS) == syms, "python binding list and header drifted apart"


def test_struct_sizes_match_header():
    # field counts of the POD structs in include/mlgi.h
    assert C.sizeof(mlgi.Config) == 12 * 4
    assert C.sizeof(mlgi.Params) == 16 * 4
    assert C.sizeof(mlgi.Camera) == 32 * 4
    assert C.sizeof(mlgi.Timing) == 32


def test_default_params_mirror_gistage_defaults():
    p = mlgi.default_params(1024)   # GIStage.cpp:215-232
    assert p.light_intensity == 5.0 and p.gi_intensity_factor == 3000.0
    assert abs(p.vpl_clamping_value - 0.001) < 1e-9
    assert (p.vpl_start_index, p.vpl_end_index) == (0, 1024)
    assert p.tess_level_factor == 2.0 and p.use_push_pull == 1 and p.enable_shadowing == 1
    assert p.use_interleaving == 1 and p.shuffle_lights == 1 and p.scale_isms == 0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is checked on the CPU container")
    with pytest.raises(mlgi.MlgiError) as e:
        mlgi.Context(64, 64, 64, 128, 128, 100, 1000)
    assert e.value.code == mlgi.E_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "many-lights-gi_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "glsl_compat" not in txt, os.path.join(dp, f)
