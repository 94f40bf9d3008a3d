"""The C-ABI library loads and exports every symbol include/rzb200.h declares; struct layouts match the header."""
import ctypes
import os
import re

import numpy as np
import pytest

from rayzath_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rzb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rzb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "librzb200.so does not export " + name


def test_binding_covers_header():
    assert sorted(capi.SYMBOLS) == declared_symbols()


def test_flag_constants_match_the_header():
    """The Python binding's FLAG_* / SCENE_* constants are the header's enumerators."""
    text = open(os.path.join(ROOT, "include", "rzb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = {name: int(value, 0) for name, value in re.findall(r"\b(RZB_[A-Z_0-9]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", text)}
    checked = 0
    for py_name, value in vars(capi).items():
        if py_name.startswith("FLAG_") or py_name.startswith("SCENE_"):
            c_name = "RZB_" + py_name
            assert c_name in enums, c_name
            assert enums[c_name] == value, (c_name, enums[c_name], value)
            checked += 1
    assert checked >= 5 and enums["RZB_FLAG_SERIAL_STAGES"] == 8


def test_abi_version():
    assert capi.lib().rzb_abi_version() == 3


def test_struct_sizes():
    for name, (dtype, size) in capi.EXPECTED_SIZES.items():
        assert dtype.itemsize == size, name
    assert ctypes.sizeof(capi.SceneStruct) == 240  # static_assert(sizeof(rzb_scene) == 240) in csrc/rzb_api.cu


def test_no_device_fails_loudly():
    """Without a CUDA device a context cannot be created: there is no CPU fallback."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.RzbError):
        capi.Context(0)


def test_product_does_not_touch_oracle():
    """Nothing under rayzath_b200/ may import, link or execute oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "rayzath_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and f != "Makefile":
                continue
            text = open(os.path.join(dirpath, f), errors="replace").read()
            for line in text.splitlines():
                code = line.split("//")[0].split("#")[0] if not f.endswith(".py") and f != "Makefile" else line.split("#")[0]
                assert "rz_oracle" not in code and "liboracle" not in code and "rz_ref_tool" not in code, (f, line)
