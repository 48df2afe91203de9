"""The C-ABI library loads and exports every symbol include/anatomix_b200.h
declares (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from anatomix_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "anatomix_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(anx_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_descriptor_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "anatomix_b200.h")).read()
    body = re.search(r"typedef struct anx_unet_desc \{(.*?)\} anx_unet_desc;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:uint32_t|int32_t|float)\s+([a-z_]+);", body)
    assert fields == [f[0] for f in _lib.UnetDesc._fields_]
    assert ctypes.sizeof(_lib.UnetDesc) == 4 * len(fields)


def test_pure_host_calls():
    lib = _lib.load()
    assert lib.anx_version() >= 100
    assert lib.anx_status_string(2) == b"unsupported input shape"
    assert lib.anx_engine_num_convs(None) == -1
    assert lib.anx_engine_workspace_bytes(None, 1, 32, 32, 32) == 0
