"""The C-ABI library loads and exports every symbol include/gorp_cuda.h declares; host-only entry points work
without a GPU; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gorp_b200 import _ffi
from gorp_b200.api import Blob, DefinitionReader, GorpCudaError
from tests import reference_vectors as V


def declared_functions():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "gorp_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gorp_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(_ffi.lib, n), n
    assert sorted(_ffi.SYMBOLS) == names
    assert _ffi.lib.gorp_abi_version() == 2


def test_blob_roundtrip_and_corruption():
    b = Blob.from_definition(V.README_DEF)
    raw = bytearray(b.bytes())
    assert raw[:8] == b"GORPDFA1"
    assert b.info() == (31, 34, 3)
    raw[5000] ^= 0xFF
    bad = bytes(raw)
    bi = _ffi.BlobInfo()
    rc = _ffi.lib.gorp_blob_get_info(bad, len(bad), C.byref(bi))
    assert rc == _ffi.GORP_E_BLOB and "checksum" in _ffi.last_error()
    rc = _ffi.lib.gorp_blob_get_info(bad[:100], 100, C.byref(bi))
    assert rc == _ffi.GORP_E_BLOB


def test_no_cpu_fallback():
    if _ffi.lib.gorp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = DefinitionReader.reader(V.README_DEF).read()
    with pytest.raises(GorpCudaError) as ei:
        g.extract("[1]: GET 2ms /a")
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(GorpCudaError):
        g.extract_batch_text(np.zeros(4, dtype=np.uint16))


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "gorp_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in src.replace("oracle layout", ""), os.path.join(dirpath, f)
