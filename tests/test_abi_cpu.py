"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every
symbol that include/csts_b200.h declares; the ctypes binding covers the same set; product code never
imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    text = open(os.path.join(ROOT, "include", "csts_b200.h")).read()
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def declared_symbols():
    return sorted(set(re.findall(r"\b(csts_[a-z0-9_]+)\s*\(", _header())))


def test_header_declares_the_abi():
    syms = declared_symbols()
    assert "csts_gemm" in syms and "csts_dwconv" in syms and "csts_kldiv_frame_softmax" in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    from csts_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.csts_version() >= 100


def test_ctypes_binding_matches_header():
    from csts_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"csts_launch_count", "csts_gemm_backend", "csts_gemm_plan", "csts_mt_chunk_elems"}
    assert bound == set(declared_symbols())
    _lib.load()


def test_struct_layouts_match_header():
    """Field order of the ctypes structures = field order of the C structs."""
    from csts_b200 import _lib
    text = _header()
    for cname, cls in (("csts_gemm_args", _lib.GemmArgs), ("csts_pool_args", _lib.PoolArgs), ("csts_wgrad_args", _lib.WgradArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(?:const\s+)?\w+\s*\*?\s*(.*)$", decl, flags=re.S)
            fields += [n.strip().lstrip("*").strip() for n in m.group(1).split(",")]
        py = ["in" if f[0] == "inp" else f[0] for f in cls._fields_]
        assert py == fields, (cname, py, fields)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from csts_b200 import kernels as K
    with pytest.raises(RuntimeError):
        K.add_f32(torch.zeros(4), torch.zeros(4))


def test_product_code_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "csts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+(oracle|csts_oracle|ref_shim|ref_train)", src, flags=re.M):
                    bad.append(f)
    assert not bad, bad
