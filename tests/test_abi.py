"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/dladmm.h declares,
its ctypes structs match the header's field lists, and argument validation fails loudly."""
import ctypes as C
import os
import re

import pytest

import dladmm_b200 as dl
from dladmm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dladmm.h")


def _declared():
    src = open(HEADER).read()
    return re.findall(r"DLADMM_API\s+[\w\s\*]+?\b(dladmm_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert set(names) == set(_lib.EXPORTS) and len(names) >= 8
    for n in names:
        assert getattr(lib, n) is not None


def _struct_fields(name):
    src = open(HEADER).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    return fields


@pytest.mark.parametrize("cname,cls", [("dladmm_bparam", _lib.BParam), ("dladmm_layer", _lib.Layer),
                                       ("dladmm_problem", _lib.Problem), ("dladmm_cotangents", _lib.Cotangents),
                                       ("dladmm_caps", _lib.Caps), ("dladmm_gen_desc", _lib.GenDesc),
                                       ("dladmm_sg_pair", _lib.SgPair), ("dladmm_metrics", _lib.Metrics)])
def test_ctypes_structs_mirror_header(cname, cls):
    assert _struct_fields(cname) == [f[0] for f in cls._fields_]


def test_invalid_arguments_are_rejected_with_messages():
    lib = _lib.load()
    p = _lib.Problem()
    p.abi_version = 999
    rc = lib.dladmm_forward(C.byref(p), None)
    assert rc == -1 and b"ABI version" in lib.dladmm_last_error()
    p.abi_version = _lib.ABI_VERSION
    p.family, p.m, p.d, p.K = 7, 4, 4, 1
    assert lib.dladmm_forward(C.byref(p), None) == -1 and b"family" in lib.dladmm_last_error()
    p.family = _lib.FAMILY_B
    assert lib.dladmm_forward(C.byref(p), None) == -1 and b"layers" in lib.dladmm_last_error()
    with pytest.raises(RuntimeError, match="libdladmm error"):
        _lib.check(lib.dladmm_forward(C.byref(p), None))
    assert lib.dladmm_workspace_bytes(None, 0) == 0
    assert lib.dladmm_gen_workspace_bytes(250, 500) >= 250 * 512 * 4


def test_library_is_in_tree():
    assert os.path.dirname(dl.library_path()).endswith(os.path.join("d-ladmm_b200", "csrc"))


def test_ctypes_struct_sizes_and_offsets_match_the_c_compiler(tmp_path):
    """The header is plain C: compile a probe with gcc and compare sizeof / offsetof of every struct with the ctypes
    mirrors the Python host uses (field names alone would not catch a type or padding drift)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = [("dladmm_bparam", _lib.BParam), ("dladmm_layer", _lib.Layer), ("dladmm_problem", _lib.Problem),
               ("dladmm_cotangents", _lib.Cotangents), ("dladmm_caps", _lib.Caps), ("dladmm_gen_desc", _lib.GenDesc),
               ("dladmm_sg_pair", _lib.SgPair), ("dladmm_metrics", _lib.Metrics)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dladmm.h"', "int main(void) {",
             '  printf("abi %d\\n", DLADMM_ABI_VERSION);']
    for cname, cls in structs:
        lines.append('  printf("%s size %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, f[0]))
    lines += ["  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["abi"]) == _lib.ABI_VERSION
    for cname, cls in structs:
        assert int(out[cname + " size"]) == C.sizeof(cls), cname
        for f in cls._fields_:
            assert int(out["%s.%s" % (cname, f[0])]) == getattr(cls, f[0]).offset, (cname, f[0])
