"""ABI guard (CPU, gcc only): the ctypes mirror in ttl_b200/_lib.py lays every struct of include/ttl_b200.h out exactly as
a C compiler does, carries the same enum values, and binds every function with the header's argument count.  A drift here
would not fail loudly at run time (ctypes passes whatever it is told), so it is pinned here."""
import ctypes as C
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ttl_b200.h")

# C struct -> ctypes mirror
STRUCTS = {"ttl_config": "TtlConfig", "ttl_hparams": "TtlHparams", "ttl_outputs": "TtlOutputs",
           "ttl_view_spec": "TtlViewSpec", "ttl_text_config": "TtlTextConfig", "ttl_gemm_record": "TtlGemmRecord",
           "ttl_deyo_options": "TtlDeyoOptions"}
# C enumerator prefix -> prefix of the Python constant
ENUM_PREFIXES = {"TTL_W_": "W_", "TTL_LORA_": "LORA_", "TTL_HEAD_": "HEAD_", "TTL_PRECISION_": "PRECISION_",
                 "TTL_VIEW_": "VIEW_", "TTL_TW_": "TW_", "TTL_AUG_": "AUG_"}


def _header_text():
    return re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)


def _struct_fields(text, name):
    body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        # "const int32_t* a" / "int32_t a, b": the field name is the last identifier of every comma-separated part
        fields += [part.replace("*", " ").split()[-1] for part in decl.split(",")]
    return fields


@pytest.fixture(scope="module")
def c_layout(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    text = _header_text()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ttl_b200.h"', 'int main(void) {', 'printf("{");']
    for s in STRUCTS:
        lines.append('printf("\\"sizeof %s\\": %%zu, ", sizeof(%s));' % (s, s))
        for f in _struct_fields(text, s):
            lines.append('printf("\\"%s.%s\\": %%zu, ", offsetof(%s, %s));' % (s, f, s, f))
    enums = sorted(set(re.findall(r"\b(TTL_[A-Z0-9_]+)\s*=\s*-?\d+", text)))
    for e in enums:
        lines.append('printf("\\"%s\\": %%d, ", (int)%s);' % (e, e))
    lines += ['printf("\\"end\\": 0}\\n");', 'return 0; }']
    d = tmp_path_factory.mktemp("abi")
    src, exe = d / "probe.c", d / "probe"
    src.write_text("\n".join(lines))
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    return json.loads(subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout)


def test_header_is_plain_c99():
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_struct_layouts_match_ctypes(c_layout):
    from ttl_b200 import _lib
    text = _header_text()
    for cname, pyname in STRUCTS.items():
        st = getattr(_lib, pyname)
        assert C.sizeof(st) == c_layout["sizeof " + cname], cname
        c_fields = _struct_fields(text, cname)
        assert [f for f, _ in st._fields_] == c_fields, cname         # same names, same order
        for f in c_fields:
            assert getattr(st, f).offset == c_layout["%s.%s" % (cname, f)], (cname, f)


def test_enum_values_match_python_constants(c_layout):
    from ttl_b200 import _lib
    checked = 0
    for cname, value in c_layout.items():
        if not cname.startswith("TTL_") or cname.startswith("TTL_E_") or cname == "TTL_OK":
            continue
        prefix = max((p for p in ENUM_PREFIXES if cname.startswith(p)), key=len, default=None)
        assert prefix is not None, cname
        pyname = ENUM_PREFIXES[prefix] + cname[len(prefix):]
        assert getattr(_lib, pyname) == value, (cname, pyname)
        checked += 1
    assert checked >= 40


def test_bound_argument_counts_match_the_header():
    from ttl_b200 import _lib
    text = _header_text()
    protos = re.findall(r"\b(ttl_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text)
    assert len(protos) >= 30
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert name in _lib._SIGS, name
        assert len(_lib._SIGS[name][1]) == n, (name, n, len(_lib._SIGS[name][1]))
