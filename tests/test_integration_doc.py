"""INTEGRATION.md shows the Rust binding a maintainer of the reference would add.  It is not compiled here (no Rust toolchain),
so this test keeps it from drifting: its `#[repr(C)]` parameter struct lists the fields of include/asph.h `asph_params` in the
same order with matching widths, and its `extern "C"` block names only functions the header declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "asph.h")) as f:
        return f.read()


def _doc():
    with open(os.path.join(ROOT, "INTEGRATION.md")) as f:
        return f.read()


def _c_params_fields(h):
    body = re.search(r"typedef struct asph_params \{(.*?)\} asph_params;", h, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        for name in names.split(","):
            name = name.strip()
            m = re.match(r"(\w+)\[(\d+)\]", name)
            fields.append((m.group(1), ctype, int(m.group(2))) if m else (name, ctype, 1))
    return fields


def _rust_params_fields(doc):
    body = re.search(r"pub struct AsphParams \{(.*?)\n\}", doc, re.S).group(1)
    body = re.sub(r"//[^\n]*", "", body)
    fields = []
    for m in re.finditer(r"pub (\w+): (\[(\w+); (\d+)\]|\w+)", body):
        fields.append((m.group(1), m.group(3) or m.group(2), int(m.group(4) or 1)))
    return fields


WIDTH = {"double": "c_double", "int32_t": "i32", "int64_t": "i64", "float": "c_float", "uint64_t": "u64"}


def test_rust_params_struct_mirrors_the_header():
    c, r = _c_params_fields(_header()), _rust_params_fields(_doc())
    assert len(c) > 50
    assert [f[0] for f in c] == [f[0] for f in r]
    for (name, ctype, n), (_, rtype, rn) in zip(c, r):
        assert WIDTH[ctype] == rtype and n == rn, (name, ctype, rtype, n, rn)


def test_rust_extern_block_names_declared_functions_only():
    declared = set(re.findall(r"\b(asph_\w+)\s*\(", _header()))
    block = re.search(r'extern "C" \{(.*?)\n\}', _doc(), re.S).group(1)
    named = re.findall(r"pub fn (asph_\w+)", block)
    assert len(named) >= 12
    assert set(named) <= declared, sorted(set(named) - declared)
    for must in ("asph_create", "asph_destroy", "asph_step", "asph_step_physics", "asph_step_adaptivity", "asph_get_field", "asph_last_error"):
        assert must in named
