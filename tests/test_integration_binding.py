"""INTEGRATION.md's Rust `extern "C"` block against include/voidray_cuda.h: every function the binding declares exists
in the header with the same number of parameters and matching scalar / pointer kinds, and the #[repr(C)] structs have
the header's fields in the header's order (Rust is not installed here, so this is the check the binding gets)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _header():
    text = _strip_comments(open(os.path.join(ROOT, "include", "voidray_cuda.h")).read())
    funcs = {}
    for m in re.finditer(r"\b(?:int32_t|uint32_t|const char\s*\*)\s+(vr_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")] if m.group(2).strip() not in ("", "void") else []
        funcs[m.group(1)] = args
    structs = {}
    for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\w+\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1] if " " in decl else decl
            for n in names.split(","):
                fields.append(re.sub(r"[\s\*]|\[.*\]", "", n.split()[-1]))
        structs[m.group(1)] = fields
    return funcs, structs


def _rust():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = _strip_comments(md[md.index("```rust"):md.index("```", md.index("```rust") + 7)])
    funcs = {}
    for m in re.finditer(r"pub fn (vr_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        funcs[m.group(1)] = args
    structs = {}
    for m in re.finditer(r"pub struct (vr_\w+)\s*\{(.*?)\}", block, flags=re.S):
        structs[m.group(1)] = re.findall(r"pub (\w+)\s*:", m.group(2))
    return funcs, structs


def _kind_c(arg):
    if "*" in arg or "[" in arg:
        return "ptr"
    if re.search(r"\bfloat\b", arg):
        return "f32"
    if re.search(r"\bdouble\b", arg):
        return "f64"
    if re.search(r"\buint64_t\b", arg):
        return "u64"
    if re.search(r"\buint32_t\b", arg):
        return "u32"
    if re.search(r"\bint32_t\b", arg):
        return "i32"
    return "other"


def _kind_rust(arg):
    ty = arg.split(":", 1)[1].strip()
    if ty.startswith("*"):
        return "ptr"
    return ty


def test_rust_binding_matches_the_header():
    c_funcs, c_structs = _header()
    r_funcs, r_structs = _rust()
    assert len(r_funcs) >= 43 and len(c_funcs) >= 50
    for name, r_args in r_funcs.items():
        assert name in c_funcs, f"{name} is not in the header"
        c_args = c_funcs[name]
        assert len(c_args) == len(r_args), f"{name}: {len(c_args)} parameters in the header, {len(r_args)} in the binding"
        for ca, ra in zip(c_args, r_args):
            assert _kind_c(ca) == _kind_rust(ra), f"{name}: `{ca}` vs `{ra}`"
    # the hot path's entry points are all bound
    for name in ("vr_scene_commit", "vr_render_begin", "vr_render_accumulate", "vr_render_cancel", "vr_render_stats",
                 "vr_render_read_accum", "vr_render_resolve", "vr_render_end"):
        assert name in r_funcs
    # everything but the gate functions is bound
    assert all(n.startswith("vr_debug_") for n in set(c_funcs) - set(r_funcs)), sorted(set(c_funcs) - set(r_funcs))
    for name in ("vr_material_desc", "vr_render_settings", "vr_stats", "vr_obj_mesh", "vr_scene_info"):
        assert name in c_structs and name in r_structs
        assert r_structs[name] == c_structs[name], (name, r_structs[name], c_structs[name])
