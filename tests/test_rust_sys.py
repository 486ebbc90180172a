"""The Rust `-sys` crate (bindings/rust/vectorx-b200-sys) is generated from include/vectorx_b200.h by
tools/gen_rust_sys.py: the committed lib.rs must be the generator's current output, declare every exported symbol with
the arity of the ctypes table the Python binding uses, and escape Rust keywords.  (No cargo in this image: the crate is
checked as text; INTEGRATION.md describes how the plonky2 fork uses it.)"""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_committed_lib_rs_is_the_generator_output():
    mod = _gen()
    src, names = mod.generate()
    assert open(mod.OUT).read() == src, "run `python tools/gen_rust_sys.py` after changing include/vectorx_b200.h"
    assert len(names) == len(set(names)) and len(names) >= 60


def test_declarations_cover_the_abi_with_the_same_arity():
    from vectorx_b200._lib import SIGNATURES
    src = open(_gen().OUT).read()
    decl = dict(re.findall(r"pub fn (vx_[a-z0-9_]+)\(([^)]*)\)", src))
    assert set(decl) == set(SIGNATURES)
    for name, (res, args) in SIGNATURES.items():
        n_rust = 0 if not decl[name].strip() else decl[name].count(":")
        assert n_rust == len(args), name
        has_ret = re.search(rf"pub fn {name}\([^)]*\) -> ", src) is not None
        assert has_ret == (res is not None), name
    assert "r#in:" in src and " in:" not in src                  # vx_ntt's `in` parameter
    for handle in ("vx_ctx", "vx_batch", "vx_tree", "vx_fri", "vx_shard_group"):
        assert f"pub struct {handle} " in src
