"""The Rust shim (rust/) cannot be compiled here (no rustc); these checks keep it from going stale anyway:
rust/src/ffi.rs is GENERATED from include/tr_prover.h (rust/gen_bindings.py) and must match a fresh generation, every function
it declares must be exported by the built libtrp.so with the header's argument count, and rust/build.rs must compile exactly
the units tiny-ram-halo2_b200/csrc/Makefile links."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust"))
import gen_bindings as G


def test_ffi_rs_is_a_fresh_generation_of_the_header():
    with open(G.OUT) as f:
        assert f.read() == G.render(), "rust/src/ffi.rs is stale: run python rust/gen_bindings.py"


def test_header_parser_sees_every_declaration():
    decls = {name: (args, ret) for name, args, ret in G.parse_header()}
    text = G.strip_comments(open(G.HEADER).read())
    # every trp_* identifier followed by '(' in the header is a declaration the generator emitted
    assert set(re.findall(r"\b(trp_[a-z0-9_]+)\s*\(", text)) == set(decls)
    # spot checks, argument counts included (the round-1 shim had dropped n_regs from trp_quotient_eval)
    assert len(decls["trp_quotient_eval"][0]) == 9 and ("n_regs", "c_uint") in decls["trp_quotient_eval"][0]
    assert len(decls["trp_dev_quotient_eval"][0]) == 10
    assert decls["trp_msm"][0][-1] == ("out_jacobian", "*mut u64") and decls["trp_msm"][1] == "c_int"
    assert decls["trp_ctx_destroy"][1] is None and decls["trp_last_error"][1] == "*const c_char"
    assert decls["trp_dev_permutation_product"][0][1] == ("d_values", "*const *const u64")
    assert decls["trp_ctx_set_stream"][0][1] == ("cuda_stream", "*mut c_void")
    assert decls["trp_bases_load"][0][-1] == ("out", "*mut *mut trp_bases")


def test_every_declared_symbol_is_exported_by_libtrp():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "tiny-ram-halo2_b200", "libtrp.so"))
    missing = [name for name, _, _ in G.parse_header() if not hasattr(lib, name)]
    assert not missing, missing


def test_lib_rs_uses_the_generated_block_and_calls_match():
    src = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    assert "mod ffi;" in src and 'extern "C"' not in src, "lib.rs must not carry hand-written extern declarations"
    decls = {name: args for name, args, _ in G.parse_header()}
    for name, body in re.findall(r"\b(trp_[a-z0-9_]+)\(([^;]*?)\)\s*[};)]", src):
        if name in decls:        # a call: the number of top-level arguments must equal the header's
            depth, count = 0, 1 if body.strip() else 0
            for ch in body:
                depth += ch in "([{"; depth -= ch in ")]}"
                count += ch == "," and depth == 0
            assert count == len(decls[name]), (name, count, len(decls[name]))


def test_build_rs_compiles_the_units_the_makefile_links():
    csrc = os.path.join(ROOT, "tiny-ram-halo2_b200", "csrc")
    mk = open(os.path.join(csrc, "Makefile")).read()
    objs = re.search(r"^OBJS := (.*)$", mk, flags=re.M).group(1).split()
    assert len(objs) >= 12
    for o in objs:
        assert os.path.exists(os.path.join(csrc, o[:-2] + ".cu")), o
    b = open(os.path.join(ROOT, "rust", "build.rs")).read()
    assert 'starts_with("OBJS :=")' in b and "capi" not in b, "build.rs must take the unit list from csrc/Makefile"
    # every extern "C" definition in csrc belongs to a listed unit
    listed = {o[:-2] + ".cu" for o in objs}
    for fn in os.listdir(csrc):
        if fn.endswith(".cu") and 'extern "C"' in open(os.path.join(csrc, fn)).read():
            assert fn in listed, fn


def test_generator_check_mode():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "rust", "gen_bindings.py"), "--check"]).returncode == 0
