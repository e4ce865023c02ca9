"""CPU checks on the Go adapter sources (no Go toolchain exists in the image, so they are validated statically):
every C symbol they call is declared in include/b2kyber.h with the same number of arguments, every type suite.go names is
defined, and the generators in g1.go / g2.go are the standard ones (g1Generator was all-zero)."""
import os
import re

from oracle import bls12381 as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "go")


def _go_files():
    for root, _, files in os.walk(GO):
        for f in files:
            if f.endswith(".go"):
                yield os.path.join(root, f)


def _split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_every_c_call_matches_a_declaration():
    hdr = open(os.path.join(ROOT, "include", "b2kyber.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"\b(b2k_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)}
    seen = 0
    for path in _go_files():
        src = open(path).read()
        for m in re.finditer(r"C\.(b2k_[a-z0-9_]+)\(", src):
            name = m.group(1)
            assert name in decls, f"{os.path.basename(path)} calls {name}, not declared in include/b2kyber.h"
            # argument list up to the matching parenthesis
            i, depth = m.end(), 1
            while depth:
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            nargs = len(_split_args(src[m.end():i - 1]))
            assert nargs == decls[name], f"{os.path.basename(path)}: {name} called with {nargs} arguments, declared with {decls[name]}"
            seen += 1
    assert seen >= 20


def test_every_type_the_suite_names_is_defined():
    src = "\n".join(open(p).read() for p in _go_files() if os.sep + "bls12381" + os.sep in p)
    for ident in ("G1Elt", "G2Elt", "GTElt", "groupBls", "Suite", "MultiGPU"):
        assert re.search(rf"\btype {ident} struct\b", src), f"type {ident} is not defined"
    for fn in ("newEmptyGT", "NullG1", "NullG2", "NewGroupG1", "NewGroupG2", "NewGroupGT", "NewScalar", "with", "acquire"):
        assert re.search(rf"\bfunc {fn}\(", src), f"func {fn} is not defined"
    # the 17 methods of kyber.Point (group.go:84-131) on every point type
    methods = ["Equal", "Null", "Base", "Pick", "Set", "Clone", "EmbedLen", "Embed", "Data", "Add", "Sub", "Neg", "Mul",
               "MarshalBinary", "UnmarshalBinary", "MarshalTo", "UnmarshalFrom", "MarshalSize", "String"]
    for t in ("G1Elt", "G2Elt", "GTElt"):
        for m in methods:
            assert re.search(rf"func \(k \*{t}\) {m}\(", src), f"{t}.{m} missing"
    assert "omitted for brevity" not in src


def test_generators_are_the_standard_ones():
    def arr(path, name):
        src = open(os.path.join(GO, "pairing", "bls12381", "b200", path)).read()
        body = re.search(rf"var {name} = \[\d+\]byte\{{(.*?)\n\}}", src, flags=re.S).group(1)
        return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})", body))
    assert arr("g1.go", "g1Generator") == o.g1_to_affine_bytes(o.G1)
    assert arr("g2.go", "g2Generator") == o.g2_to_affine_bytes(o.G2)
