"""The Fortran side of the boundary cannot be compiled in this image (no Fortran compiler), so its agreement with the
C header is checked textually: every `bind(C, name=...)` interface of fortran/mlegs_b200_c.f90 names an entry that
include/mlegs_b200.h declares, with the same number of arguments and the same by-value / by-reference passing; the two
`bind(C)` derived types list the members of the C structs in the same order and with matching kinds; every C entry the
replacement submodule calls has an interface; and (when the reference tree is present) the submodule gives a body to
every `module subroutine / function` that modules/mlegs_scalar.f90 declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "mlegs_b200.h")
F_IFACE = os.path.join(ROOT, "fortran", "mlegs_b200_c.f90")
F_SHIM = os.path.join(ROOT, "fortran", "mlegs_scalar_ops_b200.f90")
REF_IFACE = "/root/reference/src/modules/mlegs_scalar.f90"


def _c_text():
    txt = open(HDR).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def _c_prototypes():
    """name -> list of (by_value, base_type) per argument."""
    protos = {}
    for m in re.finditer(r"\b(mlegs_b200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", _c_text()):
        name, args = m.group(1), " ".join(m.group(2).split())
        out = []
        if args not in ("void", ""):
            for a in args.split(","):
                a = a.strip()
                by_value = "*" not in a and "[" not in a
                base = re.sub(r"\bconst\b", "", a).replace("*", " ").split()
                out.append((by_value, " ".join(base[:-1])))
        protos[name] = out
    return protos


def _fortran_lines(path):
    """Source lines with comments stripped and `&` continuations joined."""
    out, cur = [], ""
    for raw in open(path):
        line = raw.split("!")[0].rstrip() if "'" not in raw.split("!")[0] or raw.count("'") % 2 == 0 else raw.rstrip()
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].strip()
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        out.append(cur + line)
        cur = ""
    return out


def _fortran_interfaces():
    """name -> (dummy argument list, {dummy: (by_value, declared type)}) for every bind(C) function."""
    lines = _fortran_lines(F_IFACE)
    res = {}
    i = 0
    while i < len(lines):
        m = re.match(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)\s*result\((\w+)\)", lines[i], re.I)
        if not m:
            i += 1
            continue
        fname, args, cname, resname = m.group(1), m.group(2), m.group(3), m.group(4)
        assert fname == cname, f"{fname} binds to a different C name {cname}"
        dummies = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        i += 1
        while not re.match(r"end function", lines[i], re.I):
            d = re.match(r"(.+?)::\s*(.+)", lines[i])
            if d:
                spec, names = d.group(1), d.group(2)
                for n in re.split(r",\s*(?![^()]*\))", names):
                    n = re.sub(r"\(.*\)", "", n).strip()
                    decl[n] = ("value" in spec.lower().replace(" ", "").split(","), spec.split(",")[0].strip())
            i += 1
        assert resname in decl, f"{cname}: result {resname} not declared"
        res[cname] = (dummies, decl)
    return res


F2C = {"integer(c_int)": {"int"}, "real(c_double)": {"double"}, "integer(c_size_t)": {"size_t"},
       "integer(c_long_long)": {"long long", "unsigned long long"}, "type(c_mlegs_field)": {"mlegs_field"},
       "type(c_mlegs_params)": {"mlegs_params"}, "character(kind=c_char)": {"char"},
       "complex(c_double_complex)": {"void", "double"}, "integer(c_signed_char)": {"unsigned char", "char"},
       "integer(c_int64_t)": {"long long", "int64_t"},
       # type(c_ptr), value == void* (or any object pointer); type(c_ptr) by reference == void** / T**
       "type(c_ptr)": None}


def test_every_fortran_interface_matches_a_header_prototype():
    protos = _c_prototypes()
    ifaces = _fortran_interfaces()
    assert len(ifaces) >= 55
    for name, (dummies, decl) in ifaces.items():
        assert name in protos, f"{name}: interface in fortran/mlegs_b200_c.f90 but not declared in the header"
        cargs = protos[name]
        assert len(dummies) == len(cargs), f"{name}: {len(dummies)} Fortran dummies vs {len(cargs)} C parameters"
        for d, (c_by_value, c_type) in zip(dummies, cargs):
            assert d in decl, f"{name}: dummy {d} has no declaration"
            f_by_value, f_type = decl[d]
            f_type = f_type.replace(" ", "").lower()
            key = next((k for k in F2C if k.replace(" ", "") == f_type), None)
            assert key is not None, f"{name}: dummy {d} has unmapped type {f_type}"
            if key == "type(c_ptr)":
                # by value it is the pointer itself; by reference it is a pointer to a pointer
                assert not c_by_value, f"{name}: {d} is a c_ptr but the C parameter is passed by value"
                continue
            assert f_by_value == c_by_value, f"{name}: {d} value/reference mismatch (C: {'value' if c_by_value else 'pointer'})"
            assert c_type in F2C[key] or (not c_by_value and c_type == "void"), f"{name}: {d} is {f_type}, C has {c_type}"


def _c_struct(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), _c_text(), re.S)
    members = []
    for stmt in m.group(1).split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        ty, rest = stmt.split(" ", 1)
        for v in rest.split(","):
            v = v.strip()
            members.append((ty + ("*" if v.startswith("*") else ""), re.sub(r"[\*\s]", "", re.sub(r"\[.*\]", "", v)),
                            int(re.search(r"\[(\d+)\]", v).group(1)) if "[" in v else 1))
    return members


def _f_type(name):
    lines = _fortran_lines(F_IFACE)
    i = next(k for k, l in enumerate(lines) if re.match(r"type,\s*bind\(C\)\s*::\s*%s$" % name, l, re.I))
    members = []
    i += 1
    while not re.match(r"end type", lines[i], re.I):
        spec, names = [t.strip() for t in lines[i].split("::")]
        for n in re.split(r",\s*(?![^()]*\))", names):
            dim = re.search(r"\((\d+)\)", n)
            members.append((spec.replace(" ", "").lower(), re.sub(r"\(.*\)", "", n).strip(), int(dim.group(1)) if dim else 1))
        i += 1
    return members


@pytest.mark.parametrize("cname,fname", [("mlegs_params", "c_mlegs_params"), ("mlegs_field", "c_mlegs_field")])
def test_bind_c_types_mirror_the_structs(cname, fname):
    kinds = {"int": "integer(c_int)", "double": "real(c_double)", "void*": "type(c_ptr)", "char": "character(kind=c_char)"}
    c, f = _c_struct(cname), _f_type(fname)
    assert [(kinds[t], n, d) for t, n, d in c] == f


def test_the_submodule_calls_only_declared_entries():
    ifaces = set(_fortran_interfaces())
    called = set(re.findall(r"\b(mlegs_b200_[a-z0-9_]+)\s*\(", open(F_SHIM).read()))
    assert called, "the shim calls no C entry?"
    assert called <= ifaces, f"no interface for {sorted(called - ifaces)}"


@pytest.mark.skipif(not os.path.exists(REF_IFACE), reason="reference tree not present")
def test_the_submodule_implements_every_module_procedure_of_the_reference_interface():
    want = set(n.lower() for n in re.findall(r"^\s*module\s+(?:recursive\s+)?(?:subroutine|function)\s+(\w+)", open(REF_IFACE).read(),
                                              re.I | re.M))
    have = set(n.lower() for n in re.findall(r"^\s*module procedure\s+(\w+)", open(F_SHIM).read(), re.I | re.M))
    assert len(want) >= 37
    assert want <= have, f"no body for {sorted(want - have)}"
    assert have <= want, f"bodies for procedures the interface does not declare: {sorted(have - want)}"


def _call_sites(text):
    """(name, [top-level argument strings]) of every mlegs_b200_* call in Fortran source text."""
    out = []
    for m in re.finditer(r"\b(mlegs_b200_[a-z0-9_]+)\s*\(", text):
        i, depth, args, cur, quote = m.end(), 1, [], "", None
        while depth and i < len(text):
            ch = text[i]
            if quote:
                quote = None if ch == quote else quote
                cur += ch
            elif ch in "'\"":
                quote = ch
                cur += ch
            elif ch == "(":
                depth += 1
                cur += ch
            elif ch == ")":
                depth -= 1
                if depth:
                    cur += ch
            elif ch == "," and depth == 1:
                args.append(cur.strip())
                cur = ""
            else:
                cur += ch
            i += 1
        if cur.strip():
            args.append(cur.strip())
        out.append((m.group(1), args))
    return out


def test_every_call_in_the_submodule_passes_the_declared_number_of_arguments():
    ifaces = _fortran_interfaces()
    text = "\n".join(_fortran_lines(F_SHIM))
    sites = _call_sites(text)
    assert len(sites) >= 35
    for name, args in sites:
        dummies, _ = ifaces[name]
        assert len(args) == len(dummies), f"{name}: called with {len(args)} arguments {args}, declared with {dummies}"


@pytest.mark.skipif(not os.path.exists(REF_IFACE), reason="reference tree not present")
def test_every_dummy_argument_of_the_reference_interface_is_used_by_its_body():
    """A forwarding body that never mentions one of its dummies has dropped an argument (the kit `tfm` aside: the library
    holds the kit it was initialised with)."""
    ref = "\n".join(_fortran_lines(REF_IFACE))
    decl = {m.group(1).lower(): [a.strip().lower() for a in m.group(2).split(",") if a.strip()]
            for m in re.finditer(r"module\s+(?:recursive\s+)?(?:subroutine|function)\s+(\w+)\s*\(([^)]*)\)", ref, re.I)}
    shim = "\n".join(_fortran_lines(F_SHIM))
    bodies = {m.group(1).lower(): m.group(2).lower()
              for m in re.finditer(r"module procedure\s+(\w+)(.*?)end procedure", shim, re.I | re.S)}
    assert set(decl) == set(bodies)
    for name, dummies in decl.items():
        for d in dummies:
            # axis_input: the recursion cursor of the reference's assemble/disassemble (one gather per distributed
            # axis, dist:70-368); the slab layout has a single distributed axis and gathers it in one step
            if d in ("tfm", "axis_input"):
                continue
            assert re.search(r"\b%s\b" % re.escape(d), bodies[name]), f"{name}: dummy `{d}` is never used"
