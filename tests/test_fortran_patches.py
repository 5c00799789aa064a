"""fortran/patches/*.patch are generated from the unmodified reference sources by
fortran/make_patches.py; they must be current and apply cleanly.  Needs /root/reference (absent on
the GPU box) and the `patch` tool: skipped otherwise."""
import os
import shutil
import subprocess
import sys

import pytest

from common import ROOT

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Model")) or shutil.which("patch") is None, reason="reference tree or patch(1) not available")
def test_patches_are_current_and_apply(tmp_path):
    want = {n: open(os.path.join(ROOT, "fortran", "patches", n), encoding="latin-1", newline="").read()
            for n in sorted(os.listdir(os.path.join(ROOT, "fortran", "patches")))}
    assert set(want) == {"LTRANS.f90.patch", "boundary_module.f90.patch", "hydrodynamic_module.f90.patch", "makefile.patch", "settlement_module.f90.patch"}
    work = tmp_path / "ref"
    shutil.copytree(os.path.join(REF, "Model"), work / "Model")
    for n, text in want.items():
        r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", os.path.join(ROOT, "fortran", "patches", n)], cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, (n, r.stdout, r.stderr)
    src = open(work / "Model" / "LTRANS.f90", encoding="latin-1").read()
    for needle in ("CALL gpu_init()", "call gpu_step()", "call gpu_fetch()", "CALL gpu_fin()", "subroutine gpu_init()", "ltgpu_rotate_hydro(gpu_ctx)"):
        assert needle in src, needle
    assert "call update_particles()" not in src.split("subroutine update_particles")[0]      # the CPU loop is no longer called
    # every ltgpu_* entry point of the header has an INTERFACE in ltgpu_mod.f90
    import re
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ltrans_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(ltgpu_[a-z_0-9]+)\s*\(", hdr)))
    mod = open(os.path.join(ROOT, "fortran", "ltgpu_mod.f90")).read()
    for nm in names:
        assert "NAME='%s'" % nm in mod, nm


def test_fortran_params_type_mirrors_the_c_struct():
    """TYPE, BIND(C) :: ltgpu_params of fortran/ltgpu_mod.f90 against struct ltgpu_params of the header: the same
    fields in the same order with interoperable kinds (a BIND(C) type is laid out like the C struct only then)"""
    import re
    h = open(os.path.join(ROOT, "include", "ltrans_b200.h")).read()
    m = re.search(r"typedef struct ltgpu_params \{(.*?)\} ltgpu_params;", h, re.S) or re.search(r"struct ltgpu_params \{(.*?)\};", h, re.S)
    body = re.sub(r"//.*", "", re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S))
    cfields = []
    for decl in body.split(";"):
        parts = decl.replace(",", " ").split()
        if parts:
            cfields += [(parts[0], nm) for nm in parts[1:]]
    f = open(os.path.join(ROOT, "fortran", "ltgpu_mod.f90"), encoding="latin-1").read()
    m = re.search(r"TYPE, BIND\(C\) :: ltgpu_params(.*?)END TYPE", f, re.S | re.I)
    ffields = []
    for line in m.group(1).splitlines():
        line = line.split("!")[0].strip()
        if "::" in line:
            typ, names = line.split("::")
            ffields += [(typ.replace(" ", "").upper(), nm.strip()) for nm in names.split(",")]
    assert [n.lower() for _, n in cfields] == [n.lower() for _, n in ffields]
    ok = {"int32_t": {"INTEGER(C_INT)", "INTEGER(C_INT32_T)"}, "int64_t": {"INTEGER(C_INT64_T)", "INTEGER(C_LONG_LONG)"},
          "double": {"REAL(C_DOUBLE)"}, "float": {"REAL(C_FLOAT)"}}
    for (ct, cn), (ft, fn) in zip(cfields, ffields):
        assert ft in ok[ct], (cn, ct, ft)


def test_fortran_interfaces_take_as_many_arguments_as_the_c_prototypes():
    import re
    h = re.sub(r"//.*", "", re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ltrans_b200.h")).read(), flags=re.S))
    protos = {}
    for m in re.finditer(r"\b(ltgpu_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    f = re.sub(r"&\s*\n\s*&?", "", open(os.path.join(ROOT, "fortran", "ltgpu_mod.f90"), encoding="latin-1").read())
    iface = {}
    for m in re.finditer(r"(?:FUNCTION|SUBROUTINE)\s+(\w+)\s*\(([^)]*)\)\s*(?:RESULT\(\w+\)\s*)?BIND\(C,\s*NAME='(ltgpu_[a-z_0-9]+)'\)", f, re.I):
        args = m.group(2).strip()
        iface[m.group(3)] = 0 if not args else len(args.split(","))
    assert len(protos) == 28 and protos == iface, {k: (protos.get(k), iface.get(k)) for k in set(protos) | set(iface) if protos.get(k) != iface.get(k)}
