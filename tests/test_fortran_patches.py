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
