"""CPU-only: the C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every
function include/adseis.h declares; host-only helpers agree with the oracle; compute calls fail loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "adseis.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adseis_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(A):
    lib = A._lib.load()
    names = _declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libadseis_b200.so does not export %s" % n
    # and the Python binding table covers exactly the header
    assert sorted(A._lib.SIGNATURES) == names


def test_sass_is_sm100a_only(A):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", A._lib.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_pml_profiles_match_oracle(A, po):
    for use in [(True, True, True, True), (False, True, True, False)]:
        p = A.AcousticPropagatorParams(NX=61, NY=47, NSTEP=10, DELTAX=7.0, DELTAY=11.0, NPOINTS_PML=9, Rcoef=0.01,
                                       vp_ref=2345.0, USE_PML_XMIN=use[0], USE_PML_XMAX=use[1], USE_PML_YMIN=use[2],
                                       USE_PML_YMAX=use[3])
        A.compute_PML_Params_(p)
        s, t = po.acoustic_pml(61, 47, 7.0, 11.0, npml=9, Rcoef=0.01, vp_ref=2345.0, use=use)
        assert np.array_equal(p.Σx.reshape(-1), s) and np.array_equal(p.Σy.reshape(-1), t)


def test_cpml_profiles_match_oracle(A, po):
    lib = A._lib.load()
    p = A.ElasticPropagatorParams(NX=50, NY=40, NSTEP=10, DELTAX=2.0, DELTAY=3.0, DELTAT=1e-4, vp_ref=3300.0, f0=7.0,
                                  NPOINTS_PML=8)
    pc = p.to_c()
    for axis, n, h in ((0, 50, 2.0), (1, 40, 3.0)):
        a, b = np.empty(2 * n), np.empty(2 * n)
        A._lib.check(lib.adseis_elastic_cpml_profiles(C.byref(pc), axis, A._lib.pd(a), A._lib.pd(b)))
        a0, b0 = po.elastic_cpml_1d(n, h, 1e-4, npml=8, vp_ref=3300.0, alpha_max=2 * np.pi * 3.5)
        assert np.array_equal(a, a0) and np.array_equal(b, b0)
        assert (a != 0).sum() > 0


def test_slab_partition(A):
    lib = A._lib.load()
    for NX, n in ((4096, 8), (101, 4), (10, 3), (7, 7)):
        rows = []
        for r in range(n):
            s = A._lib.SlabC()
            A._lib.check(lib.adseis_slab_partition(NX, n, r, C.byref(s)))
            rows.append((s.row0, s.row1))
        assert rows[0][0] == 0 and rows[-1][1] == NX + 2
        assert all(rows[k][1] == rows[k + 1][0] for k in range(n - 1))
        sizes = [b - a for a, b in rows]
        interior = [sizes[0] - 1] + sizes[1:-1] + [sizes[-1] - 1] if n > 1 else [sizes[0] - 2]
        assert max(interior) - min(interior) <= 1


def test_no_cpu_fallback(A):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(A.AdseisError) as e:
        A.Context()
    assert e.value.code == A._lib.ECUDA


def test_header_is_plain_c_and_links(A, tmp_path):
    """include/adseis.h is the contract a Julia ccall / cgo / C caller binds: it must compile as strict C99 with no
    CUDA or C++ types, and a C program linked against the shared library must run (without a GPU every compute entry
    point reports ADSEIS_ECUDA through the error convention instead of crashing)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "adseis.h"
int main(void) {
  adseis_acoustic_params p;
  memset(&p, 0, sizeof p);
  p.NX = 61; p.NY = 47; p.NSTEP = 10; p.DELTAX = 7.0; p.DELTAY = 11.0; p.DELTAT = 1e-3;
  p.USE_PML_XMIN = p.USE_PML_XMAX = p.USE_PML_YMIN = p.USE_PML_YMAX = 1;
  p.NPOINTS_PML = 9; p.Rcoef = 0.01; p.vp_ref = 2345.0; p.PropagatorKernel = 1;
  double sx[63], ty[49];
  int rc = adseis_acoustic_pml_profiles(&p, sx, ty);          /* host-only helper: works without a GPU */
  adseis_ctx* ctx = 0;
  int rc2 = adseis_ctx_create(0, &ctx);                       /* needs a device */
  printf("%d %d %.17g %d %s\n", adseis_version(), rc, sx[1], rc2, rc2 ? adseis_last_error() : "ok");
  if (ctx) adseis_ctx_destroy(ctx);
  return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.dirname(A._lib.lib_path())
    libname = os.path.basename(A._lib.lib_path())
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-l:" + libname, "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    ver, rc, s1, rc2 = out.stdout.split()[:4]
    assert int(ver) >= 1 and int(rc) == 0 and float(s1) > 0
    import torch
    if not torch.cuda.is_available():
        assert int(rc2) == A._lib.ECUDA and "no CUDA device" in out.stdout or int(rc2) < 0
