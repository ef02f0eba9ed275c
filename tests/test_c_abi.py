"""The drop-in boundary without Python in the loop: a C11 client of include/tse.h (tests/c_abi/abi_check.c) and the C++
stand-alone driver (driver/prim_main.cpp: namelist, time loop, printstate, error norms -- the reference's prim_main.F90:37-203)."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NAMELIST = """&ctl_nl
  NThreads          = 1
  partmethod        = 4
  topology          = "cube"
  test_case         = "dcmip1-2"
  ne                = 8
  qsize             = 4
  ndays             = 0
  nmax              = 3                 ! three tracer steps = one remap cycle
  statefreq         = 3
  tstep             = 400
  qsplit            = 1
  rsplit            = 3
  nu_q              = 6e16 !2e16
  limiter_option    = 8
  hypervis_order    = 2
  hypervis_subcycle = 1
  prescribed_wind   = 1
/
&filter_nl
  filter_type       = "taylor"
  filter_mu         = 0.04D0
/
&vert_nl
  vform             = "ccm"
  vfile_mid         = "%(mid)s"
  vfile_int         = "%(int)s"
/
&analysis_nl
  output_varnames1 = 'Q','Q3','geo'
/
"""


def _have_gpu():
    from transport_se_b200.advection import cuda_lib
    return cuda_lib().tse_device_count() > 0


def _vcoord_files(tmp_path):
    """The ACME 72-level tables in the reference's ascii format (hybvcoord_mod.F90:84-153) and as a flat list for abi_check."""
    from transport_se_b200.mesh import load_vcoord
    hv = load_vcoord()
    fi, fm, flat = tmp_path / "acme-72i.ascii", tmp_path / "acme-72m.ascii", tmp_path / "vcoord.txt"
    fi.write_text("    73   ! hyai\n" + "\n".join(repr(float(x)) for x in hv["hyai"]) + "\n    73   ! hybi\n" + "\n".join(repr(float(x)) for x in hv["hybi"]) + "\n")
    fm.write_text("    72   ! hyam\n" + "\n".join(repr(float(x)) for x in hv["hyam"]) + "\n    72   ! hybm\n" + "\n".join(repr(float(x)) for x in hv["hybm"]) + "\n")
    flat.write_text("\n".join(repr(float(x)) for k in ("hyai", "hybi", "hyam", "hybm") for x in hv[k]) + "\n")
    return str(fi), str(fm), str(flat)


def test_header_is_c11_and_clients_link(built):
    """include/tse.h compiles as strict C11 (-Wall -Wextra -pedantic -Werror) and both clients link against libtse_cuda.so."""
    from transport_se_b200 import _build
    assert os.access(_build.build_c_abi_check(), os.X_OK)
    assert os.access(_build.build_driver(), os.X_OK)


def test_driver_reads_the_reference_namelist(built, tmp_path):
    """Namelist from stdin like the reference (prim_main < input.nl): groups, comments, quoted strings, Fortran reals.  Without a
    GPU the run must stop at tse_init with the library's message (no CPU fallback); with one it runs the three steps."""
    from transport_se_b200 import _build
    exe = _build.build_driver()
    fi, fm, _ = _vcoord_files(tmp_path)
    r = subprocess.run([exe], input=NAMELIST % {"mid": fm, "int": fi}, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    head = [l for l in r.stdout.splitlines() if l.strip().startswith("ne =")]
    assert head and re.search(r"ne = 8\s+nelem = 384\s+qsize = 4\s+tstep = 400\s+nu_q = 6e\+16\s+test_case = dcmip1-2\s+nEndStep = 3", head[0]), r.stdout
    if _have_gpu():
        assert r.returncode == 0, r.stderr
        assert "Finished main timestepping loop 3" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_c_client_matches_python_binding(built, tmp_path):
    from transport_se_b200 import _build
    from transport_se_b200.advection import TracerAdvection
    from transport_se_b200.mesh import Mesh, load_vcoord
    exe = _build.build_c_abi_check()
    _, _, flat = _vcoord_files(tmp_path)
    ne, qsize, test, tstep, nu_q = 4, 5, 11, 800.0, 5e17
    r = subprocess.run([exe, str(ne), str(qsize), str(test), repr(tstep), repr(nu_q), flat], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = re.findall(r"tracer (\d+) mass (\S+) qmin (\S+) qmax (\S+) hash ([0-9a-f]{16})", r.stdout)
    assert len(rows) == qsize
    m = Mesh(ne)
    adv = TracerAdvection(m, m.local_view(), load_vcoord(), qsize=qsize, nu_q=nu_q)
    adv.dcmip_init(test)
    nstep = adv.prim_run_subcycle(tstep, 0)
    tl = 1 if nstep % 2 == 0 else 2
    mass, (qmn, qmx), fh = adv.diag_mass(tl), adv.diag_qminmax(tl), adv.diag_field_hash(tl)
    for q, ms, a, b, hx in rows:
        q = int(q)
        assert float.fromhex(ms) == mass[q] and float.fromhex(a) == qmn[q] and float.fromhex(b) == qmx[q]
        assert int(hx, 16) == int(fh[q])      # every value of the field is bitwise the same through both bindings
    adv.close()


@pytest.mark.gpu
def test_prim_main_driver_reproduces_readme_norms(built, tmp_path):
    """The complete ne8 DCMIP 1-2 verification run (216 steps) through the stand-alone driver: README:96 to 5 digits."""
    from transport_se_b200 import _build
    exe = _build.build_driver()
    fi, fm, _ = _vcoord_files(tmp_path)
    r = subprocess.run([exe, "ne=8", "tstep=400", "nu_q=6e16", "qsize=4", "test_case=dcmip1-2", "ndays=1", "statefreq=108", "vfile_int=" + fi,
                        "vfile_mid=" + fm], stdin=subprocess.DEVNULL, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    print(r.stdout[-1500:])
    mt = re.search(r"L1 = (\S+)\s+L2 = (\S+)\s+Linf = (\S+)\s+q_max = (\S+)\s+q_min = (\S+)", r.stdout)
    got = [float(x) for x in mt.groups()]
    gold = [0.307665, 0.622099, 0.839133, 0.813105, -9.385639e-06]
    for g, w in zip(got[:4], gold[:4]):
        assert abs(g - w) < 5e-6, (got, gold)
    assert abs(got[4] - gold[4]) < 5e-11
    # tracer mass of the Hadley-layer tracer conserved over the run (prim_printstate's Q mass line)
    last = [l for l in r.stdout.splitlines() if l.strip().startswith("Q2 ")][-1]
    assert abs(float(last.split("=")[-1])) < 1e-12
