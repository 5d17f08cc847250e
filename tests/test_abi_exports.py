"""The C-ABI library loads on a CPU-only box and exports every symbol include/dumux_b200.h declares; the product path
fails loudly (no CPU fallback) when no CUDA device is visible.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dumux_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from dumux_b200 import binding
    if not os.path.exists(binding.LIB_PATH):
        g.build()
    return ctypes.CDLL(binding.LIB_PATH)


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_export_list_matches_header():
    from dumux_b200 import binding
    assert sorted(binding.EXPORTS) == _declared_symbols()


def test_header_cites_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("fvassembler.hh:179-207", "istlsolvers.hh:273", "newtonsolver.hh:976-1072", "jacobianpattern.hh:27-52",
                 "gridmanager_yasp.hh", "fvproblem.hh"):
        assert cite in src, cite


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C (gcc -std=c99), no C++ or torch types."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "dumux_b200.h"\nint main(void){ dmx_options o; dmx_newton_params p; dmx_newton_report r; (void)o;(void)p;(void)r; return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                           str(tmp_path / "t.o")])


def test_struct_layouts_match_ctypes(lib):
    """ctypes mirrors of the POD structs have the sizes the C compiler gives them."""
    import subprocess
    import tempfile
    from dumux_b200 import binding
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "dumux_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(dmx_options), '
                             'sizeof(dmx_newton_params), sizeof(dmx_newton_report), sizeof(dmx_amg_params));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(binding.DmxOptions), ctypes.sizeof(binding.DmxNewtonParams), ctypes.sizeof(binding.DmxNewtonReport),
                     ctypes.sizeof(binding.DmxAmgParams)]


def test_defaults_match_reference_parameters(lib):
    """dumux/common/parameters.cc:231-260, nonlinear/newtonsolver.hh:1213-1247, linear/linearsolverparameters.hh:56-73."""
    from dumux_b200 import binding
    o = binding.DmxOptions()
    lib.dmx_default_options(ctypes.byref(o))
    assert (o.enable_gravity, o.gravity, o.upwind_weight, o.fd_method, o.base_eps) == (1, 9.81, 1.0, 1, 1e-10)
    assert o.privar_magnitude[0] < 0 and o.privar_magnitude[1] < 0 and o.extrusion == 1.0
    p = binding.DmxNewtonParams()
    lib.dmx_default_newton_params(ctypes.byref(p))
    assert (p.max_relative_shift, p.min_steps, p.max_steps, p.lin_reduction, p.lin_maxit) == (1e-8, 2, 18, 1e-6, 250)


def test_no_cpu_fallback():
    """Without a CUDA device dmx_create must fail and the Python host layer must raise -- never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the loud-failure path is exercised on the CPU-only box")
    from dumux_b200 import binding, problems
    L = binding.load_library()
    h = ctypes.c_void_p()
    assert L.dmx_create(ctypes.byref(h), 0) != 0 and not h
    with pytest.raises(binding.DmxError):
        binding.Engine(problems.onep_incompressible((4, 4)))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under dumux_b200/ or include/ may reference it."""
    bad = []
    for base in ("dumux_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".hh", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|#include\s*[\"<][^\">]*oracle|liboracle|oracle_py|orc_[a-z_]+\()", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
