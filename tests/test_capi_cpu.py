"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
and its configuration helpers agree with the Python mirror.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from scale_letkf_b200 import capi, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from scale_letkf_b200 import build
    build.build()
    return capi.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "letkf_b200.h")).read()
    names = set(re.findall(r"\b(letkf_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert names == set(capi.PROTOTYPES), "capi.PROTOTYPES out of sync with the header"


def test_struct_layouts_match_ctypes_mirror(lib):
    sizes = (C.c_int32 * 4)()
    lib.letkf_b200_abi_sizes(C.byref(sizes))
    assert list(sizes) == [C.sizeof(capi.Config), C.sizeof(capi.CtypeInfo), C.sizeof(capi.Obs),
                           C.sizeof(capi.DasArgs)]


def test_defaults_match_python_mirror(lib):
    a = capi.Config()
    lib.letkf_b200_config_defaults(C.byref(a))
    b = config.default_config()
    assert config.config_bytes(a) == config.config_bytes(b)
    lib.letkf_b200_config_resolve(C.byref(a))
    config.resolve_config(b)
    assert config.config_bytes(a) == config.config_bytes(b)
    assert a.HORI_LOCAL[21] == 500.0e3 and a.VERT_LOCAL[21] == 1000.0 and a.VERT_LOCAL[5] == 0.4
    # default-REAL literals of letkf_obs.f90:27-28 widened to double
    assert a.dist_zero_fac == 3.6514837741851807
    assert a.dist_zero_fac_square == 13.333333015441895


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = config.default_config(MEMBER=8, nlon=8, nlat=8, nlev=2)
    h = C.c_void_p()
    r = lib.letkf_b200_create(C.byref(cfg), 0, C.byref(h))
    assert r == capi.ECUDA and not h.value   # no CPU fallback


def test_create_rejects_bad_config(lib):
    cfg = config.default_config(MEMBER=5000, nlon=8, nlat=8, nlev=2)   # > LETKF_B200_MAX_MEMBER (4096)
    h = C.c_void_p()
    assert lib.letkf_b200_create(C.byref(cfg), 0, C.byref(h)) == capi.EINVAL
    assert b"sm_100a" in lib.letkf_b200_build_info()


def test_header_is_plain_c_and_example_compiles(tmp_path):
    """include/letkf_b200.h must be consumable from C (the Fortran ISO_C_BINDING side sees exactly this ABI):
    the C example compiles with -std=c99 -Wall -Werror and every function it calls is exported by the library."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    obj = str(tmp_path / "das_from_c.o")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), "-c",
                           os.path.join(root, "examples", "das_from_c.c"), "-o", obj])
    und = subprocess.check_output(["nm", "-u", obj], text=True)
    wanted = {ln.split()[-1] for ln in und.splitlines() if "letkf_b200_" in ln}
    assert wanted and all(hasattr(capi.load_library(), w) for w in wanted), wanted
