"""The C-ABI shared library loads and exports every symbol include/mytrim_b200.h declares
(no compute calls: there is no GPU in the build container)."""
import ctypes as C
import os
import re

from mytrim_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mytrim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_header(tmp_path):
    """sizeof() of every ABI struct as compiled by gcc equals the ctypes / numpy mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu", '
                   'sizeof(mtb_config), sizeof(mtb_element), sizeof(mtb_material), sizeof(mtb_geometry), '
                   'sizeof(mtb_counters), sizeof(mtb_ion), sizeof(mtb_record), sizeof(mtb_ion_log), '
                   'sizeof(mtb_event));return 0;}\n' % os.path.join(ROOT, "include", "mytrim_b200.h"))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(capi.Config), C.sizeof(capi.Element), C.sizeof(capi.Material), C.sizeof(capi.Geometry),
            C.sizeof(capi.Counters), capi.ION_DTYPE.itemsize, capi.RECORD_DTYPE.itemsize,
            capi.IONLOG_DTYPE.itemsize, capi.EVENT_DTYPE.itemsize]
    assert got == want
    assert capi.load_library().mtb_version().startswith(b"mytrim_b200")


def test_no_device_fails_loudly():
    """Without a GPU the engine must refuse to run (no CPU fallback of any kind)."""
    import pytest
    lib = capi.load_library()
    if lib.mtb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.MytrimError) as e:
        capi.Engine()
    assert e.value.code == capi.ENODEV


def test_product_does_not_reference_oracle():
    """Nothing under mytrim_b200/, include/ or apps/ may include, link or import oracle/ or tests/."""
    bad = []
    for top in ("mytrim_b200", "include", "apps"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".so", ".o", ".pyc", ".inc")):
                    continue
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r'#include\s*"[^"]*oracle', text) or re.search(r"(import|from)\s+(oracle|tests)\b", text) \
                        or "liboracle" in text or "hostsim" in text.replace("tests/hostsim.cpp", ""):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_unmodified_reference_apps_compile_against_facade():
    """Every reference driver that does not need jsoncpp compiles, unmodified, against include/mytrim
    and links with libmytrim_b200.so (built by oracle/Makefile `facade_apps`; build container only)."""
    import pytest
    if not os.path.exists("/root/reference/trim.C"):
        pytest.skip("reference tree not mounted")
    apps = os.path.join(ROOT, "oracle", "_ref", "facade_apps")
    want = ["mytrim_layers", "mytrim_uo2", "mytrim_solid", "mytrim_solid2", "mytrim_wire", "mytrim_wire2",
            "mytrim_clusters", "mytrim_ODS", "mytrim_bobmsq", "mytrim_distance", "mytrim_moose_verification"]
    for a in want:
        assert os.path.exists(os.path.join(apps, a)), a
