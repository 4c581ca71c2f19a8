"""CPU-only tests: the C ABI library loads and exports what include/gonzag_b200.h declares,
the host-side planning logic, and the failure mode without a GPU (no CPU fallback)."""
import ctypes
import types
import os

import numpy as np
import pytest

import __graft_entry__ as ge
from gonzag_b200 import _lib as L
from gonzag_b200 import synthetic as syn
from gonzag_b200.mod2sat import plan_slabs
from gonzag_b200.multi import segment_bounds


@pytest.fixture(scope="module")
def built():
    from gonzag_b200 import build
    return build.build()


def test_library_exports_header_symbols(built):
    lib = ctypes.CDLL(built)
    syms = ge.header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(L.exported_symbols()) == syms        # the ctypes binding covers the whole header
    assert b"sm_100a" in L.lib().gzb_version()


def test_sass_is_sm100a(built):
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % built).read()
    assert "sm_100a" in out


def test_no_cpu_fallback(built):
    import gonzag_b200 as gzb
    if gzb.device_count() > 0:
        pytest.skip("a GPU is present")
    lat, lon = syn.regional_grid(8, 9)
    with pytest.raises(gzb.GonzagB200Error) as e:
        gzb.GridContext(lat, lon)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(gzb.GonzagB200Error):
        gzb.BilinTrack(np.zeros(3), np.zeros(3), lat, lon)


def test_argument_validation(built):
    import gonzag_b200 as gzb
    lat, lon = syn.regional_grid(8, 9)
    with pytest.raises(ValueError):
        gzb.GridContext(lat, lon[:, :-1])
    # C-level checks do not need a device
    h = ctypes.c_void_p()
    rc = L.lib().gzb_create(0, 2, 2, L.ptr(lat), L.ptr(lon), None, None, -1, ctypes.byref(h))
    assert rc == L.GZB_ERR_ARG and h.value is None
    assert L.lib().gzb_sync(None) == L.GZB_ERR_ARG
    assert L.lib().gzb_destroy(None) == L.GZB_OK


def test_plan_slabs():
    tmod = 3600.0 * np.arange(10)
    t = np.array([10.0, 20.0, 3599.0, 3600.0, 7300.0, 30000.0, 30001.0])
    assert plan_slabs(t, tmod, 100) == [(0, 3), (8, 9)]
    assert plan_slabs(t, tmod, 2) == [(0, 1), (1, 2), (2, 3), (8, 9)]
    assert plan_slabs(t[:3], tmod, 3) == [(0, 1)]
    assert plan_slabs(np.array([]), tmod, 3) == []
    with pytest.raises(ValueError):
        plan_slabs(np.array([-1.0]), tmod, 3)
    with pytest.raises(ValueError):
        plan_slabs(np.array([9 * 3600.0]), tmod, 3)      # the last record time is excluded
    # every bracket [k, k+1] of every point lies inside one slab
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 9 * 3600.0 - 1, 500))
    for cap in (2, 3, 5, 64):
        slabs = plan_slabs(t, tmod, cap)
        kt = np.searchsorted(tmod, t, side="right") - 1
        for k in kt:
            assert any(a <= k and k + 1 <= b for a, b in slabs)
        assert all(b - a + 1 <= cap for a, b in slabs)


def test_segment_bounds():
    assert segment_bounds(10, 3) == [0, 4, 7, 10]
    assert segment_bounds(2, 4) == [0, 1, 2, 2, 2]
    b = segment_bounds(1000003, 8)
    assert b[0] == 0 and b[-1] == 1000003 and max(np.diff(b)) - min(np.diff(b)) <= 1


def test_synthetic_inputs_have_the_advertised_shape():
    lat, lon = syn.orca_like_grid(146, 182)
    assert lat.shape == (146, 182)
    np.testing.assert_array_equal(lat[:, 0], lat[:, -2])      # exact E-W overlap copies
    np.testing.assert_array_equal(lon[:, -1], lon[:, 1])
    assert lon.min() >= 0.0 and lon.max() < 360.0
    m = syn.land_mask(lat, lon)
    assert m.dtype == np.int64 and m[0, 0] == 0 and 0.1 < 1.0 - m.mean() < 0.6
    t, y, x = syn.orbit_track(20000, **syn.JASON)
    assert np.all(np.diff(t) > 0) and np.abs(y).max() < 66.1 and len(t) < 20000   # land gaps
    rd, r = syn.search_params(1.0 / 12.0)
    assert r == 26 or r == 27
    assert abs(rd - 0.75 * 111.11 / 12.0) < 1e-9


def test_header_is_plain_c_and_links(built, tmp_path):
    """include/gonzag_b200.h is the boundary a non-Python host binds: it must compile as C99 with warnings as
    errors, and a C program must link against the library and get the argument checks the ABI promises (no
    device needed for those)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "gonzag_b200.h"
int main(void) {
    gzb_ctx* h = NULL;
    gzb_stats st;
    double lat[4] = {0, 0, 1, 1}, lon[4] = {0, 1, 0, 1};
    memset(&st, 0, sizeof st);
    if (!strstr(gzb_version(), "sm_100a")) return 1;
    if (gzb_create(0, 2, 2, lat, lon, NULL, NULL, -1, &h) != GZB_ERR_ARG || h != NULL) return 2;   /* grid too small */
    if (gzb_create(0, 3, 3, NULL, lon, NULL, NULL, -1, &h) != GZB_ERR_ARG) return 3;
    if (!gzb_last_global_error() || !gzb_last_global_error()[0]) return 4;
    if (gzb_sync(NULL) != GZB_ERR_ARG || gzb_destroy(NULL) != GZB_OK) return 5;
    if (gzb_heading(0, 0, NULL, NULL, NULL, NULL, NULL) != GZB_OK) return 6;               /* n = 0: nothing to do */
    if (gzb_radius_earth(0, 3, NULL, NULL) != GZB_ERR_ARG) return 7;
    if (gzb_haversine_sclr(0, 3, lat, lon, lat, NULL, lat) != GZB_ERR_ARG) return 8;
    printf("sizeof(gzb_stats)=%u\n", (unsigned)sizeof(gzb_stats));
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(built)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-lgonzag_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "sizeof(gzb_stats)=%d" % ctypes.sizeof(L.Stats) in r.stdout          # the ctypes mirror has the same layout


def test_pack_mask_equals_astype(built):
    """gzb_pack_mask (host-only, threaded): the int64 {0,1} mask of the reference -> int8, for every integer width,
    lengths around the thread / alignment boundaries, and values whose low byte matters."""
    rng = np.random.default_rng(0)
    for dt in (np.int64, np.int32, np.int16, np.uint16, np.uint32, np.uint64, np.int8):
        info = np.iinfo(dt)
        for n in (0, 1, 5, 63, 64, 65, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 2500001):
            m = rng.integers(max(info.min, -70000), min(info.max, 70000), n).astype(dt)
            o = np.full(n + 1, 77, np.int8)
            L.check(L.lib().gzb_pack_mask(L.ptr(m), m.dtype.itemsize, n, L.ptr(o)))
            assert np.array_equal(o[:n], m.astype(np.int8)) and o[n] == 77, (dt, n)
    assert L.lib().gzb_pack_mask(None, 8, 4, None) == L.GZB_ERR_ARG
    assert L.lib().gzb_pack_mask(L.ptr(np.zeros(4)), 3, 4, L.ptr(np.zeros(4, np.int8))) == L.GZB_ERR_ARG


def test_mask8_paths():
    """GridContext._mask8 without a device: int64 (the reference's dtype) through gzb_pack_mask, int8 / bool viewed,
    anything else through NumPy; masked arrays contribute their data."""
    from gonzag_b200.context import GridContext
    fake = types.SimpleNamespace(Nj=7, Ni=9)
    rng = np.random.default_rng(1)
    m = (rng.uniform(size=(7, 9)) < 0.6).astype(np.int64)
    for arr in (m, m.astype(np.int32), m.astype(np.int8), m.astype(bool), m.astype(np.float64), m[:, ::-1],
                np.ma.masked_array(m, mask=(m == 0)), np.asfortranarray(m)):
        got = GridContext._mask8(fake, arr)
        assert got.dtype == np.int8 and got.flags["C_CONTIGUOUS"]
        np.testing.assert_array_equal(got, np.ma.getdata(arr).astype(np.int8))
    with pytest.raises(ValueError):
        GridContext._mask8(fake, m[:, :-1])


def test_copy_swap(built):
    """gzb_copy_swap (host-only, threaded): byte-reversing copy for every item size, lengths around the thread
    boundaries, an unaligned source (a NetCDF-3 variable starts anywhere in the file)."""
    rng = np.random.default_rng(2)
    for dt in ("u1", "u2", "u4", "u8", "f4", "f8"):
        size = np.dtype(dt).itemsize
        for n in (0, 1, 7, 64, 65, (1 << 19) + 3, 3 * (1 << 20) + 1):
            raw = rng.integers(0, 256, n * size + 3, dtype=np.uint8)
            be = raw[3:].view(">" + dt)                       # unaligned, big-endian
            out = np.zeros(n + 1, dtype="<" + dt)
            L.check(L.lib().gzb_copy_swap(L.ptr(out), ctypes.c_void_p(be.__array_interface__["data"][0]), size, n, 1))
            assert np.array_equal(out[:n].view("u" + str(size)), be.astype("<" + dt).view("u" + str(size))), (dt, n)
            assert out[n] == 0
            L.check(L.lib().gzb_copy_swap(L.ptr(out), ctypes.c_void_p(be.__array_interface__["data"][0]), size, n, 0))
            assert out[:n].tobytes() == be.tobytes()
    assert L.lib().gzb_copy_swap(None, None, 4, 3, 1) == L.GZB_ERR_ARG
    assert L.lib().gzb_copy_swap(L.ptr(np.zeros(4)), L.ptr(np.zeros(4)), 3, 4, 1) == L.GZB_ERR_ARG


def test_integration_doc_matches_the_binding():
    """The ctypes stub shown in INTEGRATION.md declares the same argument lists as the packaged binding."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    env = {"C": ctypes, "_P": ctypes.c_void_p, "_I64": ctypes.c_int64}
    found = re.findall(r"_L\.(gzb_\w+)\.argtypes\s*=\s*(\[[^\]]*\])", text)
    assert len(found) >= 3
    for name, lst in found:
        assert [a for a in eval(lst, env)] == list(L._SIGNATURES[name][1]), name


def test_every_option_is_documented_and_known_to_the_emulation():
    """Every integer option gzb_set_option accepts (csrc/gzb_grid.cu) is described in the header's comment and accepted by
    the CPU emulation of the C ABI (tests/emulib.py), so that host tests and documentation follow the library."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "gonzag_b200", "csrc", "gzb_grid.cu")).read()
    body = src[src.index('extern "C" int gzb_set_option'):]
    body = body[:body.index("\n}\n")]
    names = sorted(set(re.findall(r'!strcmp\(name, "([a-z0-9_]+)"\)', body)))
    assert len(names) >= 20, names
    header = open(os.path.join(root, "include", "gonzag_b200.h")).read()
    emu = open(os.path.join(root, "tests", "emulib.py")).read()
    known = emu[emu.index("known = ("):]
    known = known[:known.index(")")]
    for n in names:
        assert '"%s"' % n in header, "option %s is not described in include/gonzag_b200.h" % n
        assert '"%s"' % n in known, "option %s is not accepted by tests/emulib.py" % n
