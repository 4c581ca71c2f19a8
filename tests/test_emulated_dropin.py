"""HOST logic of the drop-in classes on the CPU: the bodies of the GPU drop-in tests (tests/test_gpu_dropin.py) run
with the C ABI emulated by the oracle (tests/emulib.py), against the golden outputs of the unmodified reference.
What this covers: `GetEpochTimeOverlap`, `ModGrid`, `SatTrack`, `Model2SatTrack` (fused and slab paths, every record
provider incl. the NetCDF-3 fast path, mapping reuse, output packaging), the CLI and the NetCDF writers.  What it
does not: the CUDA kernels -- those are held to the same golden outputs by the very same test bodies under `-m gpu`."""
import numpy as np
import pytest

import gonzag_b200 as gz
from gonzag_b200 import _lib as L
from tests import emulib
from tests import test_gpu_dropin as T


@pytest.fixture()
def emu(monkeypatch):
    lib = emulib.EmuLib()
    monkeypatch.setattr(L, "lib", lambda: lib)
    yield lib
    gz.release_cached_context()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cli_sequence_matches_reference(emu, golden_grid, tag):
    T.test_cli_sequence_matches_reference(golden_grid, tag)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cli_sequence_from_netcdf3_files(emu, golden_grid, tmp_path, tag):
    T.test_cli_sequence_from_netcdf3_files(golden_grid, tmp_path, tag)


@pytest.mark.parametrize("tag", ["a", "c"])
def test_cli_main_from_files(emu, golden_grid, tmp_path, tag):
    T.test_cli_main_from_files(golden_grid, tmp_path, tag)


@pytest.mark.parametrize("provider", ["array", "callable", "netcdf3"])
def test_model_window_not_starting_at_record_0(emu, golden_window, tmp_path, provider):
    T.test_model_window_not_starting_at_record_0(golden_window, tmp_path, provider)


def test_streaming_provider_and_staged_paths(emu, golden_grid):
    T.test_streaming_provider_and_staged_paths(golden_grid)


def test_frame_switch(emu):
    T.test_frame_switch_on_device()


def test_multi_mission_on_one_resident_grid(emu, golden_grid):
    T.test_multi_mission_on_one_resident_grid(golden_grid)


def test_netcdf3_records_take_the_fast_path(emu, golden_grid, tmp_path, monkeypatch):
    """From file names the model records reach the staging slab through `NetCDF3Source.record_into` (one byte-swapping
    pass), not through the reader's masked arrays -- except the one probe read that tells the dtype."""
    from gonzag_b200.nc3io import NetCDF3Source
    from tests.conftest import write_grid_case_files
    calls = {"into": 0, "reader": 0}
    real_into, real_get = NetCDF3Source.record_into, NetCDF3Source.GetModel2DVar

    def into(self, *a, **k):
        ok = real_into(self, *a, **k)
        calls["into"] += bool(ok)
        return ok

    def get(self, *a, **k):
        calls["reader"] += 1
        return real_get(self, *a, **k)

    monkeypatch.setattr(NetCDF3Source, "record_into", into)
    monkeypatch.setattr(NetCDF3Source, "GetModel2DVar", get)
    g = golden_grid
    fm, ft, fl, varlsm, dist, Np = write_grid_case_files(g, "a", tmp_path)
    (it1, it2), _ = gz.GetEpochTimeOverlap(ft, fm)
    MG = gz.ModGrid(fm, it1, it2, fl, varlsm, distorded_grid=dist)
    try:
        ST = gz.SatTrack(ft, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
        T._check_results(R, g, "a")
        assert calls["reader"] == 1 and calls["into"] >= 2
    finally:
        MG.close()


def test_surface_functions(emu, golden_units, golden_surface):
    """The wrappers of tests/test_zz_gpu_surface.py (shapes, scalar / array returns, exits)."""
    from tests import test_zz_gpu_surface as S
    S.test_haversine_family(golden_units)
    S.test_iqh(golden_surface)


def test_mapping_cache_and_spectral_hand_off(emu, golden_grid, tmp_path):
    """f4: (1) a mapping saved to disk and loaded for the same grid + track skips K1-K3 and reproduces the products; a file
    that belongs to another track is ignored.  (2) The hand-off to the reference's spectral post-processing: its
    `FindUnbrokenSegments(time, distance, ssh_mod)` (gonzag/spectralysis.py:42) cuts the drop-in's output into the very
    segments it cuts the reference's own output into."""
    import types
    from oracle import ref_loader
    from tests.conftest import grid_case_sources
    g = golden_grid
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, "a")
    (it1, it2), _ = gz.GetEpochTimeOverlap(track, model)
    MG = gz.ModGrid(model, it1, it2, lsm, varlsm, distorded_grid=dist)
    try:
        ST = gz.SatTrack(track, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        path = tmp_path / "mapping.npz"
        assert gz.load_mapping(path, MG, ST) is None
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
        gz.save_mapping(path, R.BT, MG, ST)
        cached = gz.load_mapping(path, MG, ST)
        assert cached is not None and np.array_equal(cached.NP, R.BT.NP) and np.array_equal(cached.WB, R.BT.WB)
        n0 = MG.ctx.stats()["kernel_launches"]
        R2 = gz.Model2SatTrack(MG, "ssh", ST, "ssh", mapping=cached)
        assert MG.ctx.stats()["kernel_launches"] - n0 < 12
        T._check_results(R2, g, "a")
        other = types.SimpleNamespace(size=ST.size - 1, lat=ST.lat[1:], lon=ST.lon[1:])
        assert gz.load_mapping(path, MG, other) is None
        moved = types.SimpleNamespace(size=ST.size, lat=ST.lat + 1.0e-3, lon=ST.lon)
        assert gz.load_mapping(path, MG, moved) is None
        # ---- spectral hand-off
        root = ref_loader.find_reference()
        if root is None:
            pytest.skip("reference neither mounted nor copied (make -C oracle)")
        ref, _ = ref_loader.load_reference(root)
        want_mod = np.ma.MaskedArray(g["a_r_ssh_mod_data"], mask=g["a_r_ssh_mod_mask"])
        a0, a1 = ref.FindUnbrokenSegments(g["a_st_time"], g["a_r_dist"], want_mod, rcut_time=1.2, rcut_dist=7.8)
        b0, b1 = ref.FindUnbrokenSegments(R.time, R.distance, R.ssh_mod, rcut_time=1.2, rcut_dist=7.8)
        assert len(a0) > 3
        np.testing.assert_array_equal(a0, b0)
        np.testing.assert_array_equal(a1, b1)
    finally:
        MG.close()
