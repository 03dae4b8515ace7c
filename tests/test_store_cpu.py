"""Forecast store (swift_b200/store.py) against the layout the reference writes: utils/io.py:71-231 (variable compression,
array dims / shapes / chunks) and generate.py:63-74,139-152 (channel runs per variable, level stacking).  Neither zarr nor
xarray is installed here, so the format-2 directory layout is checked file by file."""
import json
import os
import zlib

import numpy as np
import pytest

from swift_b200.generate import era5_variables
from swift_b200.store import ForecastStore, compress_variables, variable_channels


def test_compress_variables_and_channel_runs_swift_b():
    v = era5_variables()
    assert len(v) == 69
    c = compress_variables(v)
    assert list(c) == ["2m_temperature", "10m_u_component_of_wind", "10m_v_component_of_wind", "mean_sea_level_pressure",
                       "geopotential", "u_component_of_wind", "v_component_of_wind", "temperature", "specific_humidity"]
    assert c["2m_temperature"] == [] and c["mean_sea_level_pressure"] == []
    assert c["geopotential"] == [50, 100, 150, 200, 250, 300, 400, 500, 600, 700, 850, 925, 1000]
    ch = variable_channels(v)
    assert ch["2m_temperature"] == [0] and ch["mean_sea_level_pressure"] == [3]
    assert ch["geopotential"] == list(range(4, 17)) and ch["specific_humidity"] == list(range(56, 69))
    # names whose suffix is not purely numeric stay single-level (io.py:75: r"^(.*)_(\d+)$")
    assert compress_variables(["a_1b", "b_2", "b_10", "c"]) == {"a_1b": [], "b": [2, 10], "c": []}


def _rollout(bs, steps, n_ch, h, w, seed):
    return np.random.default_rng(seed).standard_normal((bs, steps + 1, n_ch, h, w)).astype(np.float32)


VARS = ["t2m", "z_500", "z_850", "z_1000", "msl", "q_500", "q_850"]     # runs: t2m[0] z[1:4] msl[4] q[5:7]


@pytest.mark.parametrize("layout,level", [("trajectory", 0), ("trajectory", 1), ("step", 0)])
def test_zarr_directory_layout_and_round_trip(tmp_path, layout, level):
    path = str(tmp_path / "fc.zarr")
    n_ic, members, steps, h, w = 3, 2, 4, 6, 8
    lat, lon = np.linspace(-80, 80, h), np.arange(w) * 45.0
    st = ForecastStore.create(path, VARS, n_ic, members, steps, lat, lon, interval_hours=6, layout=layout,
                              compress_level=level)
    # ---- metadata as zarr format 2 / xarray expect it
    assert json.load(open(os.path.join(path, ".zgroup"))) == {"zarr_format": 2}
    za = json.load(open(os.path.join(path, "z", ".zarray")))
    assert za["shape"] == [n_ic, members, steps + 1, 3, h, w] and za["dtype"] == "<f4" and za["order"] == "C"
    assert za["chunks"] == ([1, 1, steps + 1, 3, h, w] if layout == "trajectory" else [1, 1, 1, 3, h, w])
    assert za["fill_value"] == 0.0 and za["zarr_format"] == 2
    assert za["compressor"] == ({"id": "zlib", "level": 1} if level else None)
    assert json.load(open(os.path.join(path, "z", ".zattrs")))["_ARRAY_DIMENSIONS"] == [
        "time", "number", "prediction_timedelta", "level", "latitude", "longitude"]
    assert json.load(open(os.path.join(path, "msl", ".zarray")))["shape"] == [n_ic, members, steps + 1, h, w]
    assert json.load(open(os.path.join(path, "msl", ".zattrs")))["_ARRAY_DIMENSIONS"] == [
        "time", "number", "prediction_timedelta", "latitude", "longitude"]
    cons = json.load(open(os.path.join(path, ".zmetadata")))
    assert cons["zarr_consolidated_format"] == 1 and "z/.zarray" in cons["metadata"] and ".zgroup" in cons["metadata"]
    np.testing.assert_array_equal(st.coordinate("prediction_timedelta"), np.arange(steps + 1) * 6)
    assert json.load(open(os.path.join(path, "prediction_timedelta", ".zattrs")))["units"] == "hours"
    assert json.load(open(os.path.join(path, "time", ".zattrs")))["units"] == "hours since 2020-01-01 00:00:00"
    np.testing.assert_array_equal(st.coordinate("number"), np.arange(members))
    np.testing.assert_allclose(st.coordinate("latitude"), lat)
    np.testing.assert_array_equal(st.coordinate("level"), np.arange(3))
    # ---- write every trajectory except (ic 2, member 1): it must read back as the fill value
    full = np.zeros((n_ic, members, steps + 1, len(VARS), h, w), dtype=np.float32)
    st2 = ForecastStore.open(path)
    assert (st2.layout, st2.n_ic, st2.members, st2.steps, st2.res) == (layout, n_ic, members, steps, (h, w))
    for j in range(n_ic):
        for m in range(members):
            if (j, m) == (2, 1):
                continue
            r = _rollout(1, steps, len(VARS), h, w, seed=10 * j + m)
            full[j, m] = r[0]
            st2.write_trajectories(j, m, r)
    np.testing.assert_array_equal(st.read_all(), full)
    np.testing.assert_array_equal(st.read("z"), full[:, :, :, 1:4])          # levels stacked after the lead-time axis
    np.testing.assert_array_equal(st.read("msl"), full[:, :, :, 4])
    # ---- one chunk file = raw C-order little-endian floats of the chunk
    key = "1.0.0.0.0.0" if layout == "trajectory" else "1.0.2.0.0.0"
    raw = open(os.path.join(path, "z", key), "rb").read()
    if level:
        raw = zlib.decompress(raw)
    want = full[1, 0, :, 1:4] if layout == "trajectory" else full[1, 0, 2, 1:4]
    assert raw == np.ascontiguousarray(want, dtype="<f4").tobytes()
    assert not os.path.exists(os.path.join(path, "z", "2.1.0.0.0.0"))
    assert not [f for f in os.listdir(os.path.join(path, "z")) if "tmp" in f]


def test_trajectory_layout_batched_chunks_and_edge_padding(tmp_path):
    """fast_create_empty_zarr's chunks (batch, 1, steps+1, ...) with n_ic not a multiple of batch."""
    path = str(tmp_path / "fc.zarr")
    n_ic, members, steps, h, w = 5, 1, 2, 4, 4
    st = ForecastStore.create(path, VARS, n_ic, members, steps, np.arange(h), np.arange(w), layout="trajectory", batch=2)
    assert json.load(open(os.path.join(path, "q", ".zarray")))["chunks"] == [2, 1, 3, 2, h, w]
    full = _rollout(n_ic, steps, len(VARS), h, w, seed=5)
    for s in range(0, n_ic, 2):
        st.write_trajectories(s, 0, full[s:s + 2])
    np.testing.assert_array_equal(st.read_all()[:, 0], full)
    assert os.path.getsize(os.path.join(path, "q", "2.0.0.0.0.0")) == 2 * 3 * 2 * h * w * 4      # full chunk on disk
    with pytest.raises(ValueError):
        st.write_trajectories(1, 0, full[1:2])                   # not aligned to a chunk row
    with pytest.raises(ValueError):
        st.write_step(0, 0, 0, full[0, 0])


def test_step_layout_streams_and_numpy_memmap(tmp_path):
    n_ic, members, steps, h, w = 2, 3, 3, 4, 6
    full = _rollout(n_ic * members, steps, len(VARS), h, w, seed=9).reshape(n_ic, members, steps + 1, len(VARS), h, w)
    st = ForecastStore.create(str(tmp_path / "s.zarr"), VARS, n_ic, members, steps, np.arange(h), np.arange(w), layout="step")
    npy = ForecastStore.create(str(tmp_path / "r.npy"), VARS, n_ic, members, steps, np.arange(h), np.arange(w), layout="numpy")
    for k in range(steps + 1):                                    # lead-time major, as the running rollout produces it
        for j in range(n_ic):
            for m in range(members):
                st.write_step(j, m, k, full[j, m, k])
                npy.write_step(j, m, k, full[j, m, k])
    npy.flush()
    np.testing.assert_array_equal(st.read_all(), full)
    loaded = np.load(str(tmp_path / "r.npy"), mmap_mode="r")     # the reference's reader (io.py:246)
    assert loaded.shape == (n_ic, members, steps + 1, len(VARS), h, w) and loaded.dtype == np.float32
    np.testing.assert_array_equal(loaded, full)
    np.testing.assert_array_equal(ForecastStore.open(str(tmp_path / "r.npy")).read_all(), full)
    npy.write_trajectories(1, 2, full[1:2, 2] * 2)
    npy.flush()
    np.testing.assert_array_equal(np.load(str(tmp_path / "r.npy"))[1, 2], full[1, 2] * 2)


def test_errors(tmp_path):
    with pytest.raises(ValueError):
        ForecastStore.create(str(tmp_path / "x"), VARS, 1, 1, 1, np.arange(2), np.arange(2), layout="hdf5")
    st = ForecastStore.create(str(tmp_path / "y.zarr"), VARS, 2, 2, 2, np.arange(2), np.arange(2))
    with pytest.raises(IndexError):
        st.write_trajectories(2, 0, _rollout(1, 2, len(VARS), 2, 2, 0))
    with pytest.raises(IndexError):
        st.write_trajectories(0, 2, _rollout(1, 2, len(VARS), 2, 2, 0))
    with pytest.raises(ValueError):
        st.write_trajectories(0, 0, _rollout(1, 3, len(VARS), 2, 2, 0))


def test_empty_store_and_single_lead(tmp_path):
    """Edge cases: no initial conditions at all (an empty --samples selection) and a zero-step rollout (lead 0 only)."""
    st = ForecastStore.create(str(tmp_path / "e.zarr"), VARS, 0, 3, 2, np.arange(4), np.arange(4))
    assert st.read("z").shape == (0, 3, 3, 3, 4, 4) and st.read_all().shape == (0, 3, 3, len(VARS), 4, 4)
    assert st.coordinate("time").shape == (0,)
    one = ForecastStore.create(str(tmp_path / "o.zarr"), VARS, 1, 1, 0, np.arange(4), np.arange(4), layout="step")
    r = _rollout(1, 0, len(VARS), 4, 4, seed=1)
    one.write_trajectories(0, 0, r)
    np.testing.assert_array_equal(one.read_all()[0, 0], r[0])


def test_round_trip_property():
    """Any (layout, batch, ragged n_ic, variable mix): what is written per (ic, member) block is what read_all returns, and
    every chunk file has exactly the size the .zarray metadata promises."""
    import tempfile
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=25, deadline=None)
    @given(layout=hs.sampled_from(["trajectory", "step"]), batch=hs.integers(1, 3), n_ic=hs.integers(1, 5),
           members=hs.integers(1, 3), steps=hs.integers(0, 3), n_lev=hs.integers(0, 3), n_single=hs.integers(0, 2),
           level=hs.sampled_from([0, 1]), seed=hs.integers(0, 10_000))
    def run(layout, batch, n_ic, members, steps, n_lev, n_single, level, seed):
        variables = [f"s{i}" for i in range(n_single)] + [f"p_{100 * (k + 1)}" for k in range(n_lev)]
        if not variables:
            variables = ["only"]
        h, w = 3, 4
        with tempfile.TemporaryDirectory() as d:
            st = ForecastStore.create(os.path.join(d, "f.zarr"), variables, n_ic, members, steps, np.arange(h), np.arange(w),
                                      layout=layout, batch=batch, compress_level=level)
            full = _rollout(n_ic * members, steps, len(variables), h, w, seed).reshape(n_ic, members, steps + 1,
                                                                                       len(variables), h, w)
            b = st.batch
            for m in range(members):
                for s0 in range(0, n_ic, b):
                    st.write_trajectories(s0, m, full[s0:s0 + b, m])
            np.testing.assert_array_equal(st.read_all(), full)
            if level == 0:
                for var in st.channels:
                    want = int(np.prod(st.chunk_shape(var))) * 4
                    files = [f for f in os.listdir(os.path.join(d, "f.zarr", var)) if not f.startswith(".")]
                    assert files and all(os.path.getsize(os.path.join(d, "f.zarr", var, f)) == want for f in files)

    run()
