"""Forecast store: the on-disk layout ``generate.py`` writes the rollout into, without the zarr / xarray / dask stack.

stockeh/swift writes every rollout into a zarr group with one array per (level-compressed) variable,

    <var>[time, number, prediction_timedelta, (level,) latitude, longitude]   float32

created by ``utils/io.py:85-157`` (``create_empty_zarr``: chunks (1, 1, 1, [L], lat, lon)) or ``utils/io.py:160-231``
(``fast_create_empty_zarr``: chunks (batch, 1, steps+1, [L], lat, lon)) and filled by ``generate.py:139-152``; the
alternative ``--dump numpy`` is one ``.npy`` memmap [samples, members, steps+1, channels, H, W] (``utils/io.py:237-260``,
``generate.py:140-142``).  This module produces the same two layouts directly:

  * a zarr **format-2 directory store** (``.zgroup`` / ``.zarray`` / ``.zattrs`` / ``.zmetadata`` JSON + one raw
    little-endian C-order file per chunk, key ``i.j.k...``), which ``zarr.open_group`` / ``xr.open_zarr(path,
    decode_timedelta=True)`` read as they read the reference's output; coordinates carry the CF attributes xarray decodes
    (``time``: "hours since ...", ``prediction_timedelta``: "hours");
  * the ``.npy`` memmap.

A chunk never spans two (time, number) pairs in either reference layout with batch = 1, so ranks that own different
trajectories write different files and need no lock -- the same property the reference relies on when every rank opens
the group in mode "a" (``generate.py:58-60``).

Neither zarr nor xarray exists in this image, so the layout is pinned by the format-2 specification
and by round-trip tests with the reader below (``tests/test_store_cpu.py``), not by the libraries themselves.
"""
from __future__ import annotations

import json
import os
import re
import threading
import zlib
from collections import defaultdict
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def compress_variables(variables: Iterable[str]) -> Dict[str, List[int]]:
    """``geopotential_500`` -> {"geopotential": [.., 500]}; names without a numeric suffix map to [] (utils/io.py:71-82)."""
    out: Dict[str, List[int]] = defaultdict(list)
    for var in variables:
        m = re.match(r"^(.*)_(\d+)$", var)
        if m:
            out[m.group(1)].append(int(m.group(2)))
        else:
            out[var] = []
    return dict(out)


def variable_channels(variables: Sequence[str]) -> Dict[str, List[int]]:
    """Channel indices of every compressed variable in the state tensor, assigned in order of first appearance with the
    levels of a variable taken as one contiguous run -- exactly the counter of generate.py:63-74."""
    idx: Dict[str, List[int]] = {}
    c = 0
    for var, levels in compress_variables(variables).items():
        n = len(levels) if levels else 1
        idx[var] = list(range(c, c + n))
        c += n
    return idx


def _write_json(path: str, obj) -> None:
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
    os.replace(tmp, path)


def _zarray(shape, chunks, dtype: str, fill, compressor) -> dict:
    return {"zarr_format": 2, "shape": [int(s) for s in shape], "chunks": [int(c) for c in chunks], "dtype": dtype,
            "fill_value": fill, "order": "C", "filters": None, "compressor": compressor, "dimension_separator": "."}


class ForecastStore:
    """Writer / reader of one rollout output.

    ``layout``: "trajectory" = fast_create_empty_zarr (one chunk per (IC block, member) holding every lead time),
    "step" = create_empty_zarr (one chunk per (IC, member, lead time): can be written while the rollout runs),
    "numpy" = the .npy memmap.
    """

    def __init__(self, path: str, layout: str, variables: Sequence[str], n_ic: int, members: int, steps: int,
                 resolution: Tuple[int, int], batch: int = 1, compress_level: int = 0):
        if layout not in ("trajectory", "step", "numpy"):
            raise ValueError(f"Unknown store layout: {layout}")
        self.path, self.layout = path, layout
        self.variables = list(variables)
        self.channels = variable_channels(self.variables)
        self.levels = compress_variables(self.variables)
        self.n_ic, self.members, self.steps = int(n_ic), int(members), int(steps)
        self.res = (int(resolution[0]), int(resolution[1]))
        self.batch = int(batch) if layout == "trajectory" else 1
        self.compress_level = int(compress_level)
        self._mm: Optional[np.memmap] = None
        self._mm_lock = threading.Lock()

    # ------------------------------------------------------------------------------------------------ creation
    @classmethod
    def create(cls, path: str, variables: Sequence[str], n_ic: int, members: int, steps: int, lat: np.ndarray,
               lon: np.ndarray, times: Optional[np.ndarray] = None, interval_hours: int = 6, layout: str = "trajectory",
               batch: int = 1, compress_level: int = 0) -> "ForecastStore":
        """Rank 0 calls this once before the rollout (generate.py:272-282 via run_on_rank0); every rank then ``open``s."""
        st = cls(path, layout, variables, n_ic, members, steps, (len(lat), len(lon)), batch, compress_level)
        if layout == "numpy":
            n_ch = sum(len(v) for v in st.channels.values())
            np.lib.format.open_memmap(path, dtype=np.float32, mode="w+",
                                      shape=(st.n_ic, st.members, st.steps + 1, n_ch, *st.res)).flush()
            return st
        os.makedirs(path, exist_ok=True)
        meta: Dict[str, dict] = {}

        def put(rel: str, obj) -> None:
            full = os.path.join(path, rel)
            os.makedirs(os.path.dirname(full), exist_ok=True)
            _write_json(full, obj)
            meta[rel] = obj

        put(".zgroup", {"zarr_format": 2})
        put(".zattrs", {"layout": layout, "variables": st.variables})
        if times is None:
            times = np.datetime64("2020-01-01T00", "h") + np.arange(st.n_ic) * np.timedelta64(12, "h")
        times = np.asarray(times).astype("datetime64[h]")
        t0 = times[0] if len(times) else np.datetime64("1970-01-01T00", "h")
        n_lev = max((len(v) for v in st.levels.values()), default=0)
        coords = {
            "time": ((times - t0).astype(np.int64), {"units": "hours since " + str(t0).replace("T", " ") + ":00:00",
                                                     "calendar": "proleptic_gregorian"}),
            "number": (np.arange(st.members, dtype=np.int64), {}),
            "prediction_timedelta": (np.arange(st.steps + 1, dtype=np.int64) * int(interval_hours), {"units": "hours"}),
            "latitude": (np.asarray(lat, dtype=np.float64), {}),
            "longitude": (np.asarray(lon, dtype=np.float64), {}),
        }
        if n_lev:
            coords["level"] = (np.arange(n_lev, dtype=np.int64), {})     # an index, as in utils/io.py:119-120
        for name, (vals, attrs) in coords.items():
            dt = vals.dtype.str
            put(f"{name}/.zarray", _zarray(vals.shape, (max(1, len(vals)),), dt, 0 if dt[1] == "i" else "NaN", None))
            put(f"{name}/.zattrs", {"_ARRAY_DIMENSIONS": [name], **attrs})
            if len(vals):
                np.ascontiguousarray(vals).tofile(os.path.join(path, name, "0"))
        comp = {"id": "zlib", "level": st.compress_level} if st.compress_level > 0 else None
        for var in st.channels:
            put(f"{var}/.zarray", _zarray(st.array_shape(var), st.chunk_shape(var), "<f4", 0.0, comp))
            put(f"{var}/.zattrs", {"_ARRAY_DIMENSIONS": st.dims(var)})
        _write_json(os.path.join(path, ".zmetadata"), {"zarr_consolidated_format": 1, "metadata": meta})
        return st

    @classmethod
    def open(cls, path: str) -> "ForecastStore":
        if os.path.isfile(path):
            mm = np.load(path, mmap_mode="r")
            n_ic, members, s1, n_ch, h, w = mm.shape
            return cls(path, "numpy", [f"channel{c}" for c in range(n_ch)], n_ic, members, s1 - 1, (h, w))
        with open(os.path.join(path, ".zattrs")) as f:
            attrs = json.load(f)
        var0 = next(iter(variable_channels(attrs["variables"])))
        with open(os.path.join(path, var0, ".zarray")) as f:
            za = json.load(f)
        shape, chunks = za["shape"], za["chunks"]
        level = (za["compressor"] or {}).get("level", 0)
        return cls(path, attrs["layout"], attrs["variables"], shape[0], shape[1], shape[2] - 1,
                   (shape[-2], shape[-1]), chunks[0], level)

    # ------------------------------------------------------------------------------------------------ geometry
    def dims(self, var: str) -> List[str]:
        d = ["time", "number", "prediction_timedelta", "latitude", "longitude"]
        return d[:3] + ["level"] + d[3:] if self.levels[var] else d

    def array_shape(self, var: str) -> Tuple[int, ...]:
        lev = (len(self.levels[var]),) if self.levels[var] else ()
        return (self.n_ic, self.members, self.steps + 1) + lev + self.res

    def chunk_shape(self, var: str) -> Tuple[int, ...]:
        lev = (len(self.levels[var]),) if self.levels[var] else ()
        lead = (self.batch, 1, self.steps + 1) if self.layout == "trajectory" else (1, 1, 1)
        return lead + lev + self.res

    def _chunk_path(self, var: str, i: int, m: int, k: int) -> str:
        tail = ".0" * (len(self.array_shape(var)) - 3)
        return os.path.join(self.path, var, f"{i}.{m}.{k}{tail}")

    def _put_chunk(self, file: str, data: np.ndarray) -> None:
        data = np.ascontiguousarray(data, dtype="<f4")
        tmp = file + ".tmp%d" % os.getpid()
        if self.compress_level > 0:
            with open(tmp, "wb") as f:
                f.write(zlib.compress(data.tobytes(), self.compress_level))
        else:
            data.tofile(tmp)
        os.replace(tmp, file)

    def _select(self, var: str, fields: np.ndarray, axis: int) -> np.ndarray:
        """Channels of ``var`` out of ``fields`` (channel axis ``axis``): a single field, or the levels stacked on that
        axis (generate.py:144-151)."""
        ch = self.channels[var]
        if self.levels[var]:
            return np.take(fields, ch, axis=axis)
        return np.take(fields, ch[0], axis=axis)

    # ------------------------------------------------------------------------------------------------ writing
    def write_trajectories(self, ic_start: int, member: int, rollout: np.ndarray) -> None:
        """``rollout`` [bs, steps+1, channels, H, W] physical fields of ICs ic_start .. ic_start+bs-1 for one member: the
        ``store[var][n:n+bs, m] = ...`` of generate.py:139-152."""
        bs = rollout.shape[0]
        if rollout.shape[1] != self.steps + 1 or tuple(rollout.shape[3:]) != self.res:
            raise ValueError(f"rollout has shape {rollout.shape}, store expects [bs, {self.steps + 1}, C, {self.res}]")
        if not (0 <= member < self.members and 0 <= ic_start and ic_start + bs <= self.n_ic):
            raise IndexError("trajectory block outside the store")
        if self.layout == "numpy":
            self._memmap()[ic_start:ic_start + bs, member] = rollout
            return
        if self.layout == "step":
            for b in range(bs):
                for k in range(self.steps + 1):
                    self.write_step(ic_start + b, member, k, rollout[b, k])
            return
        if ic_start % self.batch or bs > self.batch:
            raise ValueError(f"trajectory layout with batch={self.batch}: a write must be one chunk row "
                             f"(ic_start multiple of {self.batch}, at most {self.batch} ICs)")
        for var in self.channels:
            data = self._select(var, rollout, axis=2)                       # [bs, steps+1, (L,) H, W]
            if bs < self.batch:                                              # edge chunk: zarr stores full chunks
                pad = np.zeros((self.batch - bs,) + data.shape[1:], dtype=np.float32)
                data = np.concatenate([data, pad], axis=0)
            self._put_chunk(self._chunk_path(var, ic_start // self.batch, member, 0), data[:, None])

    def write_step(self, ic: int, member: int, lead: int, fields: np.ndarray) -> None:
        """One lead time of one trajectory, ``fields`` [channels, H, W] ("step" and "numpy" layouts: usable while the
        rollout is running, which the reference's per-batch host array cannot do)."""
        if self.layout == "trajectory":
            raise ValueError("the trajectory layout holds every lead time in one chunk: use write_trajectories")
        if not (0 <= ic < self.n_ic and 0 <= member < self.members and 0 <= lead <= self.steps):
            raise IndexError("(ic, member, lead) outside the store")
        if self.layout == "numpy":
            self._memmap()[ic, member, lead] = fields
            return
        for var in self.channels:
            self._put_chunk(self._chunk_path(var, ic, member, lead), self._select(var, fields, axis=0))

    def _memmap(self) -> np.memmap:
        if self._mm is None:
            with self._mm_lock:                                   # writer threads share one mapping
                if self._mm is None:
                    self._mm = np.lib.format.open_memmap(self.path, mode="r+")
        return self._mm

    def flush(self) -> None:
        if self._mm is not None:
            self._mm.flush()

    # ------------------------------------------------------------------------------------------------ reading
    def read(self, var: str) -> np.ndarray:
        """The full array of one variable (missing chunks read as the fill value 0, as zarr does)."""
        if self.layout == "numpy":
            raise ValueError("numpy layout has no variables: use read_all()")
        shape, chunks = self.array_shape(var), self.chunk_shape(var)
        out = np.zeros(shape, dtype=np.float32)
        with open(os.path.join(self.path, var, ".zarray")) as f:
            za = json.load(f)
        if tuple(za["shape"]) != shape or tuple(za["chunks"]) != chunks:
            raise ValueError(f"{var}: metadata {za['shape']} / {za['chunks']} does not match {shape} / {chunks}")
        n = int(np.prod(chunks))
        for i in range(-(-shape[0] // chunks[0])):
            for m in range(shape[1]):
                for k in range(-(-shape[2] // chunks[2])):
                    file = self._chunk_path(var, i, m, k)
                    if not os.path.exists(file):
                        continue
                    raw = open(file, "rb").read()
                    if za["compressor"]:
                        raw = zlib.decompress(raw)
                    blk = np.frombuffer(raw, dtype="<f4", count=n).reshape(chunks)
                    i0, k0 = i * chunks[0], k * chunks[2]
                    ni, nk = min(chunks[0], shape[0] - i0), min(chunks[2], shape[2] - k0)
                    out[i0:i0 + ni, m, k0:k0 + nk] = blk[:ni, 0, :nk]
        return out

    def read_all(self) -> np.ndarray:
        """[n_ic, members, steps+1, channels, H, W] in the channel order of the state tensor."""
        if self.layout == "numpy":
            return np.load(self.path, mmap_mode="r")
        n_ch = sum(len(v) for v in self.channels.values())
        out = np.zeros((self.n_ic, self.members, self.steps + 1, n_ch, *self.res), dtype=np.float32)
        for var, ch in self.channels.items():
            a = self.read(var)
            if self.levels[var]:
                out[:, :, :, ch] = a
            else:
                out[:, :, :, ch[0]] = a
        return out

    def coordinate(self, name: str) -> np.ndarray:
        with open(os.path.join(self.path, name, ".zarray")) as f:
            za = json.load(f)
        file = os.path.join(self.path, name, "0")
        if not os.path.exists(file):                               # an empty axis has no chunk
            return np.zeros(za["shape"], dtype=za["dtype"])
        return np.fromfile(file, dtype=za["dtype"]).reshape(za["shape"])
