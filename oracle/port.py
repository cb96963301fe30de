"""ctypes driver for oracle/liboracle_port.so (the plain-C restatement) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from uhsdr_b200.config import ChanCfg, ChanStatus, SpectrumDisplayCfg
from uhsdr_b200.tables import DEFAULT_BLOB

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "liboracle_port.so")
_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            build()
        L = ctypes.CDLL(PORT_SO)
        L.port_tables_load.restype = ctypes.c_void_p
        L.port_tables_load.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.port_tables_free.argtypes = [ctypes.c_void_p]
        L.port_chan_create.restype = ctypes.c_void_p
        L.port_chan_create.argtypes = [ctypes.c_void_p, ctypes.POINTER(ChanCfg)]
        L.port_chan_reconfigure.argtypes = [ctypes.c_void_p, ctypes.POINTER(ChanCfg)]
        L.port_chan_free.argtypes = [ctypes.c_void_p]
        L.port_rx.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
        L.port_tx.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
        L.port_spectrum.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.port_spectrum_display.argtypes = [ctypes.c_void_p, ctypes.POINTER(SpectrumDisplayCfg)] + [ctypes.c_void_p] * 4
        L.port_get_status.argtypes = [ctypes.c_void_p, ctypes.POINTER(ChanStatus)]
        _lib = L
    return _lib


class PortTables:
    def __init__(self, blob: bytes | None = None):
        if blob is None:
            with open(DEFAULT_BLOB, "rb") as f:
                blob = f.read()
        self._buf = ctypes.create_string_buffer(blob, len(blob))
        self.handle = lib().port_tables_load(self._buf, len(blob))
        if not self.handle:
            raise ValueError("bad table blob")

    def __del__(self):
        if getattr(self, "handle", None):
            lib().port_tables_free(self.handle)
            self.handle = None


_default_tables = None


def default_tables() -> PortTables:
    global _default_tables
    if _default_tables is None:
        _default_tables = PortTables()
    return _default_tables


class PortChannel:
    def __init__(self, cfg: ChanCfg, tables: PortTables | None = None):
        self._tables = tables or default_tables()
        self._h = lib().port_chan_create(self._tables.handle, ctypes.byref(cfg))
        if not self._h:
            raise ValueError("port_chan_create: unsupported configuration")

    def reconfigure(self, cfg: ChanCfg) -> None:
        rc = lib().port_chan_reconfigure(self._h, ctypes.byref(cfg))
        if rc != 0:
            raise ValueError(f"port_chan_reconfigure: {rc}")

    def rx(self, iq: np.ndarray, mute: np.ndarray | None = None):
        iq = np.ascontiguousarray(iq, dtype=np.int32)
        n = iq.shape[0]
        assert iq.ndim == 2 and iq.shape[1] == 2 and n % 32 == 0
        audio = np.empty((n, 2), dtype=np.int32)
        audio_f = np.empty(n, dtype=np.float32)
        mp = None
        if mute is not None:
            mute = np.ascontiguousarray(mute, dtype=np.uint8)
            mp = mute.ctypes.data
        rc = lib().port_rx(self._h, iq.ctypes.data, audio.ctypes.data, audio_f.ctypes.data, n // 32, mp)
        if rc != 0:
            raise RuntimeError(f"port_rx: {rc}")
        return audio, audio_f

    def tx(self, mic: np.ndarray, mute: np.ndarray | None = None):
        mic = np.ascontiguousarray(mic, dtype=np.int32)
        n = mic.shape[0]
        iq = np.empty((n, 2), dtype=np.int32)
        iq_f = np.empty((n, 2), dtype=np.float32)
        mp = None
        if mute is not None:
            mute = np.ascontiguousarray(mute, dtype=np.uint8)
            mp = mute.ctypes.data
        rc = lib().port_tx(self._h, mic.ctypes.data, iq.ctypes.data, iq_f.ctypes.data, n // 32, mp)
        if rc != 0:
            raise RuntimeError(f"port_tx: {rc}")
        return iq, iq_f

    def spectrum(self) -> np.ndarray:
        mags = np.empty(512, dtype=np.float32)
        rc = lib().port_spectrum(self._h, mags.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"port_spectrum: {rc}")
        return mags

    def spectrum_display(self, dc: SpectrumDisplayCfg):
        """UiSpectrum_RedrawSpectrum states 0-4: (mags[512], avg[512], disp[scope_width], (dbm, dbmhz, display_offset))."""
        mags, avg = np.empty(512, dtype=np.float32), np.empty(512, dtype=np.float32)
        disp, lvl = np.empty(dc.scope_width, dtype=np.float32), np.empty(3, dtype=np.float32)
        rc = lib().port_spectrum_display(self._h, ctypes.byref(dc), mags.ctypes.data, avg.ctypes.data, disp.ctypes.data, lvl.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"port_spectrum_display: {rc}")
        return mags, avg, disp, lvl

    def twinpeaks_rearm(self) -> None:
        lib().port_twinpeaks_rearm.argtypes = [ctypes.c_void_p]
        lib().port_twinpeaks_rearm(self._h)

    def status(self) -> ChanStatus:
        st = ChanStatus()
        lib().port_get_status(self._h, ctypes.byref(st))
        return st

    def close(self) -> None:
        if self._h:
            lib().port_chan_free(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        self.close()
