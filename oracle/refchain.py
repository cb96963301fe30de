"""ctypes driver for oracle/_ref/libuhsdr_ref.so -- ORACLE / TEST INFRASTRUCTURE ONLY.

libuhsdr_ref.so is the reference's own RX/TX block path (audio_driver.c, audio_agc.c, audio_nr.c,
tx_processor.c, freq_shift.c, CMSIS-DSP portable kernels ...) compiled unmodified from
/root/reference by oracle/Makefile.  The reference keeps all DSP state in statics, so one loaded
copy of the library is exactly one channel: `RefChannel` copies the .so to a unique temporary
file before dlopen() to get a private set of statics, and unloads it on close().
"""
from __future__ import annotations

import ctypes
import os
import shutil
import tempfile

import numpy as np

from uhsdr_b200.config import ChanCfg, ChanStatus, SpectrumDisplayCfg

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libuhsdr_ref.so")


def available() -> bool:
    return os.path.exists(REF_SO)


class RefChannel:
    """One fresh reference channel (fresh-process semantics)."""

    def __init__(self, cfg: ChanCfg):
        if not available():
            raise FileNotFoundError(f"{REF_SO} not built (run `make -C oracle ref` where /root/reference exists)")
        fd, self._path = tempfile.mkstemp(prefix="uhsdr_ref_", suffix=".so")
        os.close(fd)
        shutil.copyfile(REF_SO, self._path)
        self._lib = ctypes.CDLL(self._path)
        os.unlink(self._path)  # mapping stays valid; nothing left behind
        L = self._lib
        L.ref_init.argtypes = [ctypes.POINTER(ChanCfg)]
        L.ref_reconfigure.argtypes = [ctypes.POINTER(ChanCfg)]
        L.ref_rx.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ref_get_status.argtypes = [ctypes.POINTER(ChanStatus)]
        rc = L.ref_init(ctypes.byref(cfg))
        if rc != 0:
            raise RuntimeError(f"ref_init failed: {rc}")

    def reconfigure(self, cfg: ChanCfg) -> None:
        rc = self._lib.ref_reconfigure(ctypes.byref(cfg))
        if rc != 0:
            raise RuntimeError(f"ref_reconfigure failed: {rc}")

    def rx(self, iq: np.ndarray, mute: np.ndarray | None = None):
        """iq: int32 [nsamples, 2] (l=I, r=Q), nsamples % 32 == 0.
        Returns (audio int32 [nsamples, 2], audio_f float32 [nsamples])."""
        iq = np.ascontiguousarray(iq, dtype=np.int32)
        n = iq.shape[0]
        assert iq.ndim == 2 and iq.shape[1] == 2 and n % 32 == 0
        audio = np.empty((n, 2), dtype=np.int32)
        audio_f = np.empty(n, dtype=np.float32)
        mp = None
        if mute is not None:
            mute = np.ascontiguousarray(mute, dtype=np.uint8)
            assert mute.size == n // 32
            mp = mute.ctypes.data
        rc = self._lib.ref_rx(iq.ctypes.data, audio.ctypes.data, audio_f.ctypes.data, n // 32, mp)
        if rc != 0:
            raise RuntimeError(f"ref_rx failed: {rc}")
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
        self._lib.ref_tx.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        rc = self._lib.ref_tx(mic.ctypes.data, iq.ctypes.data, iq_f.ctypes.data, n // 32, mp)
        if rc != 0:
            raise RuntimeError(f"ref_tx failed: {rc}")
        return iq, iq_f

    def spectrum(self) -> np.ndarray:
        """UiSpectrum_RedrawSpectrum states 0-2: 512 magnitudes of the current spectrum ring."""
        from uhsdr_b200.tables import Tables
        t = Tables()
        win = np.ascontiguousarray(t.arrays[t.extras["spectrum_window_array"]], dtype=np.float32)
        mags = np.empty(512, dtype=np.float32)
        self._lib.ref_spectrum.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        rc = self._lib.ref_spectrum(win.ctypes.data, mags.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ref_spectrum failed: {rc}")
        return mags

    def spectrum_display_init(self, dc: SpectrumDisplayCfg) -> None:
        """UiSpectrum_InitSpectrumDisplayData (the reference's own, ui_spectrum.c:955-1083) with the given settings."""
        self._lib.ref_spectrum_display_init.argtypes = [ctypes.c_int] * 5
        rc = self._lib.ref_spectrum_display_init(dc.spectrum_db_scale, dc.spectrum_agc_rate, dc.spectrum_filter, dc.dbm_constant, dc.scope_width)
        if rc != 0:
            raise RuntimeError(f"ref_spectrum_display_init failed: {rc}")
        self._scope_w = dc.scope_width

    def spectrum_display(self):
        """The reference's UiSpectrum_RedrawSpectrum, states 0-4: (mags[512], avg[512], disp[scope_width], (dbm, dbmhz, offset))."""
        mags, avg = np.empty(512, dtype=np.float32), np.empty(512, dtype=np.float32)
        disp, lvl = np.empty(self._scope_w, dtype=np.float32), np.empty(3, dtype=np.float32)
        self._lib.ref_spectrum_redraw.argtypes = [ctypes.c_void_p] * 4
        rc = self._lib.ref_spectrum_redraw(mags.ctypes.data, avg.ctypes.data, disp.ctypes.data, lvl.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ref_spectrum_redraw failed: {rc}")
        return mags, avg, disp, lvl

    def twinpeaks_rearm(self) -> None:
        self._lib.ref_twinpeaks_rearm()

    def status(self) -> ChanStatus:
        st = ChanStatus()
        self._lib.ref_get_status(ctypes.byref(st))
        return st

    def close(self) -> None:
        if self._lib is not None:
            handle = self._lib._handle
            self._lib = None
            import _ctypes
            _ctypes.dlclose(handle)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
