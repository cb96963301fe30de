"""ctypes binding of the CUDA library (include/uhsdr_b200.h) -- the host-side mirror of the
reference's block interface:

    AudioDriver_SetProcessingChain(dmod_mode, reset)  -> Engine.configure(first, count, cfg, reset)
    AudioDriver_RxProcessor(iq, audio, 32, mute)      -> Engine.rx(iq[nch, n, 2]) -> audio[nch, n, 2]
    TxProcessor_Run(audio, iq, ...)                   -> Engine.tx(mic[nch, n, 2]) -> iq[nch, n, 2]

All arithmetic happens in libuhsdr_b200.so on the GPU.  There is no CPU fallback: if the library
is missing or no B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .config import BLOCK_SIZE, ChanCfg, ChanStatus, SpectrumDisplayCfg, SpectrumLevel
from .tables import DEFAULT_BLOB

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_FAST = os.path.join(_HERE, "csrc", "libuhsdr_b200.so")
LIB_EXACT = os.path.join(_HERE, "csrc", "libuhsdr_b200_exact.so")

OK = 0
ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_TABLES, ERR_UNSUPPORTED, ERR_STATE = -1, -2, -3, -4, -5, -6

EXPORTS = [
    "uhsdr_b200_abi_version", "uhsdr_b200_backend", "uhsdr_strerror", "uhsdr_last_error",
    "uhsdr_default_chan_cfg", "uhsdr_engine_create", "uhsdr_engine_destroy", "uhsdr_engine_num_channels",
    "uhsdr_configure_channels", "uhsdr_configure_channel", "uhsdr_configure_channels_strided", "uhsdr_rx_process", "uhsdr_rx_process_device",
    "uhsdr_tx_process", "uhsdr_tx_process_device", "uhsdr_engine_sync", "uhsdr_engine_stream",
    "uhsdr_get_spectrum", "uhsdr_get_spectrum_device", "uhsdr_get_status", "uhsdr_engine_launch_count",
    "uhsdr_channel_range", "uhsdr_multi_create", "uhsdr_multi_destroy", "uhsdr_multi_num_devices", "uhsdr_multi_num_channels",
    "uhsdr_multi_last_error", "uhsdr_multi_engine", "uhsdr_multi_configure_channels", "uhsdr_multi_configure_channels_strided",
    "uhsdr_multi_rx_process", "uhsdr_multi_tx_process", "uhsdr_multi_get_status",
    "uhsdr_twinpeaks_rearm", "uhsdr_tables_validate", "uhsdr_default_spectrum_display_cfg", "uhsdr_spectrum_display", "uhsdr_spectrum_display_device",
]

_libs: dict[str, ctypes.CDLL] = {}


class UhsdrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"uhsdr_b200 error {code}: {msg}")
        self.code = code


def load_library(exact: bool = False) -> ctypes.CDLL:
    path = LIB_EXACT if exact else LIB_FAST
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: build it with `make -C uhsdr_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    L = ctypes.CDLL(path)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.uhsdr_b200_backend.restype = ctypes.c_char_p
    L.uhsdr_strerror.restype = ctypes.c_char_p
    L.uhsdr_strerror.argtypes = [ci]
    L.uhsdr_last_error.restype = ctypes.c_char_p
    L.uhsdr_last_error.argtypes = [vp]
    L.uhsdr_default_chan_cfg.argtypes = [ctypes.POINTER(ChanCfg)]
    L.uhsdr_engine_create.argtypes = [ctypes.POINTER(vp), ci, ci, vp, ctypes.c_size_t]
    L.uhsdr_engine_destroy.argtypes = [vp]
    L.uhsdr_tables_validate.argtypes = [vp, ctypes.c_size_t]
    L.uhsdr_engine_num_channels.argtypes = [vp]
    L.uhsdr_configure_channels.argtypes = [vp, ci, ci, ctypes.POINTER(ChanCfg), ci]
    L.uhsdr_configure_channel.argtypes = [vp, ci, ctypes.POINTER(ChanCfg), ci]
    L.uhsdr_configure_channels_strided.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ChanCfg), ci]
    L.uhsdr_rx_process.argtypes = [vp, vp, vp, ci, vp]
    L.uhsdr_rx_process_device.argtypes = [vp, vp, vp, vp, ci, vp]
    L.uhsdr_tx_process.argtypes = [vp, vp, vp, ci, vp]
    L.uhsdr_tx_process_device.argtypes = [vp, vp, vp, vp, ci, vp]
    L.uhsdr_engine_sync.argtypes = [vp]
    L.uhsdr_engine_stream.restype = vp
    L.uhsdr_engine_stream.argtypes = [vp]
    L.uhsdr_get_spectrum.argtypes = [vp, ci, ci, vp]
    L.uhsdr_get_spectrum_device.argtypes = [vp, ci, ci, vp]
    L.uhsdr_get_status.argtypes = [vp, ci, ci, ctypes.POINTER(ChanStatus)]
    L.uhsdr_twinpeaks_rearm.argtypes = [vp, ci, ci]
    L.uhsdr_channel_range.argtypes = [ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.uhsdr_multi_create.argtypes = [ctypes.POINTER(vp), ci, ctypes.POINTER(ci), ci, vp, ctypes.c_size_t]
    L.uhsdr_multi_destroy.argtypes = [vp]
    L.uhsdr_multi_num_devices.argtypes = [vp]
    L.uhsdr_multi_num_channels.argtypes = [vp]
    L.uhsdr_multi_last_error.restype = ctypes.c_char_p
    L.uhsdr_multi_last_error.argtypes = [vp]
    L.uhsdr_multi_engine.restype = vp
    L.uhsdr_multi_engine.argtypes = [vp, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.uhsdr_multi_configure_channels.argtypes = [vp, ci, ci, ctypes.POINTER(ChanCfg), ci]
    L.uhsdr_multi_configure_channels_strided.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ChanCfg), ci]
    L.uhsdr_multi_rx_process.argtypes = [vp, vp, vp, ci, vp]
    L.uhsdr_multi_tx_process.argtypes = [vp, vp, vp, ci, vp]
    L.uhsdr_multi_get_status.argtypes = [vp, ci, ci, ctypes.POINTER(ChanStatus)]
    L.uhsdr_default_spectrum_display_cfg.argtypes = [ctypes.POINTER(SpectrumDisplayCfg)]
    L.uhsdr_spectrum_display.argtypes = [vp, ci, ci, ctypes.POINTER(SpectrumDisplayCfg), vp, vp, vp]
    L.uhsdr_spectrum_display_device.argtypes = [vp, ci, ci, ctypes.POINTER(SpectrumDisplayCfg), vp, vp, vp, vp]
    L.uhsdr_engine_launch_count.restype = ctypes.c_int64
    L.uhsdr_engine_launch_count.argtypes = [vp]
    _libs[path] = L
    return L


def _ptr(x) -> int | None:
    """Device/host address of a numpy array, a torch tensor or an int; None stays None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class Engine:
    """A batch of `num_channels` independent receiver/transmitter channels on one GPU."""

    def __init__(self, num_channels: int, device: int = 0, tables: bytes | None = None, exact: bool = False):
        self._lib = load_library(exact)
        if tables is None:
            with open(DEFAULT_BLOB, "rb") as f:
                tables = f.read()
        self._h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(tables, len(tables))
        rc = self._lib.uhsdr_engine_create(ctypes.byref(self._h), num_channels, device, buf, len(tables))
        if rc != OK:
            msg = self._lib.uhsdr_last_error(None).decode()
            self._h = None
            raise UhsdrError(rc, msg)
        self.num_channels = num_channels
        self.device = device

    def _check(self, rc: int) -> None:
        if rc != OK:
            raise UhsdrError(rc, self._lib.uhsdr_strerror(rc).decode() + ": " + self._lib.uhsdr_last_error(self._h).decode())

    @property
    def backend(self) -> str:
        return self._lib.uhsdr_b200_backend().decode()

    def configure(self, cfg: ChanCfg, first: int = 0, count: int | None = None, reset: bool = True, stride: int = 1) -> None:
        """AudioDriver_SetProcessingChain for channels first, first+stride, ... (count of them)."""
        if count is None:
            count = (self.num_channels - first + stride - 1) // stride
        self._check(self._lib.uhsdr_configure_channels_strided(self._h, first, count, stride, ctypes.byref(cfg), 1 if reset else 0))

    def rx(self, iq: np.ndarray, mute: np.ndarray | None = None) -> np.ndarray:
        """iq: int32 [num_channels, nsamples, 2] host array -> audio int32 of the same shape."""
        iq = np.ascontiguousarray(iq, dtype=np.int32)
        assert iq.ndim == 3 and iq.shape[0] == self.num_channels and iq.shape[2] == 2 and iq.shape[1] % BLOCK_SIZE == 0
        audio = np.empty_like(iq)
        if mute is not None:
            mute = np.ascontiguousarray(mute, dtype=np.uint8)
            assert mute.shape == (self.num_channels, iq.shape[1] // BLOCK_SIZE)
        self._check(self._lib.uhsdr_rx_process(self._h, iq.ctypes.data, audio.ctypes.data, iq.shape[1] // BLOCK_SIZE, _ptr(mute)))
        return audio

    def rx_device(self, iq_dev, audio_dev, nblocks: int, audio_f_dev=None, mute_dev=None) -> None:
        """Device-pointer variant (torch tensors or raw addresses); asynchronous on the engine stream."""
        self._check(self._lib.uhsdr_rx_process_device(self._h, _ptr(iq_dev), _ptr(audio_dev), _ptr(audio_f_dev), nblocks, _ptr(mute_dev)))

    def tx(self, mic: np.ndarray, mute: np.ndarray | None = None) -> np.ndarray:
        mic = np.ascontiguousarray(mic, dtype=np.int32)
        assert mic.ndim == 3 and mic.shape[0] == self.num_channels and mic.shape[2] == 2 and mic.shape[1] % BLOCK_SIZE == 0
        iq = np.empty_like(mic)
        if mute is not None:
            mute = np.ascontiguousarray(mute, dtype=np.uint8)
        self._check(self._lib.uhsdr_tx_process(self._h, mic.ctypes.data, iq.ctypes.data, mic.shape[1] // BLOCK_SIZE, _ptr(mute)))
        return iq

    def tx_device(self, audio_dev, iq_dev, nblocks: int, iq_f_dev=None, mute_dev=None) -> None:
        self._check(self._lib.uhsdr_tx_process_device(self._h, _ptr(audio_dev), _ptr(iq_dev), _ptr(iq_f_dev), nblocks, _ptr(mute_dev)))

    def spectrum(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.num_channels - first if count is None else count
        mags = np.empty((count, 512), dtype=np.float32)
        self._check(self._lib.uhsdr_get_spectrum(self._h, first, count, mags.ctypes.data))
        return mags

    def spectrum_display(self, dc: SpectrumDisplayCfg, first: int = 0, count: int | None = None):
        """UiSpectrum_RedrawSpectrum states 0-4: (disp [count, scope_width], levels [count, 3] = dBm, dBm/Hz, display offset,
        avg [count, 512])."""
        count = self.num_channels - first if count is None else count
        disp = np.empty((count, dc.scope_width), dtype=np.float32)
        lvl = np.empty((count, 3), dtype=np.float32)
        avg = np.empty((count, 512), dtype=np.float32)
        self._check(self._lib.uhsdr_spectrum_display(self._h, first, count, ctypes.byref(dc), disp.ctypes.data, lvl.ctypes.data, avg.ctypes.data))
        return disp, lvl, avg

    def status(self, first: int = 0, count: int | None = None) -> list[ChanStatus]:
        count = self.num_channels - first if count is None else count
        arr = (ChanStatus * count)()
        self._check(self._lib.uhsdr_get_status(self._h, first, count, arr))
        return list(arr)

    def twinpeaks_rearm(self, first: int = 0, count: int | None = None) -> None:
        count = self.num_channels - first if count is None else count
        self._check(self._lib.uhsdr_twinpeaks_rearm(self._h, first, count))

    def sync(self) -> None:
        self._check(self._lib.uhsdr_engine_sync(self._h))

    @property
    def stream(self) -> int:
        return self._lib.uhsdr_engine_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return int(self._lib.uhsdr_engine_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.uhsdr_engine_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
