"""Synthetic I/Q and microphone test signals, numpy host generator.

Signals are int32 with the sample left-justified by 16 bits (value = round(x * 2^16), x in 16-bit
units), the format the reference's codec DMA delivers (audio_driver.h:594-610).  The wanted
carrier sits at -translate_freq in the input I/Q because the receiver shifts by translate_freq
(audio_driver.c:2694-2697): +12 kHz for the default FREQ_IQ_CONV_M12KHZ.
"""
from __future__ import annotations

import numpy as np

from .config import (DEMOD_AM, DEMOD_FM, DEMOD_LSB, DEMOD_SAM, FREQ_IQ_CONV_M12KHZ, FREQ_IQ_CONV_M6KHZ,
                     FREQ_IQ_CONV_OFF, FREQ_IQ_CONV_P12KHZ, FREQ_IQ_CONV_P6KHZ, ChanCfg)

FS = 48000.0
_TRANSLATE = {FREQ_IQ_CONV_OFF: 0.0, FREQ_IQ_CONV_P6KHZ: 6000.0, FREQ_IQ_CONV_M6KHZ: -6000.0,
              FREQ_IQ_CONV_P12KHZ: 12000.0, FREQ_IQ_CONV_M12KHZ: -12000.0}


def carrier_offset(cfg: ChanCfg) -> float:
    """Input frequency (Hz) at which the wanted carrier must sit for this configuration."""
    return -_TRANSLATE[cfg.iq_freq_mode]


def rx_iq(cfg: ChanCfg, channel: int, nsamples: int, seed: int = 0x55485344, start: int = 0,
          noise_sigma: float = 100.0) -> np.ndarray:
    """int32 [nsamples, 2] (l = I, r = Q) multi-tone + interferer + fading + AWGN for `cfg`'s mode."""
    n = np.arange(start, start + nsamples, dtype=np.float64)
    t = n / FS
    fc = carrier_offset(cfg)
    rng = np.random.default_rng([seed, channel, start])
    fade = 10.0 ** ((6.0 * np.sin(2 * np.pi * 0.5 * t + 0.1 * channel)) / 20.0)
    mode = cfg.dmod_mode
    if mode == DEMOD_FM:
        dev = 2500.0
        phase = 2 * np.pi * fc * t + (dev / 1000.0) * np.sin(2 * np.pi * 1000.0 * t)
        x = 4000.0 * np.exp(1j * phase)
    elif mode in (DEMOD_AM, DEMOD_SAM):
        env = 1.0 + 0.5 * np.sin(2 * np.pi * (1000.0 + 3.0 * (channel % 64)) * t)
        x = 4000.0 * fade * env * np.exp(2j * np.pi * (fc + 37.0) * t)
    else:
        sgn = -1.0 if mode == DEMOD_LSB or (mode == 2 and cfg.cw_lsb) or (mode == 6 and cfg.digi_lsb) else 1.0
        d = 3.0 * (channel % 64)
        x = np.zeros(nsamples, dtype=np.complex128)
        for fa, amp in ((700.0, 3000.0), (1500.0, 2000.0), (2100.0, 1000.0)):
            x += amp * np.exp(2j * np.pi * (fc + sgn * (fa + d)) * t)
        x *= fade
        x += 3000.0 * np.exp(2j * np.pi * (fc - sgn * 1500.0) * t)   # opposite-sideband interferer
    x = x + noise_sigma * (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples))
    out = np.empty((nsamples, 2), dtype=np.int32)
    out[:, 0] = np.round(x.real * 65536.0).astype(np.int64).clip(-2**31, 2**31 - 1)
    out[:, 1] = np.round(x.imag * 65536.0).astype(np.int64).clip(-2**31, 2**31 - 1)
    return out


def rx_fm_subtone_iq(cfg: ChanCfg, channel: int, nsamples: int, tone_hz: float, tone_dev_hz: float = 300.0,
                     seed: int = 0x55485344) -> np.ndarray:
    """FM test signal (1 kHz tone, 2.5 kHz deviation) with a sub-audible (CTCSS) tone of `tone_dev_hz` deviation on top --
    the input of the Goertzel tone detector (audio_driver.c:1665-1734).  tone_dev_hz = 0: no sub-audible tone."""
    t = np.arange(nsamples, dtype=np.float64) / FS
    rng = np.random.default_rng([seed, channel, 0x4354])
    phase = 2 * np.pi * carrier_offset(cfg) * t + 2.5 * np.sin(2 * np.pi * 1000.0 * t)
    if tone_dev_hz:
        phase = phase + (tone_dev_hz / tone_hz) * np.sin(2 * np.pi * tone_hz * t)
    x = 4000.0 * np.exp(1j * phase) + 30.0 * (rng.standard_normal(nsamples) + 1j * rng.standard_normal(nsamples))
    out = np.empty((nsamples, 2), dtype=np.int32)
    out[:, 0] = np.round(x.real * 65536.0)
    out[:, 1] = np.round(x.imag * 65536.0)
    return out


def add_impulses(iq: np.ndarray, seed: int, count: int, amplitude: float = 25000.0) -> np.ndarray:
    """Copy of `iq` with `count` one-sample impulses (ignition-noise like) on I and Q at seeded random positions --
    the input the LPC noise blanker (alt_noise_blanking, audio_nr.c:2210) is there for."""
    out = iq.astype(np.int64)
    rng = np.random.default_rng([seed, 0x4e42])
    pos = rng.choice(np.arange(64, len(iq) - 64), size=count, replace=False)
    sg = rng.choice([-1.0, 1.0], size=(count, 2))
    out[pos, 0] += np.round(sg[:, 0] * amplitude * 65536.0).astype(np.int64)
    out[pos, 1] += np.round(sg[:, 1] * 0.7 * amplitude * 65536.0).astype(np.int64)
    return out.clip(-2**31, 2**31 - 1).astype(np.int32)


def tx_mic(channel: int, nsamples: int, seed: int = 0x55485344, start: int = 0) -> np.ndarray:
    """int32 [nsamples, 2] microphone block stream: two-tone 700 + 1900 Hz, amplitude 8000 each, in `l`."""
    n = np.arange(start, start + nsamples, dtype=np.float64)
    t = n / FS
    rng = np.random.default_rng([seed, channel, start, 7])
    x = 8000.0 * np.sin(2 * np.pi * (700.0 + channel % 16) * t) + 8000.0 * np.sin(2 * np.pi * 1900.0 * t)
    x *= 0.5 + 0.5 * np.abs(np.sin(2 * np.pi * 1.5 * t))
    x += 20.0 * rng.standard_normal(nsamples)
    out = np.zeros((nsamples, 2), dtype=np.int32)
    out[:, 0] = np.round(x * 65536.0).astype(np.int64).clip(-2**31, 2**31 - 1)
    return out


# ------------------------------------------------------------------------------------------------
# Counter-based generator: every sample is a pure function of (seed, global channel,
# sample index), evaluated in 64-bit INTEGER arithmetic only -- table look-ups, multiplies, shifts
# and a 32-bit avalanche hash for the noise -- so the numpy (host) and the torch (device) evaluation
# of the same code are bit-identical and any channel / time slice can be regenerated anywhere.
# bench.py and the at-size parity tests use it; the golden vectors keep the float generator above.
# ------------------------------------------------------------------------------------------------
_SIN_BITS = 12
_SIN_TABLE = np.round(np.sin(2.0 * np.pi * np.arange((1 << _SIN_BITS) + 1) / (1 << _SIN_BITS)) * (1 << 30)).astype(np.int64)
# fading gain 10^(6 sin(phase) / 20) in Q15 over one period of the 0.5 Hz fade
_FADE_TABLE = np.round(10.0 ** (6.0 * np.sin(2.0 * np.pi * np.arange((1 << _SIN_BITS) + 1) / (1 << _SIN_BITS)) / 20.0) * (1 << 15)).astype(np.int64)
_M32 = 0xFFFFFFFF
KIND_SSB_USB, KIND_SSB_LSB, KIND_AM, KIND_FM, KIND_MIC = range(5)


def _fword(freq_hz: float) -> int:
    """32-bit phase increment per 48 ksps sample."""
    return int(round(freq_hz / FS * 4294967296.0)) & _M32


class _Ops:
    """The handful of array operations the generator needs, for numpy or torch (int64 everywhere)."""

    def __init__(self, xp, device=None):
        self.torch = xp.__name__ == "torch"
        self.xp, self.device = xp, device
        if self.torch:
            self.sin_t = xp.from_numpy(_SIN_TABLE).to(device)
            self.fade_t = xp.from_numpy(_FADE_TABLE).to(device)
        else:
            self.sin_t, self.fade_t = _SIN_TABLE, _FADE_TABLE

    def arange(self, a, b):
        return self.xp.arange(a, b, dtype=self.xp.int64, device=self.device) if self.torch else np.arange(a, b, dtype=np.int64)

    def asarray(self, v):
        return self.xp.as_tensor(v, dtype=self.xp.int64, device=self.device) if self.torch else np.asarray(v, dtype=np.int64)

    def lut(self, table, phase32):
        """table[phase] with linear interpolation; phase32 in [0, 2^32)."""
        idx = phase32 >> (32 - _SIN_BITS)
        frac = (phase32 >> (32 - _SIN_BITS - 16)) & 0xFFFF
        a, b = table[idx], table[idx + 1]
        return a + (((b - a) * frac) >> 16)

    def hash32(self, x):
        # lowbias32 avalanche (two multiply-xorshift rounds), all values kept below 2^32
        x = x & _M32
        x = x ^ (x >> 16); x = (x * 0x7FEB352D) & _M32
        x = x ^ (x >> 15); x = (x * 0x846CA68B) & _M32
        return x ^ (x >> 16)


def counter_block(xp, kinds, channels, start: int, nsamples: int, seed: int = 0x55485344, device=None, translate_hz: float = -12000.0):
    """int32 [len(channels), nsamples, 2] block of the synthetic signal of SURVEY.md 8d.

    kinds[i] in KIND_*: SSB (three in-band tones 700 / 1500 / 2100 Hz + 3 Hz x (channel mod 64) on the wanted sideband with a
    0.5 Hz +-6 dB fade, a 1500 Hz opposite-sideband interferer), AM (carrier 4000 at +37 Hz, 50 % 1 kHz modulation, fade), FM
    (1 kHz tone, 2.5 kHz deviation, amplitude 4000) -- all at the IF offset -translate_hz, plus noise of sigma 100 per I and Q;
    KIND_MIC: two-tone 700 + 1900 Hz microphone signal in .l (TX).  channels: global channel numbers; start: first sample index."""
    o = _Ops(xp, device)
    kinds = np.asarray(kinds, dtype=np.int64)
    chans = np.asarray(channels, dtype=np.int64)
    n = o.arange(start, start + nsamples)[None, :]                        # [1, ns]
    ch = o.asarray(chans)[:, None]                                         # [nch, 1]
    fc = -translate_hz

    def col(vals):
        return o.asarray(np.asarray(vals, dtype=np.int64))[:, None]

    def tone(fw_col, ph0_col=None):
        ph = (fw_col * n) & _M32 if ph0_col is None else (fw_col * n + ph0_col) & _M32
        return o.lut(o.sin_t, (ph + (1 << 30)) & _M32), o.lut(o.sin_t, ph)           # cos, sin (x 2^30)

    sgn = np.where(kinds == KIND_SSB_LSB, -1.0, 1.0)
    d = 3.0 * (chans % 64)
    re = n * 0 + ch * 0
    im = n * 0 + ch * 0
    ssb = (kinds == KIND_SSB_USB) | (kinds == KIND_SSB_LSB)
    am, fm, mic = kinds == KIND_AM, kinds == KIND_FM, kinds == KIND_MIC
    fade = o.lut(o.fade_t, (col([_fword(0.5)] * len(chans)) * n + col([int(0.1 * c / (2 * np.pi) * 4294967296.0) & _M32 for c in chans])) & _M32)   # Q15
    if ssb.any():
        m = col(ssb.astype(np.int64))
        for fa, amp in ((700.0, 3000), (1500.0, 2000), (2100.0, 1000)):
            c_, s_ = tone(col([_fword(fc + s * (fa + dd)) for s, dd in zip(sgn, d)]))
            re = re + m * ((((c_ * amp) >> 14) * fade) >> 15)
            im = im + m * ((((s_ * amp) >> 14) * fade) >> 15)
        c_, s_ = tone(col([_fword(fc - s * 1500.0) for s in sgn]))
        re = re + m * ((c_ * 3000) >> 14)
        im = im + m * ((s_ * 3000) >> 14)
    if am.any():
        m = col(am.astype(np.int64))
        _, ms = tone(col([_fword(1000.0 + dd) for dd in d]))
        env = (1 << 30) + (ms >> 1)                                        # 1 + 0.5 sin, x 2^30
        c_, s_ = tone(col([_fword(fc + 37.0)] * len(chans)))
        re = re + m * (((((c_ >> 8) * (env >> 8)) >> 14) * 4000 >> 14) * fade >> 15)
        im = im + m * (((((s_ >> 8) * (env >> 8)) >> 14) * 4000 >> 14) * fade >> 15)
    if fm.any():
        m = col(fm.astype(np.int64))
        _, ms = tone(col([_fword(1000.0)] * len(chans)))
        dev = (ms * int(round(2.5 / (2.0 * np.pi) * (1 << 20)))) >> 18     # 2.5 rad peak (2.5 kHz deviation / 1 kHz tone) in 2^-32 turns: 2.5/(2 pi) x 2^32 x ms x 2^-30
        ph = (col([_fword(fc)] * len(chans)) * n + dev) & _M32
        re = re + m * ((o.lut(o.sin_t, (ph + (1 << 30)) & _M32) * 4000) >> 14)
        im = im + m * ((o.lut(o.sin_t, ph) * 4000) >> 14)
    if mic.any():
        m = col(mic.astype(np.int64))
        _, s1 = tone(col([_fword(700.0 + (c % 16)) for c in chans]))
        _, s2 = tone(col([_fword(1900.0)] * len(chans)))
        _, sl = tone(col([_fword(1.5)] * len(chans)))
        slow = (1 << 29) + ((sl * ((sl >> 31) * 2 + 1)) >> 1)              # 0.5 + 0.5 |sin|, x 2^30
        re = re + m * (((((s1 + s2) >> 8) * (slow >> 8)) >> 14) * 8000 >> 14)
    # noise: Irwin-Hall sum of four 16-bit uniforms per component (sigma 2^16 / sqrt(3)), scaled to sigma 100 (20 for the microphone)
    key = (ch * 0x9E3779B1 + n * 0x85EBCA77 + (seed & _M32)) & _M32
    scale = col(np.where(mic, 20, 100))
    k_sig = int(round(65536.0 * np.sqrt(3.0)))                             # 100 x 2^16 / (2^16 / sqrt 3) = 100 sqrt 3, in Q16
    for comp in (0, 1):
        h1 = o.hash32(key ^ (0xA5A5A5A5 + 0x1F123BB5 * comp))
        h2 = o.hash32(h1 + 0x68E31DA4)
        u = (h1 & 0xFFFF) + (h1 >> 16) + (h2 & 0xFFFF) + (h2 >> 16) - 2 * 65535
        nz = (u * k_sig * scale) >> 16
        if comp == 0:
            re = re + nz
        else:
            im = im + nz * col((~mic).astype(np.int64))
    lo, hi = -(1 << 31), (1 << 31) - 1
    if o.torch:
        out = xp.stack([re.clamp(lo, hi), im.clamp(lo, hi)], dim=-1).to(xp.int32)
    else:
        out = np.stack([np.clip(re, lo, hi), np.clip(im, lo, hi)], axis=-1).astype(np.int32)
    return out


def kind_of(cfg: ChanCfg) -> int:
    m = cfg.dmod_mode
    if m == DEMOD_FM:
        return KIND_FM
    if m in (DEMOD_AM, DEMOD_SAM):
        return KIND_AM
    lsb = m == DEMOD_LSB or (m == 2 and cfg.cw_lsb) or (m == 6 and cfg.digi_lsb)
    return KIND_SSB_LSB if lsb else KIND_SSB_USB
