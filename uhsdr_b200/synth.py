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
