"""Host-side reader for the coefficient-table blob (include/uhsdr_tables.h).

Mirrors the reference's FilterPathInfo[] / FilterInfo[] lookup (audio_filter.c:47-80,147-922,
audio_filter.h:96-141).  Used by tests and bench to pick filter paths; the CUDA library parses
the same blob itself.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

MAGIC = 0x42545355
VERSION = 2
DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "uhsdr_tables.bin")

FILTER_MODE_CW, FILTER_MODE_SSB, FILTER_MODE_AM, FILTER_MODE_FM, FILTER_MODE_SAM = range(5)


@dataclass
class Path:
    index: int
    id: int
    mode_mask: int
    filter_select_id: int
    fir_numtaps: int
    fir_i_array: int
    fir_q_array: int
    fir_is_new_coeffs: int
    dec_array: int
    dec_numtaps: int
    sample_rate_dec: int
    pre_lattice: int
    interpolate: int
    aa_lattice: int
    offset_hz: int
    name: str


class Tables:
    def __init__(self, blob: bytes | None = None, path: str | None = None):
        if blob is None:
            with open(path or DEFAULT_BLOB, "rb") as f:
                blob = f.read()
        self.blob = blob
        hdr = struct.unpack_from("<16I", blob, 0)
        (magic, version, total, n_arr, arr_off, n_paths, paths_off, n_filt, filt_off,
         n_lat, lat_off, n_int, int_off, extras_off) = hdr[:14]
        if magic != MAGIC or version != VERSION or total != len(blob):
            raise ValueError("bad table blob")
        self.arrays = []
        for i in range(n_arr):
            off, cnt = struct.unpack_from("<2I", blob, arr_off + 8 * i)
            self.arrays.append(np.frombuffer(blob, dtype="<f4", count=cnt, offset=off))
        self.paths = []
        psz = 14 * 4 + 24
        for i in range(n_paths):
            v = struct.unpack_from("<14i24s", blob, paths_off + psz * i)
            self.paths.append(Path(i, *v[:14], v[14].split(b"\0")[0].decode("latin-1")))
        self.filters = []
        for i in range(n_filt):
            fid, width, name = struct.unpack_from("<2i12s", blob, filt_off + 20 * i)
            self.filters.append((fid, width, name.split(b"\0")[0].decode("latin-1")))
        self.lattices = [struct.unpack_from("<3i", blob, lat_off + 12 * i) for i in range(n_lat)]
        self.interps = [struct.unpack_from("<4i", blob, int_off + 16 * i) for i in range(n_int)]
        ex = struct.unpack_from("<17i", blob, extras_off)
        self.extras = dict(zip(
            ["nr_decimate_array", "nr_interpolate_array", "sqrt_hann_256_array", "spectrum_window_array",
             "sam_c0_array", "sam_c1_array", "fm_squelch_lattice", "tx_hilbert_i_array", "tx_hilbert_q_array",
             "tx_hilbert_numtaps", "tx_lattice_soprano", "tx_lattice_tenor", "tx_lattice_bass", "tx_lattice_fm",
             "dds_table_array", "zoom_biquad_array", "zoom_decim_array"], ex))

    def width(self, path_index: int) -> int:
        return self.filters[self.paths[path_index].id][1]

    def paths_for_mode(self, filter_mode: int):
        return [p.index for p in self.paths if p.mode_mask & (1 << filter_mode)]
