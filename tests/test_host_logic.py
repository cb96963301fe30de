"""Host-side logic that needs no GPU: table blob parsing, filter-path lookup, synthetic signals,
channel partitioning across ranks (world_size 2 over gloo)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from uhsdr_b200 import synth
from uhsdr_b200.config import DEMOD_LSB, default_cfg
from uhsdr_b200.tables import FILTER_MODE_AM, FILTER_MODE_CW, FILTER_MODE_FM, FILTER_MODE_SSB, Tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_blob_matches_reference_layout():
    t = Tables()
    assert len(t.paths) == 87 and len(t.filters) == 31          # AUDIO_FILTER_PATH_NUM, AUDIO_FILTER_NUM
    p35 = t.paths[35]
    assert (p35.fir_numtaps, p35.dec_numtaps, p35.sample_rate_dec, p35.fir_is_new_coeffs) == (199, 83, 4, 1)
    assert t.width(35) == 2300 and t.width(38) == 2700
    assert t.lattices[p35.pre_lattice][0] == 10
    assert t.paths[48].fir_numtaps == 89 and t.paths[48].dec_numtaps == 43 and not t.paths[48].fir_is_new_coeffs
    assert t.paths[55].sample_rate_dec == 2 and t.paths[55].pre_lattice == -1
    assert t.paths[2].sample_rate_dec == 1 and t.paths[2].interpolate == -1
    assert set(t.paths_for_mode(FILTER_MODE_FM)) == {1, 2, 3}
    assert min(t.paths_for_mode(FILTER_MODE_AM)) == 66 and 35 in t.paths_for_mode(FILTER_MODE_CW)
    assert 36 in t.paths_for_mode(FILTER_MODE_SSB) and 36 not in t.paths_for_mode(FILTER_MODE_CW)
    assert t.arrays[t.extras["tx_hilbert_i_array"]].size == 201
    assert t.arrays[t.extras["spectrum_window_array"]].size == 1024
    with pytest.raises(ValueError):
        Tables(blob=t.blob[:-4])


def test_hilbert_pair_sideband_convention():
    """SURVEY.md 8d sanity numbers: with the CMSIS tap order, +f passes USB = I+Q with gain 2 and -f
    is rejected by >= 70 dB (199-tap pair at 12 ksps)."""
    t = Tables()
    p = t.paths[35]
    hi, hq = t.arrays[p.fir_i_array][::-1].astype(np.float64), t.arrays[p.fir_q_array][::-1].astype(np.float64)
    n = np.arange(199)
    for f, lo, hi_lim in ((1500.0, 1.95, 2.05), (-1500.0, 0.0, 2e-4)):
        w = np.exp(-2j * np.pi * f / 12000.0 * n)
        g = abs(np.sum(hi * w) * 1.0 + np.sum(hq * w) * (-1j if True else 1j) * 1.0)   # I=cos, Q=sin
        g = abs(np.sum(hi * w) + (-1j) * np.sum(hq * w))
        assert lo <= g <= hi_lim, (f, g)


def test_synth_is_deterministic_and_sliceable():
    cfg = default_cfg()
    a = synth.rx_iq(cfg, 7, 4096)
    b = synth.rx_iq(cfg, 7, 4096)
    assert np.array_equal(a, b) and a.dtype == np.int32 and a.shape == (4096, 2)
    assert not np.array_equal(a, synth.rx_iq(cfg, 8, 4096))
    assert np.max(np.abs(a)) < 2**31 - 1
    lsb = synth.rx_iq(default_cfg(dmod_mode=DEMOD_LSB), 7, 4096)
    assert not np.array_equal(a, lsb)


def test_channel_range_and_kind_partition():
    from uhsdr_b200.partition import channel_range, partition_by_kind
    for total, world in ((4096, 1), (65536, 8), (10, 4), (3, 8), (0, 2)):
        spans = [channel_range(r, world, total) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        channel_range(2, 2, 8)
    kinds = ["am"] * 5 + ["sam"] * 6 + ["fm"] * 7            # BASELINE configs[2]-style mixed plan
    owned = partition_by_kind(kinds, 2)
    assert sorted(owned[0] + owned[1]) == list(range(18))
    for r in range(2):
        mix = [kinds[c] for c in owned[r]]
        assert mix == sorted(mix, key=["am", "sam", "fm"].index)           # grouped by kind
        assert abs(mix.count("am") - 2.5) <= 0.5 and mix.count("sam") == 3 and abs(mix.count("fm") - 3.5) <= 0.5


def test_rank_partition_gloo_world2(tmp_path):
    """bench.py's sharding rule (uhsdr_b200.partition): rank r owns a contiguous global channel range,
    no data-path collective; the ranks only agree on the max-over-ranks time through one all_reduce.
    Exercised with 2 CPU processes over gloo."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from uhsdr_b200.partition import channel_range, partition_by_kind\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "total = 13\n"
        "lo, hi = channel_range(r, w, total)\n"
        "mine = torch.full((8,), -1, dtype=torch.int64)\n"
        "mine[: hi - lo] = torch.arange(lo, hi)\n"
        "allc = [torch.empty_like(mine) for _ in range(w)]\n"
        "dist.all_gather(allc, mine)\n"
        "got = torch.cat([a[a >= 0] for a in allc])\n"
        "assert torch.equal(got, torch.arange(total)), got\n"
        "kinds = ['usb', 'lsb'] * 5 + ['fm'] * 4\n"
        "own = partition_by_kind(kinds, w)[r]\n"
        "cnt = torch.tensor([sum(kinds[c] == k for c in own) for k in ('usb', 'lsb', 'fm')])\n"
        "tot = cnt.clone(); dist.all_reduce(tot)\n"
        "assert tot.tolist() == [5, 5, 4], tot\n"
        "assert all(abs(2 * c - t) <= 1 for c, t in zip(cnt.tolist(), tot.tolist()))\n"
        "t = torch.tensor([10.0 + r], dtype=torch.float64)\n"
        "dist.all_reduce(t, op=dist.ReduceOp.MAX)\n"
        "assert t.item() == 10.0 + w - 1\n"
        "dist.barrier()\n"
        "sys.stdout.write(f'rank{r}ok\\n'); sys.stdout.flush()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sk:          # a free port: a fixed one collides with concurrent runs
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rank0ok" in out.stdout and "rank1ok" in out.stdout, out.stdout


def test_bench_reference_arm_exits_cleanly_on_nonzero_rank():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_plans_shard_like_the_reference_configs():
    """bench.py's per-rank channel plans (the N > 1 legs of the other configurations): the alternating USB / LSB plan continues
    across rank boundaries as one global sequence (configs[1], configs[4] at 65536 channels over 8 ranks), mixed plans give every
    rank the same share of every kind with the kinds contiguous, and every local channel has a generator kind."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from uhsdr_b200.config import DEMOD_LSB
    world, total = 8, 65536
    seq = []
    for r in range(world):
        per = total // world
        cfgs, labels, kinds = bench.local_plan("rx_tx", per, r * per)
        assert len(cfgs) == len(labels) == len(kinds) == per
        seq += [c.dmod_mode == DEMOD_LSB for c in cfgs]
    assert seq == [bool(i % 2) for i in range(total)]
    # an odd first channel starts with the LSB kind
    cfgs, _, _ = bench.local_plan("ssb_narrow", 5, 3)
    assert [c.dmod_mode == DEMOD_LSB for c in cfgs] == [True, False, True, False, True]
    for name in ("ssb_wide", "mixed_am_sam_fm", "ssb_nr_spectrum"):
        groups = bench.plan_groups(name)
        tot = sum(g[2] for g in groups)
        cfgs, labels, kinds = bench.local_plan(name, 4096, 0)
        assert len(cfgs) == 4096 and len(set(kinds)) >= 1
        # kinds contiguous, shares as declared
        runs = [labels[0]]
        for lab in labels[1:]:
            if lab != runs[-1]:
                runs.append(lab)
        assert runs == [g[0] for g in groups]
        for lab, _, share in groups[:-1]:
            assert labels.count(lab) == 4096 * share // tot
