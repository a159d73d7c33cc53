"""GPU: merge / take / put / filter parity against the oracle (and pyarrow for filter, which the
reference does not have), at ragged sizes with all validity-bitmap combinations."""
import itertools

import numpy as np
import pytest

import arrow_gpu_b200 as ag
import oracle as O
from helpers import OArr, oracle_filter, oracle_merge, oracle_put, oracle_take
from test_gpu_parity import NAMES, SIZES, assert_same, make

pytestmark = pytest.mark.gpu

WIDTHS = [O.I8, O.U8, O.I16, O.U16, O.I32, O.U32, O.F32, O.DATE32]


def make_bool(rng, n, nulls, device, p=0.5):
    flags = rng.random(n) < p
    valid = rng.random(n) < 0.9 if nulls else None
    g = ag.BooleanArrayGPU.from_numpy(flags, valid, device)
    return g, OArr(O.BOOL, O.pack_bits(flags), n, O.pack_bits(valid) if nulls else None)


@pytest.mark.parametrize("dtype", WIDTHS, ids=lambda d: NAMES[d])
def test_merge(dtype, device):
    rng = np.random.default_rng(41 + dtype)
    for n in SIZES:
        # every combination of present/missing bitmaps (the reference only tests all-three: Q7)
        for na, nb, nm in itertools.product((False, True), repeat=3):
            a, oa = make(rng, dtype, n, na, device)
            b, ob = make(rng, dtype, n, nb, device)
            m, om = make_bool(rng, n, nm, device)
            assert_same(a.merge(b, m), oracle_merge(oa, ob, om), f"merge {NAMES[dtype]} n={n} {na}{nb}{nm}")


def test_merge_bool(device):
    rng = np.random.default_rng(42)
    for n in SIZES:
        for na, nb, nm in itertools.product((False, True), repeat=3):
            a, oa = make_bool(rng, n, na, device)
            b, ob = make_bool(rng, n, nb, device)
            m, om = make_bool(rng, n, nm, device)
            got, want = a.merge(b, m), oracle_merge(oa, ob, om)
            assert np.array_equal(device.retrive_data(got.data, O.words(n) * 4).view(np.uint32), want.data), n
            if want.valid is None:
                assert got.null_buffer is None
            else:
                assert np.array_equal(device.retrive_data(got.null_buffer.bit_buffer, O.words(n) * 4).view(np.uint32),
                                      want.valid), n


@pytest.mark.parametrize("dtype", WIDTHS, ids=lambda d: NAMES[d])
def test_take(dtype, device):
    rng = np.random.default_rng(45 + dtype)
    for src_n in (1, 33, 5000):
        for m in SIZES:
            for nulls in (False, True):
                a, oa = make(rng, dtype, src_n, nulls, device)
                idx = rng.integers(0, src_n, m).astype(np.uint32)
                if m > 10:
                    idx[3] = src_n + 7          # out of range reads as zero (robust buffer access)
                    idx[5] = 0xFFFFFFFF
                gi = ag.UInt32ArrayGPU.from_numpy(idx, None, device)
                assert_same(a.take(gi), oracle_take(oa, OArr(O.U32, idx, m)),
                            f"take {NAMES[dtype]} src={src_n} m={m} nulls={nulls}")


def test_take_bool(device):
    rng = np.random.default_rng(46)
    for src_n in (1, 100, 5000):
        for m in SIZES:
            a, oa = make_bool(rng, src_n, True, device)
            idx = rng.integers(0, src_n, m).astype(np.uint32)
            got = a.take(ag.UInt32ArrayGPU.from_numpy(idx, None, device))
            want = oracle_take(oa, OArr(O.U32, idx, m))
            assert got.len == m
            assert np.array_equal(device.retrive_data(got.data, O.words(m) * 4).view(np.uint32), want.data)
            assert np.array_equal(device.retrive_data(got.null_buffer.bit_buffer, O.words(m) * 4).view(np.uint32), want.valid)


@pytest.mark.parametrize("dtype", WIDTHS, ids=lambda d: NAMES[d])
def test_put(dtype, device):
    rng = np.random.default_rng(47 + dtype)
    for m in (0, 1, 300, 5000):     # > 256 indices: the reference's bool put breaks there (Q9)
        src_n, dst_n = 777, 9000
        src, osrc = make(rng, dtype, src_n, False, device)
        dst, odst = make(rng, dtype, dst_n, False, device)
        si = rng.integers(0, src_n, m).astype(np.uint32)
        di = rng.permutation(dst_n)[:m].astype(np.uint32)   # unique destinations: deterministic
        src.put(ag.UInt32ArrayGPU.from_numpy(si, None, device), dst, ag.UInt32ArrayGPU.from_numpy(di, None, device))
        assert_same(dst, oracle_put(osrc, OArr(O.U32, si, m), odst, OArr(O.U32, di, m)), f"put {NAMES[dtype]} m={m}")


def test_put_bool(device):
    rng = np.random.default_rng(48)
    for m in (0, 1, 300, 5000):
        src, osrc = make_bool(rng, 777, False, device)
        dst, odst = make_bool(rng, 9000, False, device)
        si = rng.integers(0, 777, m).astype(np.uint32)
        di = rng.permutation(9000)[:m].astype(np.uint32)
        src.put(ag.UInt32ArrayGPU.from_numpy(si, None, device), dst, ag.UInt32ArrayGPU.from_numpy(di, None, device))
        want = O.put(O.BOOL, osrc.data, si, odst.data, di)
        assert np.array_equal(device.retrive_data(dst.data, O.words(9000) * 4).view(np.uint32), want)


@pytest.mark.parametrize("dtype", WIDTHS, ids=lambda d: NAMES[d])
def test_filter(dtype, device):
    rng = np.random.default_rng(49 + dtype)
    for n in SIZES + [4096 * 3, 4096 * 3 + 1, 1_000_003]:
        for sel in (0.0, 0.1, 0.5, 0.9, 1.0):
            for nulls, mnulls in ((False, False), (True, True)):
                a, oa = make(rng, dtype, n, nulls, device)
                m, om = make_bool(rng, n, mnulls, device, p=sel)
                got, want = a.filter(m), oracle_filter(oa, om)
                assert_same(got, want, f"filter {NAMES[dtype]} n={n} sel={sel} nulls={nulls}")


def test_filter_matches_pyarrow(device):
    """cross-check of the new op's definition: pyarrow.compute.filter(null_selection_behavior='drop')"""
    pa = pytest.importorskip("pyarrow")
    import pyarrow.compute as pc
    rng = np.random.default_rng(50)
    n = 100_003
    vals = rng.integers(-2**31, 2**31, n).astype(np.int32)
    valid = rng.random(n) < 0.9
    flags = rng.random(n) < 0.4
    mvalid = rng.random(n) < 0.95
    a = ag.Int32ArrayGPU.from_numpy(vals, valid, device)
    m = ag.BooleanArrayGPU.from_numpy(flags, mvalid, device)
    got = a.filter(m).values()
    want = pc.filter(pa.array(vals, mask=~valid), pa.array(flags, mask=~mvalid), null_selection_behavior="drop")
    assert got == want.to_pylist()


def test_filter_large_properties(device):
    """BASELINE.json config-5 shaped check at a size the oracle cannot loop over quickly:
    500 M int32 rows (one GPU's shard of the 4 B-row column), verified by properties — count ==
    popcount(mask), output is the subsequence of kept rows (checked on samples and by checksum)."""
    n = 500_000_000
    rng = np.random.default_rng(40)
    vals = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(np.int32)
    flags = rng.random(n) < 0.5
    a = ag.Int32ArrayGPU.from_numpy(vals, None, device)
    m = ag.BooleanArrayGPU.from_numpy(flags, None, device)
    out = a.filter(m).raw_values()
    assert len(out) == int(flags.sum())
    assert np.array_equal(out, vals[flags])


def test_sharded_filter_nccl(device):
    """row-range shards + device-side count exchange (peer slots over CUDA IPC, and NCCL), global
    take over peer memory: one rank per VISIBLE GPU (a single GPU still runs every path with
    world_size 1, including the IPC mapping of its own slots)"""
    import os
    import subprocess
    import sys
    import arrow_gpu_b200._ffi as ffi
    world = max(1, ffi.device_count())
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(root, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "multi-GPU check ok" in res.stdout


def test_out_of_memory_releases_the_caches_of_every_handle(device):
    """each handle keeps freed blocks; when one handle runs out of device memory the blocks cached
    by the OTHER handles of the same GPU (an upload stream's handle, say) must be given back too"""
    import torch
    from arrow_gpu_b200 import _ffi
    lib = _ffi.lib()
    free_bytes, _total = torch.cuda.mem_get_info(device.ordinal)
    big = int(free_bytes * 0.6) // (1 << 20) * (1 << 20)
    other = ag.GpuDevice(device.ordinal)
    blk = other.create_empty_buffer(big)
    del blk                      # cached by `other`, not returned to the driver
    other.sync()
    got = device.create_empty_buffer(big)   # would not fit next to the cached block
    assert got.size >= big
    del got
    lib.agpu_trim(device.handle)
    lib.agpu_trim(other.handle)
    device.sync()
    other.sync()
