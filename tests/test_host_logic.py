"""CPU: host-side logic that needs no GPU — bitmap builders (the reference's own
BooleanBufferBuilder tests), the golden-vector extractor, row-range sharding and the 2-rank
count exchange for compaction over gloo."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from arrow_gpu_b200.array import BooleanBufferBuilder, pack_bits, unpack_bits
from arrow_gpu_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_set_bit():
    """crates/array/src/array/null_bit_buffer.rs:68-79"""
    buffer = BooleanBufferBuilder.new_with_capacity(10)
    assert len(buffer.data) == 2
    buffer.set_bit(0)
    assert buffer.data[0] == 0b00000001
    buffer.set_bit(9)
    assert buffer.data[1] == 0b00000010
    assert not buffer.is_set(5)
    assert buffer.is_set(9)
    assert buffer.is_set(0)


def test_new_set_with_capacity():
    """crates/array/src/array/null_bit_buffer.rs:81-87"""
    buffer = BooleanBufferBuilder.new_set_with_capacity(10)
    assert buffer.data[0] == 0xFF
    assert buffer.data[1] == 0b00000011


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 31, 32, 33, 1000):
        flags = rng.random(n) < 0.5
        packed = pack_bits(flags)
        assert len(packed) % 4 == 0 and len(packed) == (n + 31) // 32 * 4
        assert np.array_equal(unpack_bits(packed, n), flags)
        if n % 32:
            assert packed.view(np.uint32)[-1] >> (n % 32) == 0  # padding bits are zero


def test_extractor_is_reproducible():
    """The committed fixture equals what the extractor produces from the reference (when the
    reference tree is present, i.e. in the build container; it does not travel to the GPU box)."""
    if not os.path.isdir("/root/reference/crates"):
        pytest.skip("reference tree not present on this box")
    golden = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
    before = open(golden).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "extract_reference_vectors.py")],
                   check=True, capture_output=True)
    assert open(golden).read() == before
    assert len(json.loads(before)["cases"]) == 202


def test_row_range_partition():
    for n, world in ((0, 1), (1, 8), (1023, 4), (1024, 4), (10_000_000, 8), (4_000_000_000, 8), (4096 * 7 + 5, 3)):
        parts = [sharded.row_range(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
            assert e0 == b1 and b0 <= e0
        for b, e in parts[:-1]:
            # bitmaps split on 128-byte lines (ranges clipped at n are the empty/last shards)
            assert (b % sharded.SHARD_ALIGN == 0 or b == n) and (e % sharded.SHARD_ALIGN == 0 or e == n)
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= sharded.SHARD_ALIGN * 2 or n < sharded.SHARD_ALIGN * world


def test_exclusive_offsets():
    assert sharded.exclusive_offsets([3, 0, 5, 2]) == ([0, 3, 3, 8], 10)
    assert sharded.exclusive_offsets([]) == ([], 0)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = [7, 11][rank]
        offsets, total = sharded.exchange_counts(local)
        mx = sharded.max_over_ranks(float(rank + 1))
        words = sharded.gather_words([0x3F800000, 0xFFFFFFFF][rank])
        q.put((rank, offsets, total, mx, words))
    finally:
        dist.destroy_process_group()


def test_count_exchange_two_ranks_gloo():
    """the only collective on the path: all-gather of per-shard compaction counts (world 2, gloo)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0] == (0, [0, 7], 18, 2.0, [0x3F800000, 0xFFFFFFFF])
    assert res[1] == (1, [0, 7], 18, 2.0, [0x3F800000, 0xFFFFFFFF])


def test_combine_partial_sums_is_a_rank_ordered_fold():
    """the cross-shard sum is defined, not left to the collective: f32 adds in rank order, ints wrap"""
    f = sharded.combine_partial_sums(np.array([1e8, 1.0, -1e8, 1.0], dtype=np.float32), np.float32)
    assert f == np.float32(np.float32(np.float32(np.float32(1e8) + np.float32(1.0)) - np.float32(1e8)) + np.float32(1.0))
    assert f == np.float32(1.0)               # (1e8 + 1) rounds to 1e8 in f32: order matters and is fixed
    assert sharded.combine_partial_sums(np.array([2**31 - 1, 1], dtype=np.int32), np.int32) == np.int32(-2**31)
    assert sharded.combine_partial_sums(np.array([2**32 - 1, 2], dtype=np.uint32), np.uint32) == np.uint32(1)
    assert sharded.combine_partial_sums(np.array([], dtype=np.float32), np.float32) == np.float32(0)


def test_exchange_result_block_is_parsed_like_the_wait_kernel_writes_it():
    """agpu_exchange_wait writes {offset of every rank, total, status, count of every rank}
    (include/agpu.h); a non-zero status (a peer never posted) must raise, not return offsets"""
    from arrow_gpu_b200 import sharded
    from arrow_gpu_b200._ffi import AgpuError
    counts = [7, 0, 5, 11]
    offs, total = sharded.exclusive_offsets(counts)
    words = offs + [total, 0] + counts
    assert sharded.parse_exchange_result(words, 2, 4) == (offs, total, counts)
    assert offs == [0, 7, 7, 12] and total == 23
    words[5] = 1
    with pytest.raises(AgpuError):
        sharded.parse_exchange_result(words, 2, 4)
