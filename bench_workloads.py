"""The other BASELINE.json configurations (cfg1, cfg3, cfg4, cfg5): `per_config()` runs them inside
the default `bench.py` line (each with its own per-op table, clocks window, cpu_baseline, parity
check on a bounded sample and, for cfg 1 / 3 / 5-filter, an e2e number); `run()` runs ONE of them
alone (`bench.py --workload cfgN`, used for ncu captures and probes).

Columns of 1-4 G rows are synthesised ON THE DEVICE with torch (seeded generators) — torch is used
only as an allocator/RNG here; every measured kernel is ours, launched through the C ABI on the
library's own stream and timed with CUDA events on that stream.  oracle/ is imported only by the
cpu_baseline / parity legs (`_oracle()`), never by anything that is timed as ours.
"""
from __future__ import annotations

import ctypes as C
import statistics
import time

import numpy as np

ROWS = {"allops": 268_435_456, "cfg1": 1_048_576, "cfg3": 1_000_000_000, "cfg4": 4_000_000_000, "cfg5": 4_000_000_000,
        "sweep": 1 << 30}
SAMPLE_ROWS = 1 << 24      # CPU / parity sample: the reference's own per-op maximum (gpu_device.rs:69,133)
E2E_ROWS = 1 << 28         # rows of the e2e legs of cfg 3 and cfg 5 (host buffers of 1 GiB per f32 column)


def _oracle():
    import oracle as O
    return O


def _wrap(ag, cls, tensor, n, dev, keep):
    """zero-copy: a torch CUDA tensor's memory as one of our arrays (not owned by our pool)"""
    keep.append(tensor)
    return cls(ag.ArrowGpuBuffer(dev, tensor.data_ptr(), tensor.numel() * tensor.element_size(), owned=False), dev, n, None)


def _bitmap(torch, n, p, seed, device):
    """Bernoulli(p) bits, LSB-first, as a uint8 tensor padded to whole u32 words"""
    nbytes = (n + 31) // 32 * 4
    out = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    weights = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=device)
    chunk = 1 << 28
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        bits = (torch.rand(m, generator=g, device=device) < p)
        pad = (-m) % 8
        if pad:
            bits = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=device)])
        packed = (bits.view(-1, 8).to(torch.int32) * weights).sum(dim=1).to(torch.uint8)
        out[start // 8: start // 8 + packed.numel()] = packed
        del bits, packed
    return out


def _uniform(torch, n, lo, hi, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.float32, device=device)
    chunk = 1 << 28
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        out[start:start + m].uniform_(lo, hi, generator=g)
    return out


def _randint32(torch, n, seed, device, lo=-2**31, hi=2**31):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.int32, device=device)
    chunk = 1 << 28
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        out[start:start + m] = torch.randint(lo, hi, (m,), generator=g, device=device, dtype=torch.int64).to(torch.int32)
    return out


def _arange_u32(torch, n, device):
    """0, 1, ..., n-1 as u32 bit patterns (n may exceed 2^31) without a 8-byte-per-row temporary"""
    out = torch.empty(n, dtype=torch.int32, device=device)
    chunk = 1 << 28
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        out[start:start + m] = torch.arange(start, start + m, dtype=torch.int64, device=device).to(torch.int32)
    return out


class Timer:
    def __init__(self, dev, lib, ffi):
        self.dev, self.lib, self.ffi = dev, lib, ffi

    def event(self):
        e = C.c_void_p()
        self.ffi.check(self.lib.agpu_event_create(C.byref(e)), "event_create")
        return e

    def record(self, e):
        self.ffi.check(self.lib.agpu_event_record(self.dev.handle, e), "event_record")

    def ms(self, a, b):
        out = C.c_float(0)
        self.ffi.check(self.lib.agpu_event_elapsed_ms(a, b, C.byref(out)), "elapsed")
        return out.value


class Ctx:
    """everything a config builder needs"""

    def __init__(self, rank, world, local_rank, dev=None):
        import torch

        import arrow_gpu_b200 as ag
        from arrow_gpu_b200 import _ffi, kernels as K, sharded
        self.torch, self.ag, self.K, self.sharded, self.ffi = torch, ag, K, sharded, _ffi
        torch.cuda.set_device(local_rank)
        self.tdev = torch.device("cuda", local_rank)
        self.dev = dev or ag.GpuDevice(local_rank)
        self.lib = _ffi.lib()
        self.T = Timer(self.dev, self.lib, _ffi)
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.seed = 1000 * rank
        self.keep = []

    def wrap(self, cls, tensor, n=None):
        return _wrap(self.ag, cls, tensor, tensor.numel() if n is None else n, self.dev, self.keep)

    def bool_array(self, bits_u8, n):
        self.keep.append(bits_u8)
        ag = self.ag
        return ag.BooleanArrayGPU(ag.ArrowGpuBuffer(self.dev, bits_u8.data_ptr(), bits_u8.numel(), owned=False), self.dev, n, None)

    def nullable(self, arr, bits_u8):
        self.keep.append(bits_u8)
        ag = self.ag
        arr.null_buffer = ag.NullBitBufferGpu(ag.ArrowGpuBuffer(self.dev, bits_u8.data_ptr(), bits_u8.numel(), owned=False),
                                              arr.len, self.dev)
        return arr

    def release(self):
        """drop this config's columns and give cached blocks back before the next config"""
        import gc
        self.keep.clear()
        gc.collect()
        self.dev.sync()
        self.lib.agpu_trim(self.dev.handle)
        self.dev.sync()
        self.torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------
# config builders: list of (label, algorithmic bytes per counted row, rows counted, fn)
# ---------------------------------------------------------------------------------------------
def build_cfg3(ctx, n):
    torch, ag, K, dev = ctx.torch, ctx.ag, ctx.K, ctx.dev
    cols = []
    for k in range(4):
        t = _uniform(torch, n, -10, 10, 20 + k + ctx.seed, ctx.tdev)
        cols.append(ctx.nullable(ctx.wrap(ag.Float32ArrayGPU, t), _bitmap(torch, n, 0.95, 24 + k + ctx.seed, ctx.tdev)))

    def chain(fuse):
        p = ag.ArrowComputePipeline(dev, "chain", fuse=fuse)
        r = K.gt_op_dyn(K.add_op_dyn(K.mul_op_dyn(cols[0], cols[1], p), cols[2], p), cols[3], p)
        p.finish()
        return r

    def chain2():
        p = ag.ArrowComputePipeline(dev, "chain2")
        r = K.add_op_dyn(K.mul_op_dyn(K.sin_op_dyn(cols[0], p), cols[1], p), cols[2], p)
        p.finish()
        return r
    ops = [("fused (a*b+c)>d + 4 bitmaps", 16.75, n, lambda: K.fused_mul_add_gt(*cols)),
           # (mul, add, gt) itself is routed to the dedicated kernel; `lteq` keeps this line on the interpreter
           ("generic chain interpreter [mul b, add c, lteq d] + 4 bitmaps", 16.75, n,
            lambda: K.fused_chain(cols[0], [("mul", cols[1]), ("add", cols[2]), ("lteq", cols[3])])),
           ("unfused chain mul,add,gt (reference style, 3 kernels)", 33.25, n, lambda: chain(False)),
           ("same recorded chain on ArrowComputePipeline(fuse=True) (auto-fused -> the dedicated kernel)", 16.75, n,
            lambda: chain(True)),
           ("generic fused_chain [sin, mul b, add c] -> f32 + 3 bitmaps", 16.5, n,
            lambda: K.fused_chain(cols[0], [("sin",), ("mul", cols[1]), ("add", cols[2])])),
           ("unfused sin,mul,add (3 kernels)", 32.875, n, chain2)]
    return ops, {"cols": cols}


CFG4_RANGES = {"sqrt": (0, 1e6), "exp": (-20, 20), "sin": (-100, 100), "cos": (-100, 100)}


def build_cfg4(ctx, n):
    torch, ag = ctx.torch, ctx.ag
    col, ops, by_op = {}, [], {}
    for k, (op, (lo, hi)) in enumerate(CFG4_RANGES.items()):
        if (lo, hi) not in col:
            col[(lo, hi)] = ctx.wrap(ag.Float32ArrayGPU, _uniform(torch, n, lo, hi, 30 + k + ctx.seed, ctx.tdev))
        arr = col[(lo, hi)]
        by_op[op] = arr
        ops.append((f"f32.{op}", 8.0, n, (lambda arr=arr, op=op: getattr(arr, op)())))
    return ops, {"by_op": by_op}


CFG5_SELECTIVITIES = ((0.1, 42), (0.5, 43), (0.9, 44))


def build_cfg5(ctx, n, total_rows):
    torch, ag, sharded = ctx.torch, ctx.ag, ctx.sharded
    a = ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 40 + ctx.seed, ctx.tdev))
    b_ = ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 140 + ctx.seed, ctx.tdev))
    torch.cuda.empty_cache()
    m50 = ctx.bool_array(_bitmap(torch, n, 0.5, 41 + ctx.seed, ctx.tdev), n)
    ops = [("i32.merge", 12.125, n, lambda: a.merge(b_, m50))]
    masks = {}
    for s, sd in CFG5_SELECTIVITIES:
        mk = ctx.bool_array(_bitmap(torch, n, s, sd + ctx.seed, ctx.tdev), n)
        masks[s] = mk

        def filt(mk=mk):   # count -> [post] -> scatter -> [wait]; ONE host sync at the very end (the result block)
            out, _off, _tot = sharded.sharded_filter(a, mk)
            return out

        def filt_async(mk=mk):   # the same kernels, the counts stay on the device (a dependent GPU op could follow)
            return sharded.sharded_filter_async(a, mk)
        ops.append((f"i32.filter s={s}", 4.125 + 4 * s, n, filt))
        ops.append((f"i32.filter s={s} (enqueue only: counts left on the device)", 4.125 + 4 * s, n, filt_async))
    torch.cuda.empty_cache()
    idx_seq = ctx.wrap(ag.UInt32ArrayGPU, _arange_u32(torch, n, ctx.tdev))
    idx_rnd = ctx.wrap(ag.UInt32ArrayGPU, _randint32(torch, n, 45 + ctx.seed, ctx.tdev, 0, n))
    torch.cuda.empty_cache()
    ops.append(("i32.take sorted stride-1", 12.0, n, lambda: a.take(idx_seq)))
    ops.append(("i32.take uniform random", 12.0, n, lambda: a.take(idx_rnd)))
    extras = {"a": a, "b": b_, "m50": m50, "masks": masks}
    if ctx.world > 1:
        # global row numbers over all shards: the gather kernel reads peer shards over NVLink
        col = sharded.ShardedColumn(ag.Int32ArrayGPU, a, None, total_rows, ctx.dev)
        ctx.keep.append(col)
        m_g = min(n, 1 << 28)
        gidx = ctx.wrap(ag.UInt32ArrayGPU, _randint32(torch, m_g, 46 + ctx.seed, ctx.tdev, 0, min(total_rows, 2**31 - 1)))
        ops.append((f"i32.take GLOBAL uniform random over {ctx.world} shards (NVLink peer loads, {m_g} rows/GPU)", 12.0, m_g,
                    lambda: col.take_global(gidx)))
        extras["sharded_column"] = col
    return ops, extras


def build_cfg1(ctx, n, numa=None):
    """SURVEY.md §8(d) cfg 1: a, b ~ U(-1000, 1000) from default_rng(1), validity Bernoulli(0.9) seeds 2, 3,
    null slots zeroed — on the HOST (this is the reference's CPU-runnable config), then uploaded"""
    ag, dev = ctx.ag, ctx.dev
    rng = np.random.default_rng(1 + ctx.seed)
    a_h = rng.uniform(-1000, 1000, n).astype(np.float32)
    b_h = rng.uniform(-1000, 1000, n).astype(np.float32)
    va = np.random.default_rng(2 + ctx.seed).random(n) < 0.9
    vb = np.random.default_rng(3 + ctx.seed).random(n) < 0.9
    a_h[~va] = 0
    b_h[~vb] = 0
    a = ag.Float32ArrayGPU.from_numpy(a_h, va, dev)
    b = ag.Float32ArrayGPU.from_numpy(b_h, vb, dev)
    ops = [("f32.add+validity", 12.375, n, lambda: a.add(b)), ("f32.gt+validity", 8.5, n, lambda: a.gt(b))]
    return ops, {"a": a, "b": b, "a_h": a_h, "b_h": b_h, "va": va, "vb": vb}


def build_allops(ctx, n):
    torch, ag, tdev, dev = ctx.torch, ctx.ag, ctx.tdev, ctx.dev
    seed = ctx.seed
    # every remaining (op, dtype) family of the path at 256 Mi rows, to find kernels that fall
    # short of the roofline (config 2 already covers sub-word arithmetic/logical/shift/cast)
    f = [ctx.wrap(ag.Float32ArrayGPU, _uniform(torch, n, 0.5, 50.0, 60 + k + seed, tdev)) for k in range(2)]
    f[0] = ctx.nullable(f[0], _bitmap(torch, n, 0.9, 70 + seed, tdev))
    unit = ctx.wrap(ag.Float32ArrayGPU, _uniform(torch, n, -1.0, 1.0, 63 + seed, tdev))
    i32 = [ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 64 + k + seed, tdev)) for k in range(2)]
    u32 = [ctx.wrap(ag.UInt32ArrayGPU, _randint32(torch, n, 66 + k + seed, tdev)) for k in range(2)]
    small = ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 68 + seed, tdev, -6, 12))
    cnt = ctx.wrap(ag.UInt32ArrayGPU, _randint32(torch, n, 69 + seed, tdev, 0, 32))

    def sub(t, cls):
        g = torch.Generator(device=tdev)
        g.manual_seed(80 + seed)
        info = torch.iinfo(t)
        x = torch.randint(info.min, info.max + 1, (n,), generator=g, device=tdev, dtype=torch.int32).to(t)
        return ctx.wrap(cls, x)
    i8a, i8b = sub(torch.int8, ag.Int8ArrayGPU), sub(torch.int8, ag.Int8ArrayGPU)
    u16a, u16b = sub(torch.int16, ag.UInt16ArrayGPU), sub(torch.int16, ag.UInt16ArrayGPU)
    mbits = _bitmap(torch, n, 0.5, 90 + seed, tdev)
    m = ctx.bool_array(mbits, n)
    m2 = ctx.bool_array(mbits, n)
    sc_f = ag.Float32ArrayGPU.from_slice([1.5], dev)
    sc_i = ag.Int32ArrayGPU.from_slice([7], dev)
    ops = [
        ("f32.add (+validity)", 12.25, n, lambda: f[0].add(f[1])), ("f32.div", 12.25, n, lambda: f[0].div(f[1])),
        ("f32.min", 12.25, n, lambda: f[0].min(f[1])), ("f32.power", 12.25, n, lambda: f[0].power(f[1])),
        ("f32.rem_scalar", 8.25, n, lambda: f[0].rem_scalar(sc_f)), ("f32.neg", 8.25, n, lambda: f[0].neg()),
        ("f32.abs", 8.25, n, lambda: f[0].abs()), ("f32.cbrt", 8.25, n, lambda: f[0].cbrt()),
        ("f32.exp2", 8.25, n, lambda: f[0].exp2()), ("f32.log", 8.25, n, lambda: f[0].log()),
        ("f32.log2", 8.25, n, lambda: f[0].log2()), ("f32.acos", 8, n, lambda: unit.acos()),
        ("f32.sinh", 8, n, lambda: unit.sinh()), ("f32.gt -> bitmap", 8.375, n, lambda: f[0].gt(f[1])),
        ("f32.eq -> bitmap", 8.375, n, lambda: f[0].eq(f[1])), ("f32.sum", 4, n, lambda: f[1].sum()),
        ("f32.cast u8", 5.25, n, lambda: f[0].cast(ag.UInt8ArrayGPU)),
        ("i32.add", 12, n, lambda: i32[0].add(i32[1])), ("i32.div_scalar", 8, n, lambda: i32[0].div_scalar(sc_i)),
        ("i32.rem_scalar", 8, n, lambda: i32[0].rem_scalar(sc_i)), ("i32.max", 12, n, lambda: i32[0].max(i32[1])),
        ("i32.abs", 8, n, lambda: i32[0].abs()), ("i32.power (|p| small)", 12, n, lambda: i32[0].power(small)),
        ("i32.lt -> bitmap", 8.125, n, lambda: i32[0].lt(i32[1])), ("i32.shl", 12, n, lambda: i32[0].bitwise_shl(cnt)),
        ("i32.sum", 4, n, lambda: i32[0].sum()), ("u32.xor", 12, n, lambda: u32[0].bitwise_xor(u32[1])),
        ("u32.bitcast f32", 8, n, lambda: u32[0].bitcast(ag.Float32ArrayGPU)),
        ("i8.gt -> bitmap", 2.125, n, lambda: i8a.gt(i8b)), ("i8.min", 3, n, lambda: i8a.min(i8b)),
        ("i8.sin -> f32 (fused cast)", 5, n, lambda: i8a.sin()), ("i8.merge", 3.125, n, lambda: i8a.merge(i8b, m)),
        ("u16.lteq -> bitmap", 4.125, n, lambda: u16a.lteq(u16b)), ("u16.max", 6, n, lambda: u16a.max(u16b)),
        ("u16.cos -> f32 (fused cast)", 6, n, lambda: u16a.cos()), ("u16.merge", 6.125, n, lambda: u16a.merge(u16b, m)),
        ("f32.merge (+validity of a)", 12.375, n, lambda: f[0].merge(f[1], m)),
        ("bool.and", 0.375, n, lambda: m.bitwise_and(m2)), ("bool.not", 0.25, n, lambda: m.bitwise_not()),
        ("bool.all", 0.125, n, lambda: m.all()), ("bool.cast f32", 4.125, n, lambda: m.cast(ag.Float32ArrayGPU)),
        ("bool.merge", 0.5, n, lambda: m.merge(m2, m)),
        ("i8.filter s=0.5", 1.125 + 0.5, n, lambda: i8a.filter(m)), ("u16.filter s=0.5", 2.125 + 1, n, lambda: u16a.filter(m)),
        ("f32.filter s=0.5 (+validity)", 4.25 + 2.0625, n, lambda: f[0].filter(m)),
    ]
    # gathers / scatters outside config 5: sequential indices (the streaming bound of the kernel)
    seq = torch.arange(n, dtype=torch.int32, device=tdev)
    idx = ctx.wrap(ag.UInt32ArrayGPU, seq)
    dst = ag.Int32ArrayGPU.empty(n, dev)
    ops += [
        ("f32.take sequential (+validity gather)", 12.25, n, lambda: f[0].take(idx)),
        ("i8.take sequential", 6, n, lambda: i8a.take(idx)),
        ("bool.take sequential", 4.25, n, lambda: m.take(idx)),
        ("i32.put sequential (one index column used for both sides)", 12, n, lambda: i32[0].put(idx, dst, idx)),
        ("f32.broadcast", 4, n, lambda: ag.Float32ArrayGPU.broadcast(1.5, n, dev)),
    ]
    # fused integer chains (agpu_fused_chain_int) and the same ops one kernel each
    K = ctx.K
    sc8 = ag.Int8ArrayGPU.from_slice([3], dev)
    sc16 = ag.UInt16ArrayGPU.from_slice([3], dev)
    i8c = sub(torch.int8, ag.Int8ArrayGPU)
    i32c = ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 95 + seed, tdev))
    i32d = ctx.wrap(ag.Int32ArrayGPU, _randint32(torch, n, 96 + seed, tdev))
    ops += [
        ("i8 chain [add b, and c, mul s] fused", 4, n,
         lambda: K.fused_chain_int(i8a, [("add", i8b), ("bitwise_and", i8c), ("mul", K.DeviceScalar(sc8))])),
        ("i8 chain [add b, and c, mul s] unfused (3 kernels)", 8, n, lambda: i8a.add(i8b).bitwise_and(i8c).mul_scalar(sc8)),
        ("u16 chain [not, add s, xor b] fused", 6, n,
         lambda: K.fused_chain_int(u16a, [("bitwise_not",), ("add", K.DeviceScalar(sc16)), ("bitwise_xor", u16b)])),
        ("u16 chain [not, add s, xor b] unfused (3 kernels)", 14, n, lambda: u16a.bitwise_not().add_scalar(sc16).bitwise_xor(u16b)),
        ("i32 chain [mul b, add c, gt d] fused", 16.125, n,
         lambda: K.fused_chain_int(i32[0], [("mul", i32[1]), ("add", i32c), ("gt", i32d)])),
        ("i32 chain [mul b, add c, gt d] unfused (3 kernels)", 32.125, n, lambda: i32[0].mul(i32[1]).add(i32c).gt(i32d)),
    ]
    return ops, {}


def build_sweep(ctx, n):
    # column-size sweep of one binary op with validity (f32 add, 12.375 B/row) from the
    # reference's test sizes up to 1 Gi rows: where launch latency ends and HBM begins.
    # 20 calls back to back per measurement (no host sync in between).
    torch, ag, tdev, dev, seed = ctx.torch, ctx.ag, ctx.tdev, ctx.dev, ctx.seed
    base_a = _uniform(torch, n, -1000, 1000, 1 + seed, tdev)
    base_b = _uniform(torch, n, -1000, 1000, 2 + seed, tdev)
    va = _bitmap(torch, n, 0.9, 3 + seed, tdev)
    vb = _bitmap(torch, n, 0.9, 4 + seed, tdev)
    ctx.keep.extend([base_a, base_b, va, vb])
    ops = []
    for rows in (1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24, 1 << 26, 1 << 28, 1 << 30):
        if rows > n:
            break

        def arr(t, bits, rows=rows):
            x = ag.Float32ArrayGPU(ag.ArrowGpuBuffer(dev, t.data_ptr(), rows * 4, owned=False), dev, rows, None)
            x.null_buffer = ag.NullBitBufferGpu(ag.ArrowGpuBuffer(dev, bits.data_ptr(), (rows + 31) // 32 * 4, owned=False), rows, dev)
            return x
        xa, xb = arr(base_a, va), arr(base_b, vb)
        reps = 20

        def many(xa=xa, xb=xb):
            out = None
            for _ in range(reps):
                out = xa.add(xb)
            return out
        ops.append((f"f32.add+validity rows=2^{rows.bit_length() - 1} (x{reps} back to back)", 12.375, rows * reps, many))
    return ops, {}


# ---------------------------------------------------------------------------------------------
# measurement
# ---------------------------------------------------------------------------------------------
def measure(ctx, ops, warmup, reps, min_seconds=0.0, max_reps=None, flush=None):
    """mean CUDA-event milliseconds per op (max over ranks): `warmup` untimed calls, then at least
    `reps` timed calls and until `min_seconds` of wall time have passed (so that the 50 ms clock
    sampler sees every config under load)."""
    T, dev, sharded = ctx.T, ctx.dev, ctx.sharded

    def one(fn):
        out = fn()
        del out

    for _ in range(warmup):
        for _l, _b, _r, fn in ops:
            one(fn)
    dev.sync()
    sharded.barrier()
    per = {label: [] for label, *_ in ops}
    launches0 = dev.launch_count()
    t_begin = time.time()
    done = 0
    while True:
        for label, _b, _r, fn in ops:
            if flush is not None:
                ctx.ffi.check(ctx.lib.agpu_memset(dev.handle, flush.ptr, 0, flush.size), "flush")
            if ctx.world > 1:
                sharded.barrier()   # ranks start each op together: rank skew is not the op's cost
            e0, e1 = T.event(), T.event()
            T.record(e0)
            one(fn)
            T.record(e1)
            dev.sync()
            per[label].append(T.ms(e0, e1))
            ctx.lib.agpu_event_destroy(e0)
            ctx.lib.agpu_event_destroy(e1)
        done += 1
        more = done < reps or (time.time() - t_begin) < min_seconds
        if ctx.world > 1:   # all ranks must agree on the number of rounds (there is a barrier per op)
            more = sharded.max_over_ranks(1.0 if more else 0.0) > 0
        if not more or (max_reps is not None and done >= max_reps):
            break
    launches = dev.launch_count() - launches0
    window = (t_begin, time.time())
    means = {label: sharded.max_over_ranks(statistics.mean(v)) for label, v in per.items()}
    return means, launches, window, done


def table(ctx, ops, means, peak):
    per_op, total_ms, total_rows = {}, 0.0, 0
    for label, bpr, rows_counted, _fn in ops:
        ms = means[label]
        gbs = bpr * rows_counted / (ms * 1e-3) / 1e9          # per GPU
        per_op[label] = {"ms": round(ms, 4), "rows_per_s": rows_counted * ctx.world / (ms * 1e-3), "GBps_per_gpu": round(gbs, 1),
                         "B_per_row": bpr, "frac_measured_peak": round(gbs / peak, 4), "frac_8TBps": round(gbs / 8000, 4)}
        total_ms += ms
        total_rows += rows_counted * ctx.world
    return per_op, total_ms, total_rows


def _d2h(ctx, tensor_or_ptr, nbytes, dtype):
    """first `nbytes` of a device buffer as a numpy array (through the C ABI)"""
    ptr = tensor_or_ptr.data_ptr() if hasattr(tensor_or_ptr, "data_ptr") else tensor_or_ptr
    out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
    ctx.ffi.check(ctx.lib.agpu_d2h(ctx.dev.handle, out.ctypes.data, ptr, nbytes), "d2h")
    return out


def _col_host(ctx, arr, rows):
    """(values[:rows], validity words or None) of one of our device arrays"""
    vals = _d2h(ctx, arr.data.ptr, rows * arr.NP.itemsize, arr.NP)
    valid = None
    if arr.null_buffer is not None:
        valid = _d2h(ctx, arr.null_buffer.bit_buffer.ptr, (rows + 31) // 32 * 4, np.uint32)
    return vals, valid


def _time_cpu(fn, min_seconds=0.5, max_reps=50):
    fn()
    t0, reps = time.perf_counter(), 0
    while reps < max_reps:
        fn()
        reps += 1
        if time.perf_counter() - t0 > min_seconds:
            break
    return (time.perf_counter() - t0) / reps


def _ulp_diff(got, want_f64):
    """ULP distance between f32 results and the correctly rounded f64 reference"""
    want = want_f64.astype(np.float32)
    a = got.view(np.int32).astype(np.int64)
    b = want.view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    d = np.abs(a - b)
    both_nan = np.isnan(got) & np.isnan(want)
    d[both_nan] = 0
    return d


# ---------------------------------------------------------------------------------------------
# the per-config legs: cpu_baseline + parity on a bounded sample (+ e2e)
# ---------------------------------------------------------------------------------------------
def _cfg1_block(ctx, args, peak, handles, numa):
    ag, dev, lib, ffi, T = ctx.ag, ctx.dev, ctx.lib, ctx.ffi, ctx.T
    O = _oracle()
    O.set_num_threads(len(__import__("os").sched_getaffinity(0)))
    block = {"workload": "BASELINE.json configs[0]: f32 add + gt with null bitmaps, 1 Mi rows (the reference's CPU-runnable case)",
             "l2": "the columns fit in L2, so every iteration works on ANOTHER copy of the inputs (>= 512 MiB of copies, 4x the "
                   "126 MB L2, visited round-robin): inputs come from HBM every time, and no flush kernel leaves dirty lines "
                   "whose write-back would be charged to the timed op", "sizes": {}}
    window0 = time.time()
    for n in (1 << 20, 1 << 22, 1 << 24):
        ops, ex = build_cfg1(ctx, n)
        copies = max(4, min(64, (512 << 20) // (8 * n)))
        pairs = [(ex["a"], ex["b"])]
        for _ in range(copies - 1):
            pairs.append((ex["a"].clone_array(), ex["b"].clone_array()))
        dev.sync()
        rounds = max(2, (200 if n == 1 << 20 else 60) // copies)

        def timed(fn):
            """mean CUDA-event ms of fn(a, b) over `rounds` passes over all copies"""
            for a, b in pairs[:4]:
                fn(a, b)
            dev.sync()
            ev = [(T.event(), T.event()) for _ in pairs]
            total, count = 0.0, 0
            for _ in range(rounds):
                for (a, b), (e0, e1) in zip(pairs, ev):
                    T.record(e0)
                    fn(a, b)
                    T.record(e1)
                dev.sync()
                total += sum(T.ms(e0, e1) for e0, e1 in ev)
                count += len(ev)
            for e0, e1 in ev:
                lib.agpu_event_destroy(e0)
                lib.agpu_event_destroy(e1)
            return total / count
        add_ms = timed(lambda a, b: a.add(b))
        gt_ms = timed(lambda a, b: a.gt(b))
        means = {"f32.add+validity": add_ms, "f32.gt+validity": gt_ms}
        per_op, _tot, _rows = table(ctx, ops, means, peak)
        # the same two ops recorded ONCE per copy and submitted as one CUDA graph per iteration
        progs = []
        for a, b in pairs:
            p = ag.ArrowComputePipeline(dev, "cfg1", capture=True)
            s = a.add_op(b, p)
            g = a.gt_op(b, p)
            p.finish()
            progs.append((p, s, g))
        dev.sync()
        ev = [(T.event(), T.event()) for _ in progs]
        total, count = 0.0, 0
        for _ in range(rounds):
            for (p, _s, _g), (e0, e1) in zip(progs, ev):
                T.record(e0)
                p.replay()
                T.record(e1)
            dev.sync()
            total += sum(T.ms(e0, e1) for e0, e1 in ev)
            count += len(ev)
        prog_ms = total / count
        prog_bytes = (12.375 + 8.5) * n
        # sustained: all copies' programs submitted back to back between ONE event pair (what a host
        # loop over many recorded pipelines gets); still cold inputs for every submit
        e0, e1 = ev[0]
        T.record(e0)
        for _ in range(rounds):
            for p, _s, _g in progs:
                p.replay()
        T.record(e1)
        dev.sync()
        b2b_graph_us = T.ms(e0, e1) / (rounds * len(progs)) * 1e3
        for a, b in pairs:              # untimed pass: the output blocks of the loop below are in the allocator's cache
            x = a.add(b)
            y = a.gt(b)
        dev.sync()
        T.record(e0)
        for _ in range(rounds):
            for a, b in pairs:
                x = a.add(b)
                y = a.gt(b)
        T.record(e1)
        dev.sync()
        b2b_eager_us = T.ms(e0, e1) / (rounds * len(pairs)) * 1e3
        del x, y
        # the same program with both results out of ONE kernel (agpu_fused_chain_pair: a and b are read
        # once, the validity AND is written once): what ArrowComputePipeline(fuse=True) records for
        # add_op followed by gt_op; captured, it is a one-kernel graph
        from arrow_gpu_b200 import kernels as K
        fprogs = []
        for a, b in pairs:
            p = ag.ArrowComputePipeline(dev, "cfg1 fused", fuse=True, capture=True)
            s = a.add_op(b, p)
            g = a.gt_op(b, p)
            p.finish()
            fprogs.append((p, s, g))
        dev.sync()
        T.record(e0)
        for _ in range(rounds):
            for p, _s, _g in fprogs:
                p.replay()
        T.record(e1)
        dev.sync()
        b2b_fused_graph_us = T.ms(e0, e1) / (rounds * len(fprogs)) * 1e3
        for a, b in pairs:
            x, y = K.fused_chain_pair(a, [("add", b)], [("gt", b)])
        dev.sync()
        T.record(e0)
        for _ in range(rounds):
            for a, b in pairs:
                x, y = K.fused_chain_pair(a, [("add", b)], [("gt", b)])
        T.record(e1)
        dev.sync()
        b2b_fused_eager_us = T.ms(e0, e1) / (rounds * len(pairs)) * 1e3
        del x, y
        fused_bytes = 12.5 * n          # a, b once (8), sum (4), result bitmap, two validity reads + one write (4 x 0.125)
        for e0, e1 in ev:
            lib.agpu_event_destroy(e0)
            lib.agpu_event_destroy(e1)
        entry = {"input_copies": copies, "per_op_eager": per_op,
                 "fused_pair": {"what": "add + gt as ONE kernel (agpu_fused_chain_pair): recorded on ArrowComputePipeline(fuse=True, "
                                        "capture=True) and replayed / called eagerly as kernels.fused_chain_pair; back to back, cold "
                                        "inputs; frac_of_unfused_bytes counts the 20.875 B/row the two separate ops move, "
                                        "frac_measured_peak the 12.5 B/row this kernel moves",
                                "kernels_per_submit": fprogs[0][0].graph.kernels,
                                "captured_us_per_program": round(b2b_fused_graph_us, 2),
                                "captured_frac_measured_peak": round(fused_bytes / (b2b_fused_graph_us * 1e-6) / 1e9 / peak, 4),
                                "captured_frac_of_unfused_bytes": round(prog_bytes / (b2b_fused_graph_us * 1e-6) / 1e9 / peak, 4),
                                "eager_us_per_program": round(b2b_fused_eager_us, 2),
                                "eager_frac_measured_peak": round(fused_bytes / (b2b_fused_eager_us * 1e-6) / 1e9 / peak, 4),
                                "eager_frac_of_unfused_bytes": round(prog_bytes / (b2b_fused_eager_us * 1e-6) / 1e9 / peak, 4)},
                 "captured_program": {"what": "add_op + gt_op recorded on ArrowComputePipeline(capture=True), one graph launch per iteration, "
                                              "one event pair per submit",
                                      "ms": round(prog_ms, 4), "GBps": round(prog_bytes / (prog_ms * 1e-3) / 1e9, 1),
                                      "frac_measured_peak": round(prog_bytes / (prog_ms * 1e-3) / 1e9 / peak, 4),
                                      "kernels_per_submit": progs[0][0].graph.kernels},
                 "back_to_back": {"what": "programs submitted back to back, one event pair around all of them (cold inputs each time)",
                                  "captured_us_per_program": round(b2b_graph_us, 2),
                                  "captured_frac_measured_peak": round(prog_bytes / (b2b_graph_us * 1e-6) / 1e9 / peak, 4),
                                  "eager_us_per_program": round(b2b_eager_us, 2),
                                  "eager_frac_measured_peak": round(prog_bytes / (b2b_eager_us * 1e-6) / 1e9 / peak, 4)}}
        if n == 1 << 20:
            a, b = pairs[0]
            _p, s, g = progs[0]
            # parity: full size, against the oracle; values AND validity words, eager and captured
            want_s = O.binary(O.ADD, O.F32, ex["a_h"], ex["b_h"])
            want_g = O.compare(O.GT, O.F32, ex["a_h"], ex["b_h"])
            want_v = O.validity_and(O.pack_bits(ex["va"]), O.pack_bits(ex["vb"]), n)
            bad = 0
            for got_s, got_g in ((a.add(b), a.gt(b)), (s, g), fprogs[0][1:], K.fused_chain_pair(a, [("add", b)], [("gt", b)])):
                bad += int(np.count_nonzero(got_s.raw_values().view(np.uint32) != want_s.view(np.uint32)))
                bad += int(np.count_nonzero(_d2h(ctx, got_g.data.ptr, O.words(n) * 4, np.uint32) != want_g))
                for arr in (got_s, got_g):
                    bad += int(np.count_nonzero(_d2h(ctx, arr.null_buffer.bit_buffer.ptr, O.words(n) * 4, np.uint32) != want_v))
            block["parity"] = {"ops": 2, "rows": n, "mismatches": bad, "how": "values, result bitmap and validity words; eager, captured, fused pair (captured and eager) vs the oracle"}
            cpu_s = _time_cpu(lambda: (O.binary(O.ADD, O.F32, ex["a_h"], ex["b_h"]), O.compare(O.GT, O.F32, ex["a_h"], ex["b_h"]),
                                       O.validity_and(O.pack_bits(ex["va"]), O.pack_bits(ex["vb"]), n)), 0.3, 200)
            block["cpu_baseline"] = {"value": 2 * n / cpu_s, "unit": "rows/s", "cores": O.num_threads(), "kind": "port",
                                     "sample": f"oracle add + gt + validity AND on the full {n}-row columns (the reference's lavapipe path cannot run here)"}
            # e2e: pinned host columns -> H2D -> add, gt -> D2H of sum, result bitmap, validity
            with numa.bound():
                pa, pb = dev.pinned_empty(n, np.float32), dev.pinned_empty(n, np.float32)
                pva, pvb = dev.pinned_empty(O.words(n), np.uint32), dev.pinned_empty(O.words(n), np.uint32)
                land = dev.pinned_empty(n * 4 + O.words(n) * 8, np.uint8)
            pa[:], pb[:] = ex["a_h"], ex["b_h"]
            pva[:], pvb[:] = O.pack_bits(ex["va"]).view(np.uint32), O.pack_bits(ex["vb"]).view(np.uint32)

            def e2e_once():
                def col(vals, bits):
                    buf = dev.create_gpu_buffer_with_data(vals, wait=False)
                    nb = ag.NullBitBufferGpu(dev.create_gpu_buffer_with_data(bits, wait=False), n, dev)
                    return ag.Float32ArrayGPU(buf, dev, n, nb)
                xa, xb = col(pa, pva), col(pb, pvb)
                ss, gg = xa.add(xb), xa.gt(xb)
                o = 0
                for ptr, nb_ in ((ss.data.ptr, n * 4), (gg.data.ptr, O.words(n) * 4), (ss.null_buffer.bit_buffer.ptr, O.words(n) * 4)):
                    ffi.check(lib.agpu_d2h_async(dev.handle, land.ctypes.data + o, ptr, nb_), "d2h")
                    o += nb_
                dev.sync()
                return 2 * (n * 4 + O.words(n) * 4), o
            e2e_once()
            t0 = time.perf_counter()
            for _ in range(50):
                h2d, d2h = e2e_once()
            e2e_s = (time.perf_counter() - t0) / 50
            block["e2e"] = {"value": 2 * n / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "ms_per_step": round(e2e_s * 1e3, 4)}
            for buf in (pa, pb, pva, pvb, land):
                dev.pinned_free(buf)
            del a, b
        block["sizes"][f"{n} rows"] = entry
        del progs, fprogs, pairs, ops, ex, p, s, g
        import gc
        gc.collect()
    block["_window"] = (window0, time.time())
    block["per_op"] = block["sizes"][f"{1 << 20} rows"]["per_op_eager"]
    # the same config through the compiled host mirror (what a Rust caller of the crates would see):
    # arrow_gpu_b200/cpp/bench_small links only libagpu.so; its own process, its own device handle
    import json
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arrow_gpu_b200", "cpp", "bench_small")
    if os.path.exists(exe):
        dev.sync()
        try:
            res = subprocess.run([exe, "--json"], capture_output=True, text=True, timeout=120,
                                 env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(ctx.local_rank))))
            block["cpp_mirror"] = json.loads(res.stdout.strip().splitlines()[-1]) if res.returncode == 0 else {"error": res.stderr[-300:]}
        except Exception as exc:  # noqa: BLE001
            block["cpp_mirror"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    return block


def _cfg3_block(ctx, args, peak, handles, numa, scale):
    ag, dev, lib, ffi, K = ctx.ag, ctx.dev, ctx.lib, ctx.ffi, ctx.K
    O = _oracle()
    n = max(1 << 20, int(ROWS["cfg3"] * scale))
    ops, ex = build_cfg3(ctx, n)
    cols = ex["cols"]
    means, launches, window, _reps = measure(ctx, ops, 2, 3, min_seconds=0.25)
    per_op, _tot, _rows = table(ctx, ops, means, peak)
    block = {"workload": f"BASELINE.json configs[2]: chained f32 expression (a*b+c) > d AND validity mask, {n} rows, 1 B200",
             "rows": n, "per_op": per_op, "gpu_launches": launches, "_window": window, "l2": "inputs larger than L2"}
    # parity + cpu_baseline on the first SAMPLE_ROWS rows of the FULL-SIZE run's output
    m = min(SAMPLE_ROWS, n) // 32 * 32
    host = [_col_host(ctx, c, m) for c in cols]
    got = K.fused_mul_add_gt(*cols)
    got_bits = _d2h(ctx, got.data.ptr, m // 8, np.uint32)
    got_valid = _d2h(ctx, got.null_buffer.bit_buffer.ptr, m // 8, np.uint32)
    del got

    def cpu():
        t = O.binary(O.MUL, O.F32, host[0][0], host[1][0])
        t = O.binary(O.ADD, O.F32, t, host[2][0])
        bits = O.compare(O.GT, O.F32, t, host[3][0])
        v = O.validity_and(O.validity_and(host[0][1], host[1][1], m), O.validity_and(host[2][1], host[3][1], m), m)
        return bits, v
    O.set_num_threads(len(__import__("os").sched_getaffinity(0)))
    want_bits, want_valid = cpu()
    bad = int(np.count_nonzero(got_bits != want_bits)) + int(np.count_nonzero(got_valid != want_valid))
    block["parity"] = {"ops": 1, "rows": m, "mismatches": bad,
                       "how": f"result bitmap + validity words of the first {m} rows of the full-size fused run vs the oracle's mul, add, gt, AND"}
    cpu_s = _time_cpu(cpu, 0.5, 20)
    block["cpu_baseline"] = {"value": m / cpu_s, "unit": "rows/s (expression evaluations)", "cores": O.num_threads(), "kind": "port",
                             "sample": f"oracle mul, add, gt + validity ANDs on the first {m} rows"}
    # e2e: the expression from pinned host columns (bounded: E2E_ROWS rows), results back to the host
    r = min(E2E_ROWS, n) // 1024 * 1024
    up, down = handles[1], handles[2]
    words = r // 32
    with numa.bound():
        pv = [dev.pinned_empty(r, np.float32) for _ in range(4)]
        pb = [dev.pinned_empty(words, np.uint32) for _ in range(4)]
        land = dev.pinned_empty(2 * words, np.uint32)
    for k, c in enumerate(cols):
        ffi.check(lib.agpu_d2h(dev.handle, pv[k].ctypes.data, c.data.ptr, r * 4), "d2h")
        ffi.check(lib.agpu_d2h(dev.handle, pb[k].ctypes.data, c.null_buffer.bit_buffer.ptr, words * 4), "d2h")

    def e2e_once():
        xs = []
        for k in range(4):
            buf = up.create_gpu_buffer_with_data(pv[k], wait=False)
            nb = ag.NullBitBufferGpu(up.create_gpu_buffer_with_data(pb[k], wait=False), r, dev)
            xs.append(ag.Float32ArrayGPU(buf, dev, r, nb))
        dev.wait_event(up.record_event())
        out = K.fused_mul_add_gt(*xs)
        ffi.check(lib.agpu_d2h_async(dev.handle, land.ctypes.data, out.data.ptr, words * 4), "d2h")
        ffi.check(lib.agpu_d2h_async(dev.handle, land.ctypes.data + words * 4, out.null_buffer.bit_buffer.ptr, words * 4), "d2h")
        dev.sync()
        return 4 * (r * 4 + words * 4), 2 * words * 4
    e2e_once()
    t0 = time.perf_counter()
    for _ in range(3):
        h2d, d2h = e2e_once()
    e2e_s = (time.perf_counter() - t0) / 3
    block["e2e"] = {"value": r / e2e_s, "unit": "rows/s (expression evaluations)", "rows": r, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3),
                    "sample": f"{r} of the {n} rows (4 pinned f32 columns + 4 bitmaps), fused kernel, both result bitmaps read back"}
    for buf in pv + pb + [land]:
        dev.pinned_free(buf)
    return block


def _cfg4_block(ctx, args, peak, scale):
    dev = ctx.dev
    O = _oracle()
    total = max(1 << 22, int(ROWS["cfg4"] * scale))
    b, e = ctx.sharded.row_range(total, ctx.rank, ctx.world)
    n = e - b
    ops, ex = build_cfg4(ctx, n)
    means, launches, window, _reps = measure(ctx, ops, 2, 3, min_seconds=0.25)
    per_op, tot_ms, tot_rows = table(ctx, ops, means, peak)
    block = {"workload": f"BASELINE.json configs[3]: f32 sqrt/exp/sin/cos over {total} rows, row-range sharded over {ctx.world} GPU(s)",
             "rows_total": total, "rows_per_gpu": n, "scaling": "strong", "per_op": per_op, "gpu_launches": launches, "_window": window,
             "value": tot_rows / (tot_ms * 1e-3), "unit": "rows/s", "collective": "none (independent rows)"}
    m = min(SAMPLE_ROWS, n)
    O.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0)) // ctx.world))
    ref = {"sqrt": np.sqrt, "exp": np.exp, "sin": np.sin, "cos": np.cos}
    oid = {"sqrt": O.SQRT, "exp": O.EXP, "sin": O.SIN, "cos": O.COS}
    tol = {"sqrt": 0, "exp": 2, "sin": 2, "cos": 2}
    bad, max_ulp, cpu_s = 0, {}, 0.0
    for op, arr in ex["by_op"].items():
        x = _d2h(ctx, arr.data.ptr, m * 4, np.float32)
        out = getattr(arr, op)()
        got = _d2h(ctx, out.data.ptr, m * 4, np.float32)
        del out
        d = _ulp_diff(got, ref[op](x.astype(np.float64)))
        max_ulp[op] = int(d.max())
        bad += int(np.count_nonzero(d > tol[op]))
        t0 = time.perf_counter()
        want = O.unary(oid[op], O.F32, x)
        cpu_s += time.perf_counter() - t0
        d2 = _ulp_diff(got, want.astype(np.float64))
        max_ulp[op + " vs oracle (libm)"] = int(d2.max())
    bad = int(ctx.sharded.sum_over_ranks(bad))
    block["parity"] = {"ops": 4, "rows": m, "mismatches": bad, "max_ulp": max_ulp, "tolerance_ulp": tol,
                       "how": f"first {m} rows of every rank's full-size output vs the correctly rounded f64 value (sqrt exact, "
                              "exp/sin/cos <= 2 ULP: the bound tests/test_gpu_parity.py states)"}
    if ctx.rank == 0:
        block["cpu_baseline"] = {"value": 4 * m / cpu_s, "unit": "rows/s", "cores": O.num_threads(), "kind": "port",
                                 "sample": f"oracle sqrt, exp, sin, cos (libm) on {m} rows, one pass"}
    return block


def _cfg5_block(ctx, args, peak, handles, numa, scale):
    ag, dev, lib, ffi, sharded, torch = ctx.ag, ctx.dev, ctx.lib, ctx.ffi, ctx.sharded, ctx.torch
    O = _oracle()
    total = max(1 << 22, int(ROWS["cfg5"] * scale))
    b, e = sharded.row_range(total, ctx.rank, ctx.world)
    n = e - b
    ops, ex = build_cfg5(ctx, n, total)
    means, launches, window, _reps = measure(ctx, ops, 2, 3, min_seconds=0.25)
    per_op, tot_ms, tot_rows = table(ctx, ops, means, peak)
    block = {"workload": f"BASELINE.json configs[4]: take and mask-driven merge/filter on {total} int32 rows, row-range sharded over "
                         f"{ctx.world} GPU(s), per-shard counts exchanged on the device",
             "rows_total": total, "rows_per_gpu": n, "scaling": "strong", "per_op": per_op, "gpu_launches": launches, "_window": window}
    # the exchange alone: post + wait of one u64 per rank (peer slots over NVLink), and NCCL for comparison
    if ctx.world > 1 or sharded._dist() is not None:
        ex_ctx = sharded.count_exchange(dev)
        val = dev.create_gpu_buffer_with_data(np.array([123 + ctx.rank], dtype=np.uint64))
        info = dev.create_empty_buffer((2 * ctx.world + 2) * 8)
        T = ctx.T

        def timed(fn, reps=50):
            for _ in range(5):
                fn()
            dev.sync()
            sharded.barrier()
            e0, e1 = T.event(), T.event()
            T.record(e0)
            for _ in range(reps):
                fn()
            T.record(e1)
            dev.sync()
            return sharded.max_over_ranks(T.ms(e0, e1) / reps * 1e3)

        def peer():
            ex_ctx.post(val.ptr)
            ex_ctx.wait(info.ptr)
        block["collective_us"] = {"peer slots (agpu_exchange_post + agpu_exchange_wait)": round(timed(peer), 2)}
        got = dev.retrive_data(info, (2 * ctx.world + 2) * 8).view(np.uint64)
        assert int(got[ctx.world]) == sum(123 + r for r in range(ctx.world)) and int(got[ctx.world + 1]) == 0, got
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_backend() == "nccl":
            ext = torch.cuda.ExternalStream(dev.stream_ptr, device=ctx.tdev)
            mine = torch.zeros(1, dtype=torch.int64, device=ctx.tdev)
            gathered = torch.zeros(ctx.world, dtype=torch.int64, device=ctx.tdev)

            def nccl():
                with torch.cuda.stream(ext):
                    dist.all_gather_into_tensor(gathered, mine)
            block["collective_us"]["NCCL all_gather_into_tensor on the same stream"] = round(timed(nccl), 2)
    # parity + cpu_baseline on a bounded sample: the first m rows of this rank's shard
    m = min(SAMPLE_ROWS, n) // 1024 * 1024
    O.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0)) // ctx.world))
    a, b_, m50 = ex["a"], ex["b"], ex["m50"]
    a_h = _d2h(ctx, a.data.ptr, m * 4, np.int32)
    b_h = _d2h(ctx, b_.data.ptr, m * 4, np.int32)
    bad, cpu_s, cpu_ops = 0, 0.0, 0
    mask_h = _d2h(ctx, m50.data.ptr, m // 8, np.uint32)
    out = a.merge(b_, m50)
    got = _d2h(ctx, out.data.ptr, m * 4, np.int32)
    del out
    t0 = time.perf_counter()
    want = O.merge(O.I32, a_h, b_h, mask_h)
    cpu_s += time.perf_counter() - t0
    cpu_ops += 1
    bad += int(np.count_nonzero(got != want))
    for s, mk in ex["masks"].items():
        mk_h = _d2h(ctx, mk.data.ptr, m // 8, np.uint32)
        t0 = time.perf_counter()
        want, _v, k = O.filter(O.I32, a_h, None, mk_h, None)
        cpu_s += time.perf_counter() - t0
        cpu_ops += 1
        out, _off, _tot = sharded.sharded_filter(a, mk)     # full size; its first k rows come from the first m input rows
        got = _d2h(ctx, out.data.ptr, k * 4, np.int32) if k else np.zeros(0, np.int32)
        del out
        bad += int(np.count_nonzero(got != want))
    # take: the same kernel on the sample (source = the first m rows, m random / sorted indices into it)
    a_view = ag.Int32ArrayGPU(ag.ArrowGpuBuffer(dev, a.data.ptr, m * 4, owned=False), dev, m, None)
    for label, idx_t in (("random", _randint32(torch, m, 47 + ctx.seed, ctx.tdev, 0, m)),
                         ("sorted", torch.arange(m, dtype=torch.int32, device=ctx.tdev))):
        idx = ctx.wrap(ag.UInt32ArrayGPU, idx_t)
        idx_h = _d2h(ctx, idx.data.ptr, m * 4, np.uint32)
        out = a_view.take(idx)
        got = _d2h(ctx, out.data.ptr, m * 4, np.int32)
        del out
        t0 = time.perf_counter()
        want = O.take(O.I32, a_h, m, idx_h)
        cpu_s += time.perf_counter() - t0
        cpu_ops += 1
        bad += int(np.count_nonzero(got != want))
    bad = int(sharded.sum_over_ranks(bad))
    block["parity"] = {"ops": 6, "rows": m, "mismatches": bad,
                       "how": f"merge and the three filters: the part of every rank's FULL-SIZE output that comes from its first {m} "
                              "input rows vs the oracle; take (random, sorted): the same kernel on that sample vs the oracle"}
    if ctx.rank == 0:
        block["cpu_baseline"] = {"value": cpu_ops * m / cpu_s, "unit": "rows/s", "cores": O.num_threads(), "kind": "port",
                                 "sample": f"oracle merge, filter x3, take x2 on {m} rows, one pass"}
    # e2e of filter s=0.5: pinned host column + mask -> H2D -> filter -> D2H of the kept rows
    r = min(E2E_ROWS, n) // 1024 * 1024
    mk = ex["masks"][0.5]
    with numa.bound():
        pcol, pmask = dev.pinned_empty(r, np.int32), dev.pinned_empty(r // 32, np.uint32)
        land = dev.pinned_empty(r, np.int32)
    ffi.check(lib.agpu_d2h(dev.handle, pcol.ctypes.data, a.data.ptr, r * 4), "d2h")
    ffi.check(lib.agpu_d2h(dev.handle, pmask.ctypes.data, mk.data.ptr, r // 8), "d2h")

    def e2e_once():
        col = ag.Int32ArrayGPU(dev.create_gpu_buffer_with_data(pcol, wait=False), dev, r, None)
        msk = ag.BooleanArrayGPU(dev.create_gpu_buffer_with_data(pmask, wait=False), dev, r, None)
        out = col.filter(msk)
        out.raw_values(out=land, wait=True)
        return r * 4 + r // 8, out.len * 4
    e2e_once()
    sharded.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        h2d, d2h = e2e_once()
    e2e_s = sharded.max_over_ranks((time.perf_counter() - t0) / 3)
    block["e2e"] = {"value": r * ctx.world / e2e_s, "unit": "rows/s (input rows filtered)", "rows_per_gpu": r, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s * 1e3, 3),
                    "sample": f"filter s=0.5 on {r} rows per GPU from a pinned host column + mask, kept rows read back"}
    for buf in (pcol, pmask, land):
        dev.pinned_free(buf)
    return block


def per_config(args, rank, world, local_rank, helpers, handles):
    """the blocks of the default bench line for BASELINE.json configs 1, 3, 4, 5"""
    peak, _src = helpers["peak"]()
    numa = helpers["numa"]
    scale = args.per_config_scale
    out = {"note": "per-GPU GB/s; strong-scaled configs split BASELINE.json's row count over the ranks; clocks sampled per config; "
                   f"cpu_baseline / parity on samples of <= {SAMPLE_ROWS} rows (the reference's per-op maximum, gpu_device.rs:69,133)"}
    todo = [("cfg4", lambda c: _cfg4_block(c, args, peak, scale)), ("cfg5", lambda c: _cfg5_block(c, args, peak, handles, numa, scale))]
    if world == 1:
        todo = [("cfg1", lambda c: _cfg1_block(c, args, peak, handles, numa)),
                ("cfg3", lambda c: _cfg3_block(c, args, peak, handles, numa, scale))] + todo
    only = getattr(args, "per_config_only", None)
    if only:
        todo = [t for t in todo if t[0] in only.split(",")]
    for name, fn in todo:
        ctx = Ctx(rank, world, local_rank, dev=handles[0])
        t0 = time.time()
        try:
            block = fn(ctx)
        except Exception as exc:  # noqa: BLE001 — one config must not take the whole bench line down
            import sys
            import traceback
            traceback.print_exc(file=sys.stderr)
            block = {"error": f"{type(exc).__name__}: {exc}"[:400], "trace": traceback.format_exc()[-1200:]}
            h = C.c_void_p()   # an exception inside a captured pipeline leaves the stream capturing: close it
            if ctx.lib.agpu_graph_end(ctx.dev.handle, C.byref(h)) == 0 and h:
                ctx.lib.agpu_graph_destroy(h)
        block["seconds"] = round(time.time() - t0, 1)
        out[name] = block
        ctx.release()
        for h in handles[1:]:
            ctx.lib.agpu_trim(h.handle)
    return out


# ---------------------------------------------------------------------------------------------
# one workload alone (bench.py --workload ...)
# ---------------------------------------------------------------------------------------------
def run(args, rank, world, local_rank, helpers):
    """returns the JSON dict for --workload cfg1|cfg3|cfg4|cfg5|allops|sweep"""
    ctx = Ctx(rank, world, local_rank)
    dev, sharded = ctx.dev, ctx.sharded
    name = args.workload
    total_rows = args.rows if args.rows_given else ROWS[name]
    if name in ("cfg4", "cfg5"):
        b, e = sharded.row_range(total_rows, rank, world)   # strong scaling: the named column is split
        n = e - b
        scaling = "strong"
    else:
        n = total_rows
        scaling = "weak"
    if name == "cfg1":
        ops, _ex = build_cfg1(ctx, n)
    elif name == "cfg3":
        ops, _ex = build_cfg3(ctx, n)
    elif name == "cfg4":
        ops, _ex = build_cfg4(ctx, n)
    elif name == "cfg5":
        ops, _ex = build_cfg5(ctx, n, total_rows)
    elif name == "allops":
        ops, _ex = build_allops(ctx, n)
    elif name == "sweep":
        ops, _ex = build_sweep(ctx, n)
    else:
        raise SystemExit(f"unknown workload {name}")
    ctx.torch.cuda.synchronize()
    flush = dev.create_empty_buffer(512 << 20) if name == "cfg1" else None   # columns fit in L2: flush between iterations
    means, launches, window, _reps = measure(ctx, ops, args.warmup, args.steps, min_seconds=0.3, flush=flush)
    sharded.barrier()
    peak, peak_src = helpers["peak"]()
    per_op, total_ms, total_rows_done = table(ctx, ops, means, peak)
    worst = min(per_op, key=lambda k: per_op[k]["frac_measured_peak"])
    clocks = helpers["sampler"].window(*window)
    return {
        "metric": f"rows/s (row-operations per second over the ops of {name}; achieved HBM GB/s per op in per_op)",
        "value": total_rows_done / (total_ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(total_ms, 4), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32" if name != "cfg5" else "i32", "data": "synthetic (generated on device, seeded)",
        "config": {"workload": f"BASELINE.json {name}", "rows_total": total_rows, "rows_per_gpu": n,
                   "l2": "L2 flushed between iterations" if flush is not None else "inputs larger than L2"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": None,
        "roofline": {"bound": "hbm", "kernel": worst, "achieved": per_op[worst]["GBps_per_gpu"], "peak": peak, "unit": "GB/s",
                     "frac": per_op[worst]["frac_measured_peak"], "traffic": helpers["traffic"](worst), "peak_source": peak_src},
        "per_op": per_op, "cpu_baseline": None,
    }
