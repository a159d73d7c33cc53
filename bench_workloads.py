"""The other BASELINE.json configurations (cfg1, cfg3, cfg4, cfg5) for `bench.py --workload`.

These are the parity-test / scaling configurations; the default bench line is config 2 (bench.py).
Columns of 1-4 G rows are synthesised ON THE DEVICE with torch (seeded generators) — torch is used
only as an allocator/RNG here; every measured kernel is ours, launched through the C ABI on the
library's own stream and timed with CUDA events on that stream.  No e2e number is produced for
these workloads (their inputs never exist on the host), except cfg1.
"""
from __future__ import annotations

import ctypes as C
import statistics

ROWS = {"allops": 268_435_456, "cfg1": 1_048_576, "cfg3": 1_000_000_000, "cfg4": 4_000_000_000, "cfg5": 4_000_000_000, "sweep": 1 << 30}


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


def run(args, rank, world, local_rank, helpers):
    """returns the JSON dict for --workload cfg1|cfg3|cfg4|cfg5"""
    import torch

    import arrow_gpu_b200 as ag
    from arrow_gpu_b200 import _ffi, kernels as K, sharded

    torch.cuda.set_device(local_rank)
    tdev = torch.device("cuda", local_rank)
    dev = ag.GpuDevice(local_rank)
    lib = _ffi.lib()
    T = Timer(dev, lib, _ffi)
    keep = []
    name = args.workload
    total_rows = args.rows if args.rows_given else ROWS[name]
    if name in ("cfg4", "cfg5"):
        b, e = sharded.row_range(total_rows, rank, world)   # strong scaling: the named column is split
        n = e - b
        scaling = "strong"
    else:
        n = total_rows
        scaling = "weak"
    seed = 1000 * rank

    ops = []  # (label, bytes_per_row (per input row unless noted), rows_counted, fn)

    def nullable(arr, p, s):
        bits = _bitmap(torch, n, p, s, tdev)
        keep.append(bits)
        arr.null_buffer = ag.NullBitBufferGpu(ag.ArrowGpuBuffer(dev, bits.data_ptr(), bits.numel(), owned=False), n, dev)
        return arr

    if name == "cfg1":
        a = nullable(_wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, -1000, 1000, 1 + seed, tdev), n, dev, keep), 0.9, 2 + seed)
        b_ = nullable(_wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, -1000, 1000, 11 + seed, tdev), n, dev, keep), 0.9, 3 + seed)
        ops = [("f32.add+validity", 12.375, n, lambda: a.add(b_)), ("f32.gt+validity", 8.5, n, lambda: a.gt(b_))]
    elif name == "cfg3":
        cols = [nullable(_wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, -10, 10, 20 + k + seed, tdev), n, dev, keep), 0.95, 24 + k + seed)
                for k in range(4)]

        def chain():
            p = ag.ArrowComputePipeline(dev, "chain")
            r = K.gt_op_dyn(K.add_op_dyn(K.mul_op_dyn(cols[0], cols[1], p), cols[2], p), cols[3], p)
            p.finish()
            return r
        def chain_fused():
            p = ag.ArrowComputePipeline(dev, "chain", fuse=True)
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
               ("unfused chain mul,add,gt (reference style, 3 kernels)", 33.25, n, chain),
               ("same recorded chain on ArrowComputePipeline(fuse=True) (auto-fused -> the dedicated kernel)", 16.75, n, chain_fused),
               ("generic fused_chain [sin, mul b, add c] -> f32 + 3 bitmaps", 16.5, n,
                lambda: K.fused_chain(cols[0], [("sin",), ("mul", cols[1]), ("add", cols[2])])),
               ("unfused sin,mul,add (3 kernels)", 32.875, n, chain2)]
    elif name == "cfg4":
        rng = {"sqrt": (0, 1e6), "exp": (-20, 20), "sin": (-100, 100), "cos": (-100, 100)}
        col = {}
        for k, (op, (lo, hi)) in enumerate(rng.items()):
            if (lo, hi) not in col:
                col[(lo, hi)] = _wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, lo, hi, 30 + k + seed, tdev), n, dev, keep)
            arr = col[(lo, hi)]
            ops.append((f"f32.{op}", 8.0, n, (lambda arr=arr, op=op: getattr(arr, op)())))
    elif name == "cfg5":
        a = _wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 40 + seed, tdev), n, dev, keep)
        b_ = _wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 140 + seed, tdev), n, dev, keep)

        def mask(p, s):
            bits = _bitmap(torch, n, p, s, tdev)
            keep.append(bits)
            return ag.BooleanArrayGPU(ag.ArrowGpuBuffer(dev, bits.data_ptr(), bits.numel(), owned=False), dev, n, None)
        m50 = mask(0.5, 41 + seed)
        ops.append(("i32.merge", 12.125, n, lambda: a.merge(b_, m50)))
        for s, sd in ((0.1, 42), (0.5, 43), (0.9, 44)):
            mk = mask(s, sd + seed)

            def filt(mk=mk):
                out, _off, _tot = sharded.sharded_filter(a, mk)   # count exchange over NCCL when world > 1
                return out
            ops.append((f"i32.filter s={s}", 4.125 + 4 * s, n, filt))
        seq = torch.arange(n, dtype=torch.int64, device=tdev).to(torch.int32)
        idx_seq = _wrap(ag, ag.UInt32ArrayGPU, seq, n, dev, keep)
        idx_rnd = _wrap(ag, ag.UInt32ArrayGPU, _randint32(torch, n, 45 + seed, tdev, 0, n), n, dev, keep)
        ops.append(("i32.take sorted stride-1", 12.0, n, lambda: a.take(idx_seq)))
        ops.append(("i32.take uniform random", 12.0, n, lambda: a.take(idx_rnd)))
        if world > 1:
            # global row numbers over all shards: the gather kernel reads peer shards over NVLink
            col = sharded.ShardedColumn(ag.Int32ArrayGPU, a, None, total_rows, dev)
            keep.append(col)
            m_g = min(n, 1 << 28)
            gidx = _wrap(ag, ag.UInt32ArrayGPU, _randint32(torch, m_g, 46 + seed, tdev, 0, min(total_rows, 2**31 - 1)), m_g, dev, keep)
            ops.append((f"i32.take GLOBAL uniform random over {world} shards (NVLink peer loads, {m_g} rows/GPU)", 12.0, m_g,
                        lambda: col.take_global(gidx)))
    elif name == "allops":
        # every remaining (op, dtype) family of the path at 256 Mi rows, to find kernels that fall
        # short of the roofline (config 2 already covers sub-word arithmetic/logical/shift/cast)
        f = [_wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, 0.5, 50.0, 60 + k + seed, tdev), n, dev, keep) for k in range(2)]
        f[0] = nullable(f[0], 0.9, 70 + seed)
        unit = _wrap(ag, ag.Float32ArrayGPU, _uniform(torch, n, -1.0, 1.0, 63 + seed, tdev), n, dev, keep)
        i32 = [_wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 64 + k + seed, tdev), n, dev, keep) for k in range(2)]
        u32 = [_wrap(ag, ag.UInt32ArrayGPU, _randint32(torch, n, 66 + k + seed, tdev), n, dev, keep) for k in range(2)]
        small = _wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 68 + seed, tdev, -6, 12), n, dev, keep)
        cnt = _wrap(ag, ag.UInt32ArrayGPU, _randint32(torch, n, 69 + seed, tdev, 0, 32), n, dev, keep)

        def sub(t, cls):
            g = torch.Generator(device=tdev)
            g.manual_seed(80 + seed)
            info = torch.iinfo(t)
            x = torch.randint(info.min, info.max + 1, (n,), generator=g, device=tdev, dtype=torch.int32).to(t)
            return _wrap(ag, cls, x, n, dev, keep)
        i8a, i8b = sub(torch.int8, ag.Int8ArrayGPU), sub(torch.int8, ag.Int8ArrayGPU)
        u16a, u16b = sub(torch.int16, ag.UInt16ArrayGPU), sub(torch.int16, ag.UInt16ArrayGPU)
        mbits = _bitmap(torch, n, 0.5, 90 + seed, tdev)
        keep.append(mbits)
        m = ag.BooleanArrayGPU(ag.ArrowGpuBuffer(dev, mbits.data_ptr(), mbits.numel(), owned=False), dev, n, None)
        m2 = ag.BooleanArrayGPU(ag.ArrowGpuBuffer(dev, mbits.data_ptr(), mbits.numel(), owned=False), dev, n, None)
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
        keep.append(seq)
        idx = ag.UInt32ArrayGPU(ag.ArrowGpuBuffer(dev, seq.data_ptr(), n * 4, owned=False), dev, n, None)
        dst = ag.Int32ArrayGPU.empty(n, dev)
        ops += [
            ("f32.take sequential (+validity gather)", 12.25, n, lambda: f[0].take(idx)),
            ("i8.take sequential", 6, n, lambda: i8a.take(idx)),
            ("bool.take sequential", 4.25, n, lambda: m.take(idx)),
            ("i32.put sequential (one index column used for both sides)", 12, n, lambda: i32[0].put(idx, dst, idx)),
            ("f32.broadcast", 4, n, lambda: ag.Float32ArrayGPU.broadcast(1.5, n, dev)),
        ]
        # fused integer chains (agpu_fused_chain_int) and the same ops one kernel each
        from arrow_gpu_b200 import kernels as K
        sc8 = ag.Int8ArrayGPU.from_slice([3], dev)
        sc16 = ag.UInt16ArrayGPU.from_slice([3], dev)
        i8c = sub(torch.int8, ag.Int8ArrayGPU)
        i32c = _wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 95 + seed, tdev), n, dev, keep)
        i32d = _wrap(ag, ag.Int32ArrayGPU, _randint32(torch, n, 96 + seed, tdev), n, dev, keep)
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
    elif name == "sweep":
        # column-size sweep of one binary op with validity (f32 add, 12.375 B/row) from the
        # reference's test sizes up to 1 Gi rows: where launch latency ends and HBM begins.
        # 20 calls back to back per measurement (no host sync in between).
        base_a = _uniform(torch, n, -1000, 1000, 1 + seed, tdev)
        base_b = _uniform(torch, n, -1000, 1000, 2 + seed, tdev)
        va = _bitmap(torch, n, 0.9, 3 + seed, tdev)
        vb = _bitmap(torch, n, 0.9, 4 + seed, tdev)
        keep.extend([base_a, base_b, va, vb])
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
    else:
        raise SystemExit(f"unknown workload {name}")

    torch.cuda.synchronize()
    flush = None
    if name == "cfg1":   # columns fit in L2: flush it between timed iterations
        flush = dev.create_empty_buffer(512 << 20)

    def one(fn):
        out = fn()
        del out

    for _ in range(args.warmup):
        for _l, _b, _r, fn in ops:
            one(fn)
    dev.sync()
    sharded.barrier()
    clocks_proc = helpers["clocks_start"](local_rank)
    launches0 = dev.launch_count()
    per = {label: [] for label, *_ in ops}
    for _ in range(args.steps):
        for label, _b, _r, fn in ops:
            if flush is not None:
                _ffi.check(lib.agpu_memset(dev.handle, flush.ptr, 0, flush.size), "flush")
            if world > 1:
                sharded.barrier()   # ranks start each op together: rank skew is not the op's cost
            e0, e1 = T.event(), T.event()
            T.record(e0)
            one(fn)
            T.record(e1)
            dev.sync()
            per[label].append(T.ms(e0, e1))
    launches = dev.launch_count() - launches0
    clocks = helpers["clocks_stop"](clocks_proc)
    sharded.barrier()

    peak, peak_src = helpers["peak"]()
    per_op, total_ms, total_rows_done = {}, 0.0, 0
    for label, bpr, rows_counted, _fn in ops:
        ms = sharded.max_over_ranks(statistics.mean(per[label]))
        gbs = bpr * rows_counted / (ms * 1e-3) / 1e9          # per GPU
        per_op[label] = {"ms": round(ms, 4), "rows_per_s": rows_counted * world / (ms * 1e-3), "GBps_per_gpu": round(gbs, 1),
                         "B_per_row": bpr, "frac_measured_peak": round(gbs / peak, 4), "frac_8TBps": round(gbs / 8000, 4)}
        total_ms += ms
        total_rows_done += rows_counted * world
    worst = max(per_op, key=lambda k: per_op[k]["ms"])
    res = {
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
    return res
