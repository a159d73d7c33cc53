#!/usr/bin/env python3
"""bench.py — rows/s and achieved HBM GB/s per op of the columnar compute hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (`value`, N = 1): BASELINE.json configs[1] — "i8/u8/i16/u16 arithmetic, logical
and cast across types, 256M rows, 1 B200" with the op list of SURVEY.md §8(d) cfg 2.  One STEP =
one pass of all 55 ops over their 268 435 456-row synthetic columns.
  value      row-operations per second (ops x rows / device time), inputs resident in HBM, the K
             steps launched back to back (no events inside the timed region)
  per_op     the same ops timed one by one with CUDA events on the launching stream
  roofline   the op FURTHEST BELOW the HBM roofline (lowest algorithmic GB/s / measured copy peak)
  e2e        the same step through the public array API from pinned HOST buffers: H2D of every
             input column on an upload stream, D2H of every output column on a download stream
  parity     every one of the 55 full-size outputs compared bit for bit with the oracle's output on
             the same host columns (non-timed); a mismatch fails the run (exit code 1)
  cpu_baseline  the oracle port (OpenMP) on the host cores, rank 0, full columns
  per_config the OTHER BASELINE.json configs, each with its own per-op table, clocks window,
             cpu_baseline, parity on a bounded sample and (cfg 1, 3, 5-filter) e2e:
             cfg1 (1 Mi rows f32 add + gt with nulls; eager and as ONE captured submit), cfg3 (fused
             (a*b+c)>d, 1 G rows) — N = 1 only; cfg4 (f32 sqrt/exp/sin/cos over 4 G rows, row-range
             sharded: strong scaling) and cfg5 (merge / filter with the device-side count exchange /
             take on 4 G int32 rows, sharded; `collective_us` per exchange) at every N.
N > 1: cfg 2 is weak-scaled — every rank runs the same step on its own row-range shard, no
collective on the data path, time = max over ranks.

`--impl reference` times the reference's CPU stand-in (oracle/, the C restatement of its shaders;
the reference itself is Rust+WGSL on wgpu/lavapipe and cannot be built in this image) on the host
cores for the same config/metric.  oracle/ is imported ONLY in that arm, in the cpu_baseline legs
and in the parity checks.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_CFG2 = 268_435_456
SIZES = {"i8": 1, "u8": 1, "i16": 2, "u16": 2, "i32": 4, "u32": 4, "f32": 4}
NPT = {"i8": np.int8, "u8": np.uint8, "i16": np.int16, "u16": np.uint16, "i32": np.int32, "u32": np.uint32,
       "f32": np.float32}
METRIC = "rows/s (row-operations per second over the 55 ops of config 2; achieved HBM GB/s per op in per_op)"
DTYPE = "u8/i8/u16/i16 (+u32 shift counts, f32 for casts)"


def config_dict(rows: int):
    """the SAME dict in both arms (the driver compares them)"""
    return {"workload": "BASELINE.json configs[1]: i8/u8/i16/u16 arithmetic, logical, shift and cast, "
                        f"{rows} rows per GPU, 55 ops per step", "rows_per_gpu": rows, "ops_per_step": 55,
            "l2": "inputs larger than L2 (every column >= 256 MiB vs 126 MB L2)", "sharding": "row-range, no collective"}


# ---------------------------------------------------------------------------------------------
# workload description (shared by both arms): (label, kind, dtype, op, dst dtype)
# ---------------------------------------------------------------------------------------------
def cfg2_ops():
    ops = []
    for t in ("i8", "u8", "i16", "u16"):
        for op in ("add", "sub", "mul"):
            ops.append((f"{t}.{op}", "binary", t, op, t))
        for op in ("add", "mul"):
            ops.append((f"{t}.{op}_scalar", "scalar", t, op, t))
        for op in ("and", "or", "xor"):
            ops.append((f"{t}.{op}", "binary", t, op, t))
        ops.append((f"{t}.not", "unary", t, "not", t))
        for op in ("shl", "shr"):
            ops.append((f"{t}.{op}", "shift", t, op, t))
    for s, d in (("i8", "i16"), ("i8", "i32"), ("i8", "f32"), ("u8", "u16"), ("u8", "u32"), ("u8", "f32"),
                 ("i16", "i32"), ("i16", "f32"), ("u16", "u32"), ("u16", "f32"), ("f32", "u8")):
        ops.append((f"cast.{s}->{d}", "cast", s, "cast", d))
    return ops


def pinned_by_reference(spec) -> bool:
    """False for the ops BASELINE.json config 2 names but the reference does not implement
    (sub-word + - x, array and scalar; only `u16 + scalar` exists: arithmetic/compute_shaders/u16/
    scalar.wgsl:15-23).  For those the oracle is the definition — parity unpinned."""
    label, kind, _t, op, _d = spec
    if kind == "binary" and op in ("add", "sub", "mul"):
        return False
    if kind == "scalar":
        return label == "u16.add_scalar"
    return True


def bytes_per_row(spec) -> float:
    _label, kind, t, _op, d = spec
    es, ds = SIZES[t], SIZES[d]
    return {"binary": 3 * es, "scalar": 2 * es, "unary": 2 * es, "shift": 2 * es + 4, "cast": es + ds}[kind]


def cfg2_columns(rows: int, seed0: int = 10, out=None):
    """synthetic columns of SURVEY.md §8(d) cfg 2: full-range uniform ints (seed 10+k), shift
    counts U{0..width-1}, f32 U(-10, 70000) for the narrowing cast.  `out(name, n, dtype)` may
    provide the destination arrays (pinned host memory)."""
    cols = {}

    def put(name, values):
        if out is None:
            cols[name] = values
        else:
            dst = out(name, len(values), values.dtype)
            dst[:] = values
            cols[name] = dst

    for k, t in enumerate(("i8", "u8", "i16", "u16")):
        info = np.iinfo(NPT[t])
        rng = np.random.default_rng(seed0 + k)
        put(f"{t}.a", rng.integers(info.min, int(info.max) + 1, rows, dtype=NPT[t]))
        put(f"{t}.b", rng.integers(info.min, int(info.max) + 1, rows, dtype=NPT[t]))
    rng = np.random.default_rng(seed0 + 8)
    put("cnt8", rng.integers(0, 8, rows, dtype=np.uint32))
    put("cnt16", rng.integers(0, 16, rows, dtype=np.uint32))
    put("f32.a", rng.uniform(-10, 70000, rows).astype(np.float32))
    return cols


def inputs_of(spec):
    _label, kind, t, _op, _d = spec
    if kind == "binary":
        return [f"{t}.a", f"{t}.b"]
    if kind == "shift":
        return [f"{t}.a", "cnt8" if SIZES[t] == 1 else "cnt16"]
    return [f"{t}.a"]


# ---------------------------------------------------------------------------------------------
# our arm: the public array API over the C ABI
# ---------------------------------------------------------------------------------------------
def gpu_runner():
    import arrow_gpu_b200 as ag
    cls = {"i8": ag.Int8ArrayGPU, "u8": ag.UInt8ArrayGPU, "i16": ag.Int16ArrayGPU, "u16": ag.UInt16ArrayGPU,
           "i32": ag.Int32ArrayGPU, "u32": ag.UInt32ArrayGPU, "f32": ag.Float32ArrayGPU}
    meth = {"add": "add", "sub": "sub", "mul": "mul", "and": "bitwise_and", "or": "bitwise_or", "xor": "bitwise_xor",
            "shl": "bitwise_shl", "shr": "bitwise_shr", "not": "bitwise_not"}

    def run(spec, arrs, scalars):
        _label, kind, t, op, d = spec
        ins = [arrs[c] for c in inputs_of(spec)]
        if kind == "binary" or kind == "shift":
            return getattr(ins[0], meth[op])(ins[1])
        if kind == "scalar":
            return getattr(ins[0], f"{op}_scalar")(scalars[t])
        if kind == "unary":
            return getattr(ins[0], meth[op])()
        return ins[0].cast(cls[d])
    return ag, cls, run


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled for the WHOLE run, every 20 ms, through
    NVML — the library behind the recipe's `nvidia-smi --query-gpu=clocks.sm,...` line
    (B200_PROFILING.md); a pipe to an nvidia-smi child loses its last block-buffered samples when
    the child is killed, and a 100 ms period misses millisecond-long configs.  `window(t0, t1)`
    summarises the samples between two time.time() marks, so every config reports the clocks seen
    during ITS timed region.  Falls back to one nvidia-smi child (line-buffered, 50 ms)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.02):
        import threading
        self.samples, self.proc, self.thread, self.source = [], None, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn(h)

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        r = int(reasons_fn(h))
                        self.samples.append((time.time(), sm, mx, [n for n, b in bits.items() if r & b]))
                    except Exception:  # noqa: BLE001
                        pass
                    self._stop.wait(period_s)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = f"NVML (pynvml), every {int(period_s * 1e3)} ms"
        except Exception:  # noqa: BLE001 — no NVML binding: the nvidia-smi child
            try:
                cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)]
                import shutil
                if shutil.which("stdbuf"):
                    cmd = ["stdbuf", "-oL"] + cmd
                self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi -lms 50"
            except OSError:
                self.source = None
        self._stopped = False

    def stop(self):
        if self._stopped:
            return
        self._stopped = True
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                t = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                sm, mx = float(f[2]), float(f[3])
            except ValueError:
                continue
            self.samples.append((t, sm, mx, [n for n, v in zip(self.NAMES, f[6:10]) if v.lower().startswith("active")]))

    def window(self, t0: float, t1: float):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML / nvidia-smi"]}
        if self.proc is not None:
            self.stop()
        samples = list(self.samples)
        inside = [s for s in samples if t0 <= s[0] <= t1]
        nearest = False
        if not inside and samples:   # a window shorter than the sampling period: the nearest sample
            mid = 0.5 * (t0 + t1)
            inside = [min(samples, key=lambda s: abs(s[0] - mid))]
            nearest = True
        reasons = sorted({r for s in inside for r in s[3]})
        out = {"sm_mhz": statistics.median(s[1] for s in inside) if inside else None,
               "sm_max_mhz": max((s[2] for s in inside), default=None), "samples": len(inside), "reasons": reasons,
               "window_s": round(t1 - t0, 3), "source": self.source}
        if nearest:
            out["note"] = "window shorter than the sampling period: nearest sample"
        return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(label: str):
    """dram bytes per launch of the roofline kernel from the committed ncu --set full capture"""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(label)
    return None


class NumaBinding:
    """`with numa.bound():` pins the calling thread to the CPUs next to this rank's GPU while pinned
    staging buffers are allocated (first touch places their pages on that NUMA node, so the PCIe
    traffic of the 8 ranks does not cross the socket interconnect); everything else — the launch
    loop, the OpenMP CPU legs — runs with the process's full affinity."""

    def __init__(self, index: int):
        self.full = os.sched_getaffinity(0)
        self.cpus, self.info = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & self.full
            if cpus and cpus != self.full:
                self.cpus = cpus
                self.info = {"bound": True, "cpus": len(cpus), "of": len(self.full)}
            else:
                self.info = {"bound": False, "cpus": len(self.full), "why": "GPU-local CPU set == process affinity"}
        except Exception as e:  # noqa: BLE001 — NVML absent / not permitted: run unbound, say so
            self.info = {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}

    def bound(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            if self.cpus:
                os.sched_setaffinity(0, self.cpus)
            try:
                yield
            finally:
                if self.cpus:
                    os.sched_setaffinity(0, self.full)
        return ctx()


class Events:
    def __init__(self, dev, lib, ffi):
        import ctypes as C
        self.C, self.dev, self.lib, self.ffi = C, dev, lib, ffi

    def new(self):
        e = self.C.c_void_p()
        self.ffi.check(self.lib.agpu_event_create(self.C.byref(e)), "event_create")
        return e

    def record(self, e, dev=None):
        self.ffi.check(self.lib.agpu_event_record((dev or self.dev).handle, e), "event_record")

    def ms(self, a, b):
        out = self.C.c_float(0)
        self.ffi.check(self.lib.agpu_event_elapsed_ms(a, b, self.C.byref(out)), "event_elapsed")
        return out.value


def link_ceiling_probe(dev, up, down, lib, ffi, sharded, h2d_bytes, d2h_bytes, src_pinned, dst_pinned, dev_buf):
    """What the host <-> device links give ALL ranks at once with plain cudaMemcpyAsync from/to pinned
    memory and nothing else running: the e2e step's own byte counts in both directions, H2D and
    D2H concurrently on two streams.  e2e cannot beat this; `link_frac` says how close it gets."""
    dst_list = dst_pinned if isinstance(dst_pinned, (list, tuple)) else [dst_pinned]   # the e2e step alternates landing buffers
    chunk = min(len(src_pinned), min(len(d) for d in dst_list), dev_buf.size)

    def once():
        sent = 0
        while sent < h2d_bytes:
            n = min(chunk, h2d_bytes - sent)
            ffi.check(lib.agpu_h2d(up.handle, dev_buf.ptr, src_pinned.ctypes.data, n), "h2d")
            sent += n
        got, k = 0, 0
        while got < d2h_bytes:
            n = min(chunk, d2h_bytes - got)
            ffi.check(lib.agpu_d2h_async(down.handle, dst_list[k % len(dst_list)].ctypes.data, dev_buf.ptr, n), "d2h")
            got += n
            k += 1
        up.sync()
        down.sync()

    once()
    sharded.barrier()
    t0 = time.perf_counter()
    once()
    ms = sharded.max_over_ranks((time.perf_counter() - t0) * 1e3)
    return ms


def run_ours(args, rank, world, local_rank, sampler, numa):
    import ctypes as C
    from arrow_gpu_b200 import _ffi, sharded
    ag, cls, run = gpu_runner()
    dev = ag.GpuDevice(local_rank)
    lib = _ffi.lib()
    ev = Events(dev, lib, _ffi)
    rows = args.rows
    ops = cfg2_ops()

    # pinned host staging: inputs (copied every e2e step) and two output landing buffers
    with numa.bound():
        pinned = cfg2_columns(rows, seed0=10 + 100 * rank, out=lambda _name, n, dt: dev.pinned_empty(n, dt))
        landing = [dev.pinned_empty(rows * 4, np.uint8) for _ in range(2)]

    col_cls = {"cnt8": cls["u32"], "cnt16": cls["u32"]}

    def upload(name, wait):
        c = col_cls.get(name) or cls[name.split(".")[0]]
        return c.from_numpy(pinned[name], None, dev, wait=wait)

    arrs = {name: upload(name, True) for name in pinned}
    scalars = {t: cls[t].from_slice([3], dev) for t in ("i8", "u8", "i16", "u16")}

    # ---- resident-input throughput (`value`): K steps back to back, nothing but kernels in the
    # timed region (consecutive streaming kernels overlap tail and ramp: programmatic dependent launch)
    def resident_step(events=None):
        for k, spec in enumerate(ops):
            out = run(spec, arrs, scalars)
            del out  # stream-ordered free: the pool hands the block to the next op
            if events is not None:
                ev.record(events[k + 1])

    t_window0 = time.time()
    for _ in range(args.warmup):
        resident_step()
    dev.sync()
    sharded.barrier()
    launches0 = dev.launch_count()
    t_start, t_stop = ev.new(), ev.new()
    dev.sync()
    ev.record(t_start)
    for _ in range(args.steps):
        resident_step()
    ev.record(t_stop)
    dev.sync()
    launches = dev.launch_count() - launches0
    total_ms = ev.ms(t_start, t_stop)
    sharded.barrier()
    total_ms = sharded.max_over_ranks(total_ms)

    # ---- the same ops one by one (CUDA events between them: no overlap across ops)
    step_events = [[ev.new() for _ in range(len(ops) + 1)] for _ in range(args.steps)]
    for s in range(args.steps):
        ev.record(step_events[s][0])
        resident_step(step_events[s])
    dev.sync()
    t_window1 = time.time()
    # median over the K passes: one host hiccup (GC, a page fault) while the queue is shallow lands
    # in whichever op was waiting for its launch and would otherwise dominate its mean
    per_op_ms = [statistics.median(ev.ms(step_events[s][k], step_events[s][k + 1]) for s in range(args.steps))
                 for k in range(len(ops))]
    peak, peak_src = measured_peak()
    per_op = {}
    for spec, ms in zip(ops, per_op_ms):
        ms = sharded.max_over_ranks(ms)
        gbs = bytes_per_row(spec) * rows / (ms * 1e-3) / 1e9
        per_op[spec[0]] = {"ms": round(ms, 4), "rows_per_s": rows / (ms * 1e-3), "GBps": round(gbs, 1),
                           "B_per_row": bytes_per_row(spec), "frac_measured_peak": round(gbs / peak, 4),
                           "frac_8TBps": round(gbs / 8000.0, 4), "pinned": pinned_by_reference(spec)}
    # the roofline line names the op FURTHEST BELOW the roofline, not the one that moves most bytes
    wl = min(per_op, key=lambda k: per_op[k]["frac_measured_peak"])
    step_bytes = sum(bytes_per_row(s) for s in ops) * rows
    roof = {"bound": "hbm", "kernel": wl, "achieved": per_op[wl]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": per_op[wl]["frac_measured_peak"], "traffic": known_traffic(wl), "peak_source": peak_src,
            "share_of_step": round(per_op[wl]["ms"] / sum(v["ms"] for v in per_op.values()), 4),
            "step_mean_frac": round(step_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peak, 4),
            "step_mean_frac_one_by_one": round(step_bytes / (sum(v["ms"] for v in per_op.values()) * 1e-3) / 1e9 / peak, 4),
            "below_target": sorted(k for k, v in per_op.items() if v["frac_8TBps"] < 0.80),
            "target": ">= 0.80 of 8 TB/s per op (BASELINE.json north_star); the measured copy peak itself is "
                      f"{peak / 8000:.3f} of 8 TB/s"}
    row_ops = len(ops) * rows * args.steps
    value = row_ops * world / (total_ms * 1e-3)

    # ---- end to end through the public API from pinned host buffers ----
    # Three handles = three streams on the GPU: `up` uploads the input columns, `dev` computes (each
    # op waits on the GPU for the upload events of its inputs), `down` copies every result column to
    # one of two pinned landing buffers.  PCIe is full duplex; the allocator keeps a block that
    # another stream still uses out of circulation until that stream is done (agpu_buffer_record_use).
    up, down = ag.GpuDevice(local_rank), ag.GpuDevice(local_rank)
    order = []
    for spec in ops:
        for c in inputs_of(spec):
            if c not in order:
                order.append(c)
    done_ev = [ag.GpuEvent() for _ in ops]

    def e2e_step(read_back=True):
        live, ready = {}, {}
        for name in order:
            c = col_cls.get(name) or cls[name.split(".")[0]]
            arr = c.from_numpy(pinned[name], None, up, wait=False)
            ready[name] = up.record_event()
            arr.gpu_device = dev          # ops on this column run on the compute handle
            live[name] = arr
        h2d = sum(a.nbytes for a in pinned.values())
        d2h = 0
        waited = set()
        for k, spec in enumerate(ops):
            for c in inputs_of(spec):
                if c not in waited:
                    dev.wait_event(ready[c])
                    waited.add(c)
            out = run(spec, live, scalars)
            if read_back:
                nbytes = out.len * out.NP.itemsize
                dev.record_event(done_ev[k])
                down.wait_event(done_ev[k])
                down.record_use(out.data)
                _ffi.check(lib.agpu_d2h_async(down.handle, landing[k & 1].ctypes.data, out.data.ptr, nbytes), "d2h")
                d2h += nbytes
        dev.sync()
        up.sync()
        down.sync()
        return h2d, d2h

    e2e = None
    if not args.no_e2e:
        e2e_step()
        sharded.barrier()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))

        def timed_e2e():
            t0 = time.perf_counter()
            each = []
            for _ in range(e2e_steps):
                t1 = time.perf_counter()
                moved = e2e_step()
                each.append(round((time.perf_counter() - t1) * 1e3, 1))
            return sharded.max_over_ranks((time.perf_counter() - t0) * 1e3), each, moved

        def e2e_block(e2e_ms, each, moved):
            h2d, d2h = moved
            step_ms = e2e_ms / e2e_steps
            return {"value": len(ops) * rows * e2e_steps * world / (e2e_ms * 1e-3), "unit": "rows/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(step_ms, 3),
                    "pcie_GBps": {"h2d": round(h2d / (step_ms * 1e-3) / 1e9, 1), "d2h": round(d2h / (step_ms * 1e-3) / 1e9, 1)},
                    "ms_each_step_rank0": each,
                    "streams": "upload / compute / download on three handles, two landing buffers"}

        e2e = e2e_block(*timed_e2e())
        h2d, d2h = e2e["h2d_bytes_per_step"], e2e["d2h_bytes_per_step"]
        # where the e2e time goes: the same step with the 55 result columns left on the device
        t0 = time.perf_counter()
        e2e_step(read_back=False)
        e2e["upload_and_compute_only_ms_per_step"] = round(sharded.max_over_ranks((time.perf_counter() - t0) * 1e3), 3)
        # and what the links give all ranks at once for these byte counts with nothing else going on
        scratch = dev.create_empty_buffer(rows * 4)
        probe_ms = link_ceiling_probe(dev, up, down, lib, _ffi, sharded, h2d, d2h, pinned["cnt8"].view(np.uint8),
                                      landing, scratch)
        del scratch
        e2e["link_probe_ms_per_step"] = round(probe_ms, 3)
        e2e["link_GBps_ceiling"] = {"h2d": round(h2d / (probe_ms * 1e-3) / 1e9, 1), "d2h": round(d2h / (probe_ms * 1e-3) / 1e9, 1),
                                    "how": "plain cudaMemcpyAsync, pinned, H2D and D2H concurrently, all ranks at once, "
                                           "the e2e step's byte counts"}
        e2e["link_frac"] = round(probe_ms / e2e["ms_per_step"], 4)
        # The host links are shared with whatever else runs on the box.  A measurement far below what
        # the same links gave seconds later (plain copies of the same bytes) saw an outside
        # disturbance: like a run with a thermal slowdown it is rejected and re-measured ONCE; the
        # rejected attempt stays in the line.
        if e2e["link_frac"] < args.e2e_remeasure_below:
            first = {k: e2e[k] for k in ("value", "ms_per_step", "ms_each_step_rank0", "link_frac")}
            again = e2e_block(*timed_e2e())
            again["link_frac"] = round(probe_ms / again["ms_per_step"], 4)
            for k in ("upload_and_compute_only_ms_per_step", "link_probe_ms_per_step", "link_GBps_ceiling"):
                again[k] = e2e[k]
            again["remeasured_once"] = {"why": f"first attempt ran at < {args.e2e_remeasure_below} of the link ceiling measured right after it",
                                        "rejected_attempt": first}
            e2e = again

    # ---- full-size parity: each of the 55 outputs vs the oracle on the same host columns ----
    parity = None
    if not args.no_parity:
        parity = parity_pass(ops, run, arrs, scalars, pinned, landing[0], rows, world)
        parity["mismatches"] = int(sharded.sum_over_ranks(parity["mismatches"]))
        parity["ranks_checked"] = world
        parity["unpinned_ops"] = sorted(s[0] for s in ops if not pinned_by_reference(s))

    run_ours.host_columns = pinned     # reused by the cpu_baseline leg (same synthetic columns)
    run_ours.handles = (dev, up, down)
    run_ours.landing = landing
    result = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_dict(rows),
        "gpu_launches": int(launches), "clocks": None, "e2e": e2e, "roofline": roof, "parity": parity, "per_op": per_op,
    }
    run_ours.clock_window = (t_window0, t_window1)
    return result


def parity_pass(ops, run, arrs, scalars, host_cols, landing, rows, world):
    """GPU output of every op (resident inputs = uploads of `host_cols`) vs the oracle on `host_cols`."""
    O, cpu_run = cpu_runner()
    O.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))   # every rank checks its own shard at once
    outs = {t: np.empty(rows, dtype=NPT[t]) for t in ("i8", "u8", "i16", "u16", "i32", "u32", "f32")}
    bad_ops, mismatches = {}, 0
    t0 = time.perf_counter()
    for spec in ops:
        got_dev = run(spec, arrs, scalars)
        view = landing[: got_dev.len * got_dev.NP.itemsize].view(got_dev.NP)
        got_dev.raw_values(out=view, wait=True)
        want = cpu_run(spec, host_cols, outs)
        # bit patterns (f32 results of the int -> f32 casts are exact, so NaN never occurs)
        a = view.view(np.uint8)
        b = want.view(np.uint8)
        if not np.array_equal(a, b):
            wrong = int(np.count_nonzero(view.view(f"u{view.dtype.itemsize}") != want.view(f"u{want.dtype.itemsize}")))
            bad_ops[spec[0]] = wrong
            mismatches += wrong
        del got_dev
    return {"ops": len(ops), "rows": rows, "mismatches": mismatches, "bad_ops": bad_ops,
            "seconds": round(time.perf_counter() - t0, 1),
            "how": "bit-exact compare of every full-size output column (D2H) with oracle/oracle.c on the same host columns"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (OpenMP) on the host cores — reference stand-in and cpu_baseline
# ---------------------------------------------------------------------------------------------
def cpu_runner():
    import oracle as O
    dt = {"i8": O.I8, "u8": O.U8, "i16": O.I16, "u16": O.U16, "i32": O.I32, "u32": O.U32, "f32": O.F32}
    binop = {"add": O.ADD, "sub": O.SUB, "mul": O.MUL, "and": O.AND, "or": O.OR, "xor": O.XOR}

    def run(spec, cols, outs):
        _label, kind, t, op, d = spec
        ins = [cols[c] for c in inputs_of(spec)]
        out = outs[d][: len(ins[0])]
        if kind == "binary":
            return O.binary(binop[op], dt[t], ins[0], ins[1], out=out)
        if kind == "scalar":
            return O.scalar(binop[op], dt[t], ins[0], 3, out=out)
        if kind == "unary":
            return O.unary(O.NOT, dt[t], ins[0], out=out)
        if kind == "shift":
            return O.shift(O.SHL if op == "shl" else O.SHR, dt[t], ins[0], ins[1], out=out)
        return O.cast(dt[t], dt[d], ins[0], out=out)
    return O, run


def time_cpu(sample_rows: int, steps: int, warmup: int, cols=None):
    O, run = cpu_runner()
    # all host threads this process may use — torchrun exports OMP_NUM_THREADS=1 to its workers,
    # which would otherwise time the CPU arm on a single core
    O.set_num_threads(len(os.sched_getaffinity(0)))
    ops = cfg2_ops()
    if cols is None:
        cols = cfg2_columns(sample_rows)
    else:
        cols = {k: v[:sample_rows] for k, v in cols.items()}
    outs = {t: np.empty(sample_rows, dtype=NPT[t]) for t in NPT}
    for _ in range(warmup):
        for spec in ops:
            run(spec, cols, outs)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for spec in ops:
            run(spec, cols, outs)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": len(ops) * sample_rows * steps / total, "unit": "rows/s", "cores": O.num_threads(),
            "kind": "port", "sample": f"{len(ops)} ops x {sample_rows} rows x {steps} steps of config 2 "
                                      f"(oracle/oracle.c, OpenMP, {O.num_threads()} threads; the reference's own "
                                      "wgpu/lavapipe path cannot be built here)",
            "ms_per_step": round(total / steps * 1e3, 2)}


def time_arrow_cpu(cols, sample_rows: int = 1 << 24):
    """cross-check named by the north star: Arrow's own CPU kernels (pyarrow.compute, the C++ Arrow
    library; arrow-rs is not in this image) on the ops of config 2 whose semantics they share —
    wrapping add/sub/mul, and/or/xor/not, widening casts; one thread per call, bounded sample"""
    try:
        import pyarrow as pa
        import pyarrow.compute as pc
    except ImportError:
        return None
    fn = {"add": pc.add, "sub": pc.subtract, "mul": pc.multiply, "and": pc.bit_wise_and, "or": pc.bit_wise_or,
          "xor": pc.bit_wise_xor, "not": pc.bit_wise_not}
    pat = {"i8": pa.int8(), "u8": pa.uint8(), "i16": pa.int16(), "u16": pa.uint16(), "i32": pa.int32(),
           "u32": pa.uint32(), "f32": pa.float32()}
    arrs = {k: pa.array(v[:sample_rows]) for k, v in cols.items()}
    todo = [s for s in cfg2_ops() if s[1] in ("binary", "scalar", "unary") or (s[1] == "cast" and s[2] != "f32")]
    t0 = time.perf_counter()
    for _label, kind, t, op, d in todo:
        a = arrs[f"{t}.a"]
        if kind == "binary":
            fn[op](a, arrs[f"{t}.b"])
        elif kind == "scalar":
            fn[op](a, pa.scalar(3, pat[t]))
        elif kind == "unary":
            fn[op](a)
        else:
            pc.cast(a, pat[d])
    dt = time.perf_counter() - t0
    return {"value": len(todo) * sample_rows / dt, "unit": "rows/s", "cores": 1, "ops": len(todo),
            "sample": f"pyarrow.compute {pa.__version__}, {len(todo)} of the 55 ops x {sample_rows} rows, one pass"}


def run_reference(args, rank):
    if rank != 0:
        return None
    base = time_cpu(args.cpu_rows, args.steps, args.warmup)
    return {
        "impl": "reference", "metric": METRIC,
        "value": base["value"], "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": config_dict(args.rows),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=None, help="rows (default: the workload's own size; cfg2: 256 Mi per GPU)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "sweep", "allops"],
                    help="BASELINE.json configs[k-1]; cfg2 is the bench line (with the others in per_config); naming "
                         "another one runs only that config (profiling)")
    ap.add_argument("--cpu-rows", type=int, default=None,
                    help="rows of the CPU sample (default: the full columns — a step is 1-4 s of CPU work on 16-32 host "
                         "threads, and a sample that fits the host's last-level cache would flatter the CPU)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-remeasure-below", type=float, default=0.6,
                    help="re-measure e2e once when it ran below this fraction of the link ceiling (outside disturbance)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--per-config-only", default=None, help="debugging: skip config 2 and run only these per_config blocks (cfg1,cfg3,...)")
    ap.add_argument("--per-config-scale", type=float, default=1.0,
                    help="shrink the per_config row counts (smoke runs on small GPUs); 1.0 = BASELINE.json sizes")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that chat on fd 1 (NCCL prints its version
    # banner there) are sent to stderr, the result goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()
    args.rows_given = args.rows is not None
    if args.rows is None:
        args.rows = ROWS_CFG2
    if args.cpu_rows is None:
        args.cpu_rows = args.rows
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup   # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        res = run_reference(args, rank)
        if res is not None:
            emit(res)
        return 0

    numa = NumaBinding(local_rank)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rc = 0
    sampler = ClockSampler(local_rank)
    helpers = {"sampler": sampler, "peak": measured_peak, "traffic": known_traffic, "numa": numa}
    try:
        if args.workload != "cfg2":
            import bench_workloads
            res = bench_workloads.run(args, rank, world, local_rank, helpers)
            if rank == 0:
                emit(res)
            return 0
        from arrow_gpu_b200 import sharded
        if args.per_config_only:
            import arrow_gpu_b200 as ag
            import bench_workloads
            handles = tuple(ag.GpuDevice(local_rank) for _ in range(3))
            per_config = bench_workloads.per_config(args, rank, world, local_rank, helpers, handles)
            for block in per_config.values():
                if isinstance(block, dict) and "_window" in block:
                    block["clocks"] = sampler.window(*block.pop("_window"))
            if rank == 0:
                emit({"per_config": per_config})
            return 0
        res = run_ours(args, rank, world, local_rank, sampler, numa)
        res["host_numa"] = numa.info
        cpu = None
        if not args.no_cpu_baseline:
            sharded.barrier()
            if rank == 0:   # the other ranks idle at the barrier below: the CPU leg has the host to itself
                cpu = time_cpu(min(args.cpu_rows, args.rows), 3, 1, cols=run_ours.host_columns)
                cpu["arrow_cross_check"] = time_arrow_cpu(run_ours.host_columns, min(1 << 24, args.rows))
            sharded.barrier()
        # free config 2 before the other configs take the GPU
        dev = run_ours.handles[0]
        for name in list(run_ours.host_columns):
            dev.pinned_free(run_ours.host_columns.pop(name))
        for buf in run_ours.landing:
            dev.pinned_free(buf)
        run_ours.landing = []
        import gc
        gc.collect()
        per_config = None
        if not args.no_per_config:
            import bench_workloads
            per_config = bench_workloads.per_config(args, rank, world, local_rank, helpers, run_ours.handles)
        sampler.stop()
        res["clocks"] = sampler.window(*run_ours.clock_window)
        res["clocks"]["window"] = "warm-up + timed steps + per-op pass of config 2"
        if per_config is not None:
            for name, block in per_config.items():
                if isinstance(block, dict) and "_window" in block:
                    block["clocks"] = sampler.window(*block.pop("_window"))
        res["per_config"] = per_config
        res["cpu_baseline"] = cpu
        bad = (res.get("parity") or {}).get("mismatches", 0)
        for block in (per_config or {}).values():
            if isinstance(block, dict):
                bad += (block.get("parity") or {}).get("mismatches", 0)
        if bad:
            res["parity_failed"] = True
            rc = 1
        if rank == 0:
            emit(res)
    finally:
        sampler.stop()
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
