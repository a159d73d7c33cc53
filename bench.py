#!/usr/bin/env python3
"""bench.py — rows/s and achieved HBM GB/s per op of the columnar compute hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default, N=1): BASELINE.json configs[1] — "i8/u8/i16/u16 arithmetic, logical and cast
across types, 256M rows, 1 B200" with the op list of SURVEY.md §8(d) cfg 2.  One STEP = one pass
of all 55 ops over their 268 435 456-row synthetic columns.  `value` = row-operations per second
(ops x rows / device time) with inputs resident in HBM; `e2e` = the same step through the public
array API from pinned HOST buffers (H2D of every input column, D2H of every output column inside
the timed region); `roofline` = algorithmic bytes of the slowest op / its CUDA-event time against
the measured HBM copy peak; `cpu_baseline` = the oracle port (OpenMP) on a bounded sample.
N > 1: weak scaling — every rank runs the same step on its own row-range shard, no collective on
the data path (element-wise ops shard with zero communication), time = max over ranks.

`--impl reference` times the reference's CPU stand-in (oracle/, the C restatement of its shaders;
the reference itself is Rust+WGSL on wgpu/lavapipe and cannot be built in this image) on the host
cores for the same config/metric.  oracle/ is imported ONLY in that arm and in the cpu_baseline leg.
"""
from __future__ import annotations

import argparse
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


def bytes_per_row(spec) -> float:
    _label, kind, t, _op, d = spec
    es, ds = SIZES[t], SIZES[d]
    return {"binary": 3 * es, "scalar": 2 * es, "unary": 2 * es, "shift": 2 * es + 4, "cast": es + ds}[kind]


def cfg2_columns(rows: int, seed0: int = 10):
    """synthetic columns of SURVEY.md §8(d) cfg 2: full-range uniform ints (seed 10+k), shift
    counts U{0..width-1}, f32 U(-10, 70000) for the narrowing cast"""
    cols = {}
    for k, t in enumerate(("i8", "u8", "i16", "u16")):
        info = np.iinfo(NPT[t])
        rng = np.random.default_rng(seed0 + k)
        cols[f"{t}.a"] = rng.integers(info.min, int(info.max) + 1, rows, dtype=NPT[t])
        cols[f"{t}.b"] = rng.integers(info.min, int(info.max) + 1, rows, dtype=NPT[t])
    rng = np.random.default_rng(seed0 + 8)
    cols["cnt8"] = rng.integers(0, 8, rows, dtype=np.uint32)
    cols["cnt16"] = rng.integers(0, 16, rows, dtype=np.uint32)
    cols["f32.a"] = rng.uniform(-10, 70000, rows).astype(np.float32)
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


def sample_clocks_start(index: int):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return None


def sample_clocks_stop(proc):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=10)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, _ = proc.communicate()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in out.splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1]))
            mx.append(float(f[2]))
        except ValueError:
            continue
        for name, val in zip(names, f[5:9]):
            if val.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm), "reasons": sorted(reasons)}


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


def run_ours(args, rank, world, local_rank):
    import ctypes as C
    from arrow_gpu_b200 import _ffi, sharded
    ag, cls, run = gpu_runner()
    dev = ag.GpuDevice(local_rank)
    lib = _ffi.lib()
    rows = args.rows
    ops = cfg2_ops()
    host = cfg2_columns(rows, seed0=10 + 100 * rank)

    # pinned host staging: inputs (copied every e2e step) and one output landing buffer
    pinned = {}
    for name, arr in host.items():
        p = dev.pinned_empty(len(arr), arr.dtype)
        p[:] = arr
        pinned[name] = p
    del host
    out_stage = dev.pinned_empty(rows * 4, np.uint8)

    col_cls = {"cnt8": cls["u32"], "cnt16": cls["u32"]}

    def upload(name, wait):
        c = col_cls.get(name) or cls[name.split(".")[0]]
        return c.from_numpy(pinned[name], None, dev, wait=wait)

    arrs = {name: upload(name, True) for name in pinned}
    scalars = {t: cls[t].from_slice([3], dev) for t in ("i8", "u8", "i16", "u16")}

    def new_event():
        e = C.c_void_p()
        _ffi.check(lib.agpu_event_create(C.byref(e)), "event_create")
        return e

    def record(e):
        _ffi.check(lib.agpu_event_record(dev.handle, e), "event_record")

    def elapsed(a, b):
        ms = C.c_float(0)
        _ffi.check(lib.agpu_event_elapsed_ms(a, b, C.byref(ms)), "event_elapsed")
        return ms.value

    # ---- resident-input throughput (`value`) with per-op CUDA events on the launching stream ----
    def resident_step(events=None):
        for k, spec in enumerate(ops):
            out = run(spec, arrs, scalars)
            del out  # stream-ordered free: the pool hands the block to the next op
            if events is not None:
                record(events[k + 1])

    # the sampler runs from the warm-up on (same load as the timed steps): the timed region of
    # K=5 steps lasts ~50 ms, shorter than nvidia-smi's sampling period
    clocks_proc = sample_clocks_start(local_rank)
    for _ in range(max(args.warmup, 1) * 8):
        resident_step()
    dev.sync()
    sharded.barrier()
    step_events = [[new_event() for _ in range(len(ops) + 1)] for _ in range(args.steps)]
    launches0 = dev.launch_count()
    t_start, t_stop = new_event(), new_event()
    dev.sync()
    record(t_start)
    for s in range(args.steps):
        record(step_events[s][0])
        resident_step(step_events[s])
    record(t_stop)
    dev.sync()
    launches = dev.launch_count() - launches0
    total_ms = elapsed(t_start, t_stop)
    clocks = sample_clocks_stop(clocks_proc)
    sharded.barrier()
    total_ms = sharded.max_over_ranks(total_ms)

    per_op_ms = [statistics.mean(elapsed(step_events[s][k], step_events[s][k + 1]) for s in range(args.steps))
                 for k in range(len(ops))]
    peak, peak_src = measured_peak()
    per_op = {}
    for spec, ms in zip(ops, per_op_ms):
        gbs = bytes_per_row(spec) * rows / (ms * 1e-3) / 1e9
        per_op[spec[0]] = {"ms": round(ms, 4), "rows_per_s": rows / (ms * 1e-3), "GBps": round(gbs, 1),
                           "B_per_row": bytes_per_row(spec), "frac_measured_peak": round(gbs / peak, 4),
                           "frac_8TBps": round(gbs / 8000.0, 4)}
    # dominant kernel = the op family that takes the largest share of the step
    worst = max(range(len(ops)), key=lambda k: per_op_ms[k])
    wl = ops[worst][0]
    roof = {"bound": "hbm", "kernel": wl, "achieved": per_op[wl]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": round(per_op[wl]["GBps"] / peak, 4), "traffic": known_traffic(wl), "peak_source": peak_src,
            "share_of_step": round(per_op_ms[worst] / sum(per_op_ms), 4),
            "step_mean_frac": round(sum(bytes_per_row(s) for s in ops) * rows / (sum(per_op_ms) * 1e-3) / 1e9 / peak, 4)}
    row_ops = len(ops) * rows * args.steps
    value = row_ops * world / (total_ms * 1e-3)

    # ---- end to end through the public API from pinned host buffers ----
    # A second device handle (= a second stream on the same GPU) uploads the input columns while
    # the compute handle works; each op waits on the GPU for the upload events of its inputs.
    # Outputs are read back on the compute stream into a pinned landing buffer.  PCIe is full
    # duplex, so H2D hides behind the (larger) D2H traffic.
    up = ag.GpuDevice(local_rank)
    order = []
    for spec in ops:
        for c in inputs_of(spec):
            if c not in order:
                order.append(c)

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
        for spec in ops:
            for c in inputs_of(spec):
                if c not in waited:
                    dev.wait_event(ready[c])
                    waited.add(c)
            out = run(spec, live, scalars)
            if read_back:
                view = out_stage[: out.len * out.NP.itemsize].view(out.NP)
                out.raw_values(out=view, wait=False)
                d2h += view.nbytes
        dev.sync()
        up.sync()
        return h2d, d2h

    e2e = None
    if not args.no_e2e:
        e2e_step()
        sharded.barrier()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        ev0, ev1 = new_event(), new_event()
        record(ev0)
        for _ in range(e2e_steps):
            h2d, d2h = e2e_step()
        record(ev1)
        dev.sync()
        e2e_ms = sharded.max_over_ranks(max(elapsed(ev0, ev1), (time.perf_counter() - t0) * 1e3))
        e2e = {"value": len(ops) * rows * e2e_steps * world / (e2e_ms * 1e-3), "unit": "rows/s", "steps": e2e_steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms / e2e_steps, 3),
               "pcie_GBps": {"h2d": round(h2d / (e2e_ms / e2e_steps * 1e-3) / 1e9, 1),
                             "d2h": round(d2h / (e2e_ms / e2e_steps * 1e-3) / 1e9, 1)}}
        # where the e2e time goes: the same step with the 55 result columns left on the device
        # (uploads + compute only) — not the headline, it shows that e2e is the D2H link's time
        t0 = time.perf_counter()
        e2e_step(read_back=False)
        e2e["upload_and_compute_only_ms_per_step"] = round(
            sharded.max_over_ranks((time.perf_counter() - t0) * 1e3), 3)

    run_ours.host_columns = pinned     # reused by the cpu_baseline leg (same synthetic columns)
    result = {
        "metric": "rows/s (row-operations per second over the 55 ops of config 2; achieved HBM GB/s per op in per_op)",
        "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/i8/u16/i16 (+u32 shift counts, f32 for casts)", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: i8/u8/i16/u16 arithmetic, logical, shift and cast, "
                               f"{rows} rows per GPU, 55 ops per step", "rows_per_gpu": rows, "ops_per_step": len(ops),
                   "l2": "inputs larger than L2 (every column >= 256 MiB vs 126 MB L2)", "sharding": "row-range, no collective"},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "per_op": per_op,
    }
    return result


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
        "impl": "reference",
        "metric": "rows/s (row-operations per second over the 55 ops of config 2; achieved HBM GB/s per op in per_op)",
        "value": base["value"], "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/i8/u16/i16 (+u32 shift counts, f32 for casts)", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: i8/u8/i16/u16 arithmetic, logical, shift and cast, "
                               f"bounded sample of {args.cpu_rows} rows per step, 55 ops per step",
                   "ops_per_step": 55},
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
                    help="BASELINE.json configs[k-1]; cfg2 is the bench line, the others are scaling/parity configs")
    ap.add_argument("--cpu-rows", type=int, default=ROWS_CFG2,
                    help="rows of the CPU sample (default: the full 256 Mi-row columns — a step is 1-4 s of CPU work on "
                         "16-24 host threads, and a sample that fits the host's last-level cache would flatter the CPU)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        res = run_reference(args, rank)
        if res is not None:
            emit(res)
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload != "cfg2":
            import bench_workloads
            res = bench_workloads.run(args, rank, world, local_rank,
                                      {"clocks_start": sample_clocks_start, "clocks_stop": sample_clocks_stop,
                                       "peak": measured_peak, "traffic": known_traffic})
            if rank == 0:
                emit(res)
            return
        res = run_ours(args, rank, world, local_rank)
        if rank == 0:
            if not args.no_cpu_baseline and world == 1:
                res["cpu_baseline"] = time_cpu(min(args.cpu_rows, args.rows), 3, 1, cols=run_ours.host_columns)
                res["cpu_baseline"]["arrow_cross_check"] = time_arrow_cpu(run_ours.host_columns,
                                                                          min(1 << 24, args.rows))
            else:
                res["cpu_baseline"] = None
            emit(res)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
