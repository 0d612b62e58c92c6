#!/usr/bin/env python
"""bench.py — candidate-sites/sec of the Clair3 network forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload pileup|fa]

A *step* is one forward of the hot path over one synthetic candidate-site batch of the configuration the metric is
quoted on: pileup ``1024x33x18`` int32 (``BASELINE.json`` configs[1]; ``fa`` = configs[2], ``256x89x33x8`` int8).
Timed region: K steps issued round-robin over a few CUDA streams of ONE model (each stream owns an activation
workspace), inputs already resident in HBM, rotated over > 126 MB of distinct batches so no step re-reads its input
from L2; bracketed by barrier + synchronize, timed with CUDA events, max over ranks.  ``e2e`` repeats the measurement
through the reference-facing module call (``Clair3_P.__call__``) with pinned HOST input and HOST output, H2D and D2H
inside the timed region.  N>1: one process per GPU (torchrun), sites sharded with no data-path collective, one
weight broadcast from rank 0 before the timed region (``scaling: weak``).

``--impl reference`` times the reference's own CPU implementation of the same step (torch CPU ops, all host threads)
through ``oracle/torch_port.py`` (the Python reference cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from clair3_b200 import synth  # noqa: E402

FLOP_PER_SITE = {"pileup": 47_785_984, "fa": 451_538_432}
# algorithmic FLOPs per site of each tensor-core kernel (2*M*N*K of the layer shapes, clair3/model.py:96-110, 317-344)
KERNEL_FLOP_PER_SITE = {
    "pileup": {"lstm1": 33 * 2 * 2 * 512 * (18 + 128), "proj2": 33 * 2 * 2 * 640 * 256, "lstm2": 33 * 2 * 2 * 640 * 160,
               "l4": 2 * 10560 * 128},
    "fa": {"conv0": 2 * 45 * 17 * 64 * 72, "conv1": 2 * 45 * 17 * 64 * 576, "conv2": 2 * 45 * 17 * 64 * 576,
           "conv3": 2 * 23 * 9 * 128 * 576, "conv4": 2 * 23 * 9 * 128 * 1152, "conv5": 2 * 23 * 9 * 128 * 1152,
           "conv6": 2 * 12 * 5 * 256 * 1152, "conv7": 2 * 12 * 5 * 256 * 2304, "conv8": 2 * 12 * 5 * 256 * 2304,
           "l4": 2 * 3584 * 256},
}
BATCH = {"pileup": 1024, "fa": 256}
LSTM_TILE = [0]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "which": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "which": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_model(workload, device, load_real_weights):
    from clair3_b200.model import Clair3_F, Clair3_P
    if workload == "pileup":
        sd = synth.pileup_state_dict(False, seed=0)
        m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
    else:
        sd = synth.fa_state_dict(True, channels=8, seed=0)
        m = Clair3_F(add_indel_length=True, predict=True, input_channels=8)
    if not load_real_weights:      # non-root ranks start from zeros and receive the packed image by broadcast
        sd = {k: np.zeros_like(v) for k, v in sd.items()}
    m.to(device)
    m.eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return m, sd


def make_inputs(workload, n_batches, seed):
    b = BATCH[workload]
    if workload == "pileup":
        return [synth.pileup_inputs(b, seed=seed + i) for i in range(n_batches)]
    base = [synth.fa_inputs(b, depth=89, channels=8, seed=seed + i) for i in range(min(n_batches, 4))]
    out = []
    for i in range(n_batches):        # cheap distinct batches: roll sites of a few generated ones
        out.append(np.roll(base[i % len(base)], i // len(base) + 1, axis=0))
    return out


def cpu_port(workload, sd):
    from oracle import torch_port          # cpu_baseline / reference arm only
    if workload == "pileup":
        return torch_port.PileupPort(sd, False)
    return torch_port.FullAlignmentPort(sd, True)


def best_cpu_threads(port, x, ncores):
    """The reference lets torch pick its thread count (CallVariantsFromCffi.py:56-63 sets it from --threads); oneDNN's
    LSTM/conv primitives do not scale to every core of a large host, so probe a few counts and keep the fastest."""
    best, best_t = ncores, None
    for n in sorted({c for c in (8, 16, 32, 64, ncores) if c <= ncores}):
        torch.set_num_threads(n)
        port(x)
        t0 = time.perf_counter()
        port(x)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def time_cpu(workload, sd, budget_s, threads, min_iters=2):
    port = cpu_port(workload, sd)
    xs = make_inputs(workload, 2, seed=900)
    threads = best_cpu_threads(port, xs[0], threads)
    port(xs[0])                            # warm-up
    t0 = time.perf_counter()
    iters = 0
    while iters < min_iters or (time.perf_counter() - t0 < budget_s and iters < 64):
        port(xs[iters % 2])
        iters += 1
    dt = time.perf_counter() - t0
    return BATCH[workload] * iters / dt, iters, dt, threads


def cpu_deployment_shape(workload, ncores, seconds=6.0, steps=0):
    """The reference's own CPU deployment: many single-threaded worker processes (`--threads N` -> N*3/4 callers with
    torch.set_num_threads(1), scripts/clair3_c_impl.sh + CallVariantsFromCffi.py:56-63).  Runs min(3/4 cores, 64) processes of
    the oracle port concurrently - for `seconds` each, or (steps > 0) exactly `steps` calls sized to take about `seconds` - and
    sums their rates."""
    nproc = max(1, min(ncores * 3 // 4, 64))
    sites = 64 if workload == "pileup" else 8
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    root = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "oracle.torch_port", workload, str(seconds), str(sites)] + ([str(steps)] if steps > 0 else [])
    procs = [subprocess.Popen(cmd, cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(nproc)]
    total, ok, per_call, secs = 0.0, 0, 0, 0.0
    for p in procs:
        try:
            out, _ = p.communicate(timeout=seconds * 10 + 180)
            r = json.loads(out.strip().splitlines()[-1])
            total += r["sites"] / r["seconds"]
            per_call = r["sites_per_call"]
            secs = max(secs, r["seconds"])
            ok += 1
        except Exception:
            p.kill()
    return {"value": total, "unit": "sites/s", "processes": ok, "threads_each": 1, "seconds": secs,
            "sample": "%d single-thread processes, %s of ~%d-site calls each, rates summed"
                      % (ok, ("%d calls" % steps) if steps > 0 else ("%.0f s" % seconds), per_call)}


def run_reference_arm(args, rank, world):
    """The reference's CPU path on this box's host cores (rank 0 only under torchrun)."""
    if rank != 0:
        return
    workload = args.workload
    sd = synth.pileup_state_dict(False, seed=0) if workload == "pileup" else synth.fa_state_dict(True, channels=8, seed=0)
    port = cpu_port(workload, sd)
    xs = make_inputs(workload, 2, seed=900)
    threads = best_cpu_threads(port, xs[0], len(os.sched_getaffinity(0)))
    for i in range(args.warmup):
        port(xs[i % 2])
    # a "step" is a bounded sample of the workload's batch: the whole K-step run must end within ~2 minutes on the host cores
    t0 = time.perf_counter()
    port(xs[0])
    t1 = time.perf_counter() - t0
    n_sites = BATCH[workload]
    if args.steps * t1 > 120.0:
        n_sites = max(16, int(BATCH[workload] * 120.0 / (args.steps * t1)) // 16 * 16)
        xs = [x[:n_sites] for x in xs]
        port(xs[0])
    t0 = time.perf_counter()
    for i in range(args.steps):
        port(xs[i % 2])
    dt = time.perf_counter() - t0
    val = n_sites * args.steps / dt
    single = {"value": val, "cores": threads, "sample": "%d steps of %d sites in one process" % (args.steps, n_sites)}
    dep = cpu_deployment_shape(workload, len(os.sched_getaffinity(0)), seconds=20.0, steps=args.steps)
    cores, sample, ms_step = threads, "%d steps of %d sites, torch CPU ops of the reference forward (oracle/torch_port.py)" % (args.steps, n_sites), dt / args.steps * 1e3
    if dep["value"] > val:            # all the host threads the reference can use: its many-single-thread-workers deployment
        val, cores, sample = dep["value"], dep["processes"], dep["sample"] + ", torch CPU ops of the reference forward (oracle/torch_port.py)"
        ms_step = dep["seconds"] / args.steps * 1e3
    line = {
        "impl": "reference", "metric": "candidate-sites/sec", "value": val, "unit": "sites/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(workload, 0, 0),
        "cpu_baseline": {"value": val, "unit": "sites/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_process": single, "deployment_shape": dep},
        "e2e": {"value": val, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(workload, streams, pool):
    if workload == "pileup":
        w = "Pileup net forward, synthetic batch 1024x33x18 int32 per step (BASELINE configs[1])"
    else:
        w = "Full-alignment net forward, synthetic batch 256x89x33x8 int8 per step (BASELINE configs[2])"
    return {"workload": w, "batch_per_step": BATCH[workload], "streams_in_flight": streams, "lstm_subtile_sites": LSTM_TILE[0],
            "l2_policy": "inputs rotated over %d distinct device-resident batches (> 126 MB L2)" % pool if pool else "n/a",
            "weights": "seeded synthetic checkpoint (clair3_b200.synth), random-init of the reference architecture",
            "parallelism": "site-sharded, one process per GPU"}


def timed_steps(model, xs_dev, ys_dev, streams, steps, warmup, device):
    """K forwards round-robin over the streams; returns elapsed ms measured with CUDA events."""
    main = torch.cuda.current_stream(device)
    def issue(n, offset):
        for i in range(n):
            st = streams[(offset + i) % len(streams)]
            with torch.cuda.stream(st):
                j = (offset + i) % len(xs_dev)
                ys_dev[(offset + i) % len(ys_dev)] = model(xs_dev[j])
    issue(warmup, 0)
    torch.cuda.synchronize(device)
    launches0 = model.launch_count
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(main)
    for st in streams:
        st.wait_event(start)
    issue(steps, warmup)
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
    end.record(main)
    torch.cuda.synchronize(device)
    return start.elapsed_time(end), model.launch_count - launches0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0,
                    help="timed steps (default: 4000 pileup / 1500 full-alignment = ~0.4 s, so the 100 ms clock sampler sees the run)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pileup", choices=["pileup", "fa"])
    ap.add_argument("--streams", type=int, default=12)
    ap.add_argument("--lstm-wg", type=int, default=0, help="epilogue warpgroups per LSTM sub-tile (0 = library default)")
    ap.add_argument("--lstm-tile", type=int, default=64,
                    help="sites per LSTM sub-tile (16|32|64; 0 = library auto = latency-oriented 16 at this batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = 4000 if args.workload == "pileup" else 1500
        if args.impl == "reference":
            args.steps = 20
    if args.warmup < 3:
        args.warmup = 3
    args.warmup = max(args.warmup, args.streams)      # every stream's workspace exists before the timed region

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    workload = args.workload
    model, sd = make_model(workload, device, load_real_weights=(rank == 0))
    if args.lstm_wg and workload == "pileup":
        model.set_option("lstm_wg", args.lstm_wg)
    if args.lstm_tile and workload == "pileup":
        model.set_option("lstm_tile", args.lstm_tile)
        LSTM_TILE[0] = args.lstm_tile
    bcast_bytes = 0
    if world > 1:
        from clair3_b200 import sharding
        bcast_bytes = sharding.broadcast_weights(model, src=0)       # the one NCCL collective, before the timed region
        sd = synth.pileup_state_dict(False, seed=0) if workload == "pileup" else synth.fa_state_dict(True, channels=8, seed=0)

    b = BATCH[workload]
    in_bytes = int(np.prod(make_inputs(workload, 1, 0)[0].shape)) * (4 if workload == "pileup" else 1)
    pool = max(8, int(140e6 // in_bytes) + 1)
    xs_host = make_inputs(workload, pool, seed=100 + 1000 * rank)    # every rank gets its own sites (weak scaling)
    xs_dev = [torch.from_numpy(x).to(device) for x in xs_host]
    n_streams = max(1, args.streams)
    streams = [torch.cuda.Stream(device) for _ in range(n_streams)]
    ys_dev = [None] * (2 * n_streams)

    # ---- parity spot check before timing (rank-local, tiny): the bench never times a wrong kernel
    from oracle import clair3_oracle as orc                           # checker only
    chk = xs_host[0][:8]
    ref = orc.pileup_forward(sd, chk, False) if workload == "pileup" else orc.fa_forward(sd, chk, True)
    got = model(torch.from_numpy(chk).to(device)).cpu().numpy()
    parity = float(np.abs(got - ref).max())
    if not (parity < 2e-2):
        raise SystemExit("parity check failed before timing: max |dp| = %g" % parity)

    # ---- timed region: device-resident inputs
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed_steps(model, xs_dev, ys_dev, streams, args.steps, args.warmup, device)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    value = b * args.steps * world / (ms * 1e-3)

    # ---- e2e: host tensors in and out through the module API, H2D and D2H inside the timed region.
    #  (a) pipelined: Clair3_X.forward_async(pinned x, pinned y) round-robin over the streams, one sync at the end;
    #  (b) synchronous: y = model(x_pinned) per step, exactly the shape of the reference's _torch_predict.
    e2e_steps = max(10, args.steps // 2)
    xs_pin = [torch.from_numpy(x).pin_memory() for x in xs_host[:max(8, n_streams)]]
    ys_pin = [torch.empty((b, model.out_dim), dtype=torch.float32).pin_memory() for _ in range(n_streams)]

    def issue_e2e(n, offset):
        for i in range(n):
            st = streams[(offset + i) % n_streams]
            with torch.cuda.stream(st):
                model.forward_async(xs_pin[(offset + i) % len(xs_pin)], ys_pin[(offset + i) % n_streams])
    issue_e2e(2 * n_streams, 0)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    main_stream = torch.cuda.current_stream(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_stream)
    for st in streams:
        st.wait_event(e0)
    issue_e2e(e2e_steps, 0)
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        main_stream.wait_event(ev)
    e1.record(main_stream)
    torch.cuda.synchronize(device)
    e2e_ms = e0.elapsed_time(e1)
    ref_y = model(xs_pin[(e2e_steps - 1) % len(xs_pin)])              # the last pipelined result must equal a sync call
    assert float((ref_y - ys_pin[(e2e_steps - 1) % n_streams]).abs().max()) < 1e-4
    # (b) synchronous per step
    sync_steps = max(10, args.steps // 8)
    for i in range(3):
        model(xs_pin[i % len(xs_pin)])
    torch.cuda.synchronize(device)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(sync_steps):
        y_host = model(xs_pin[i % len(xs_pin)])
    s1.record()
    torch.cuda.synchronize(device)
    sync_ms = s0.elapsed_time(s1)
    assert y_host.device.type == "cpu"
    if world > 1:
        t = torch.tensor([e2e_ms, sync_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, sync_ms = float(t[0].item()), float(t[1].item())
    e2e_value = b * e2e_steps * world / (e2e_ms * 1e-3)
    e2e_sync_value = b * sync_steps * world / (sync_ms * 1e-3)

    # ---- per-kernel device time (single stream, CUDA events around every launch inside the library)
    pk = peaks()
    roofline = None
    kernels = {}
    if rank == 0:
        from clair3_b200._ffi import check, ffi, lib
        model.set_option("profile", 1)
        prof_steps = min(args.steps, 40)
        for i in range(prof_steps):
            model(xs_dev[i % len(xs_dev)])
        torch.cuda.synchronize(device)
        names = list(KERNEL_FLOP_PER_SITE[workload]) + ["ingest", "heads"] + (["spp"] if workload == "fa" else [])
        tot = 0.0
        for nme in names:
            pms, pn = ffi.new("double *"), ffi.new("int64_t *")
            check(lib().c3b_get_profile(model._handle, nme.encode(), pms, pn))
            if pn[0]:
                kernels[nme] = {"ms_per_launch": pms[0] / pn[0], "launches": int(pn[0])}
                tot += pms[0] / pn[0]
        model.set_option("profile", 0)
        for nme, k in kernels.items():
            k["share"] = k["ms_per_launch"] / tot
            fl = KERNEL_FLOP_PER_SITE[workload].get(nme)
            if fl:
                k["tflops"] = fl * b / (k["ms_per_launch"] * 1e-3) / 1e12
                k["frac_of_bf16_burst"] = k["tflops"] / pk["bf16_burst"]
        dom = max((n for n in kernels if n in KERNEL_FLOP_PER_SITE[workload]), key=lambda n: kernels[n]["ms_per_launch"])
        # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json, written by
        # tools/ncu_summary.py from `ncu --set full`: dram__bytes_read.sum + dram__bytes_write.sum), if it was taken at this
        # workload's batch size
        traffic = None
        try:
            tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")))
            ent = tj.get(workload, {}).get(dom)
            if ent and ent.get("batch") == b:
                traffic = ent["dram_bytes_per_launch"]
        except (OSError, ValueError):
            pass
        roofline = {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": pk["bf16_burst"],
                    "unit": "TFLOP/s", "frac": kernels[dom]["tflops"] / pk["bf16_burst"], "traffic": traffic,
                    "peak_source": pk["which"] + ", burst figure (kernel timed alone between CUDA events)",
                    "flop_per_launch": KERNEL_FLOP_PER_SITE[workload][dom] * b,
                    "whole_step": {"achieved": FLOP_PER_SITE[workload] * value / world / 1e12,
                                   "frac_of_sustained": FLOP_PER_SITE[workload] * value / world / 1e12 / pk["bf16_sustained"]}}
        if dom in ("lstm1", "lstm2"):
            # The recurrent kernels are not tensor-bound: their epilogue needs 5 MUFU.TANH per (site, step, direction, unit)
            # and the SFU pipe issues 16 lanes/clk/SM.  One launch occupies 2 * ceil(B / (2*tile)) CTAs (one per SM), so the
            # honest ceiling for THIS launch is those SMs' SFU rate; the other SMs are filled by the other streams.
            units = 128 if dom == "lstm1" else 160
            tile = int(args.lstm_tile) or 64
            if dom == "lstm2":
                tile = min(tile, 32)        # LSTM2's ten accumulator blocks fit TMEM only up to 32 sites per sub-tile
            ctas = 2 * ((b + 2 * tile - 1) // (2 * tile))
            mufu = 5.0 * 33 * 2 * units * b
            clk_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            per_clk_sm = mufu / (kernels[dom]["ms_per_launch"] * 1e-3 * clk_hz) / ctas
            roofline["limiter"] = {"resource": "SFU (MUFU.TANH) issue, 16 lanes/clk/SM", "mufu_ops_per_launch": mufu, "ctas": ctas,
                                   "achieved_per_clk_per_sm": per_clk_sm, "peak_per_clk_per_sm": 16.0, "frac": per_clk_sm / 16.0}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        v, iters, dt, threads = time_cpu(workload, sd, 12.0, threads)
        cpu = {"value": v, "unit": "sites/s", "cores": threads, "kind": "port",
               "sample": "%d steps of %d sites in %.1f s; torch CPU ops of the reference forward (oracle/torch_port.py), "
                         "fastest of {8,16,32,64,all=%d} threads" % (iters, b, dt, len(os.sched_getaffinity(0)))}
        cpu["single_process"] = {"value": v, "cores": threads}
        dep = cpu_deployment_shape(workload, len(os.sched_getaffinity(0)))
        cpu["deployment_shape"] = dep
        if dep["value"] > v:          # report the stronger CPU configuration as the baseline
            cpu["value"], cpu["cores"] = dep["value"], dep["processes"]
            cpu["sample"] = dep["sample"] + " (the reference's --threads deployment; beats one multi-threaded process)"

    if rank == 0:
        line = {
            "metric": "candidate-sites/sec", "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": workload_config(workload, n_streams, pool),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": b * model.out_dim * 4,
                    "steps": e2e_steps,
                    "mode": "Clair3_%s.forward_async(pinned host x, pinned host y) pipelined over %d streams (H2D + kernels + D2H "
                            "stream-ordered per step, one synchronise at the end)" % ("P" if workload == "pileup" else "F", n_streams),
                    "synchronous_per_step": {"value": e2e_sync_value, "unit": "sites/s", "steps": sync_steps,
                                             "mode": "y = model(x_pinned): H2D, forward, D2H, stream sync every step (the _torch_predict shape)"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "parity_max_abs_dp": parity,
            "weight_broadcast_bytes": bcast_bytes,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
