"""GPU diagnostics: per-case, per-tap error report for both precisions (prints, never asserts).

    python tools/diag.py fp32|tc [case ...]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import GOLDEN_FA, GOLDEN_PILEUP, golden_case  # noqa: E402


def relerr(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def run_case(name, precision, opts):
    from clair3_b200.model import Clair3_F, Clair3_P
    z, meta, sd, x = golden_case(name)
    cls = Clair3_P if meta["kind"] == "pileup" else Clair3_F
    ch = 18 if meta["kind"] == "pileup" else meta["channels"]
    m = cls(add_indel_length=meta["add_indel_length"], predict=True, input_channels=ch)
    m.set_option("precision", precision)
    for k, v in opts.items():
        m.set_option(k, v)
    m.to(torch.device("cuda"))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    t0 = time.time()
    y = m(torch.from_numpy(x).cuda()).cpu().numpy()
    dt = time.time() - t0
    d = np.abs(y - z["y"])
    print(f"[{name}] prec={precision} opts={opts} out max|dp|={d.max():.3e} mean|dp|={d.mean():.3e} "
          f"finite={np.isfinite(y).all()} ({dt*1e3:.1f} ms first call)", flush=True)
    for tap in ("lstm1", "lstm2", "conv1", "res_block1", "conv3", "res_block2", "conv5", "res_block3", "spp", "l4_pre"):
        key = "tap_" + tap
        if key not in z.files:
            continue
        want = z[key]
        try:
            got = m.tap(tap)
        except Exception as e:  # noqa: BLE001
            print(f"    tap {tap}: unavailable ({e})")
            continue
        if tap in ("conv1", "res_block1", "conv3", "res_block2", "conv5", "res_block3"):
            got = got.reshape(x.shape[0], want.shape[2], want.shape[3], want.shape[1])[:1].transpose(0, 3, 1, 2)
        else:
            got = got.reshape(x.shape[0], -1)
            if tap == "l4_pre":
                got = got + sd["L4.bias"][None, :]
            got = got[:want.shape[0]].reshape(want.shape)
        print(f"    tap {tap:11s} rel-L2 {relerr(got, want):.3e}  max|d| {np.abs(got - want).max():.3e}  "
              f"(|ref| max {np.abs(want).max():.2f})", flush=True)


def run_ptrace(conv=1):
    """Cycle stamps of CTA 0 of one Clair3_F conv (pconv): MMA warp and epilogue per macro-tile."""
    from clair3_b200 import synth
    from clair3_b200._ffi import check, ffi, lib
    from clair3_b200.model import Clair3_F
    sd = synth.fa_state_dict(True, channels=8, seed=0)
    x = synth.fa_inputs(256, seed=0)
    m = Clair3_F(add_indel_length=True, predict=True, input_channels=8)
    m.set_option("lstm_trace", 10 + conv)
    m.to(torch.device("cuda"))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    xd = torch.from_numpy(x).cuda()
    for _ in range(3):
        m(xd)
    print(f"--- conv{conv}")
    buf = np.zeros(2 * 33 * 4, dtype=np.int64)
    check(lib().c3b_debug_lstm_trace(m._handle, ffi.cast("int64_t *", buf.ctypes.data)))
    tr = buf[:64].reshape(8, 8)
    t0 = tr[0, 0]
    for li in range(8):
        r = tr[li]
        if r[0] == 0:
            continue
        if li == 0 and buf[64]:
            x = buf[64:72]
            print(f"  ring: piece0 issued+commit queued {x[0]-t0}; loader: first refill released {x[1]-t0}; MMA thread: starts waiting for it "
                  f"{x[5]-t0}, own half landed {x[3]-t0}, peer half forwarded {x[4]-t0}", flush=True)
        print(f"macro {li}: MMA warp: start {r[0]-t0:7d} tmem_wait {r[2]-r[0]:6d} first_block_wait {r[1]-r[2]:6d} mma_issue+run {r[3]-r[1]:6d} | "
              f"epilogue: start {r[4]-t0:7d} wait_full {r[5]-r[4]:6d} work {r[6]-r[5]:6d}", flush=True)


def run_stress(opts):
    """Concurrent-stream consistency: every pipelined forward must equal the single-stream result."""
    from clair3_b200 import synth
    from clair3_b200.model import Clair3_P
    sd = synth.pileup_state_dict(False, seed=0)
    xs = [synth.pileup_inputs(1024, seed=100 + i) for i in range(8)]
    m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
    for k, v in opts.items():
        m.set_option(k, v)
    m.to(torch.device("cuda"))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    xd = [torch.from_numpy(x).cuda() for x in xs]
    ref = [m(x).cpu().numpy() for x in xd]
    again = [m(x).cpu().numpy() for x in xd]
    print("single-stream repeatability max diff", max(float(np.abs(a - b).max()) for a, b in zip(ref, again)))
    reftaps = {}
    m.set_option("tap_ws", 0)
    for j in range(8):
        m(xd[j])
        reftaps[j] = {t: m.tap(t).copy() for t in ("lstm1", "lstm2", "l4_pre")} if opts.get("precision", 0) == 0 or True else {}
    streams = [torch.cuda.Stream() for _ in range(8)]
    worst = 0.0
    for rep in range(4):
        outs = [None] * 8
        for i in range(8):
            with torch.cuda.stream(streams[i]):
                outs[i] = m(xd[i])
        torch.cuda.synchronize()
        for i in range(8):
            d = np.abs(outs[i].cpu().numpy() - ref[i])
            worst = max(worst, float(d.max()))
            if d.max() > 1e-4:
                rows = np.where(d.max(1) > 1e-4)[0]
                msg = f"rep {rep} stream {i}: out max diff {d.max():.3e} in {len(rows)} rows {rows[:10]}"
                m.set_option("tap_ws", i + 1)          # workspace 0 belongs to the default stream
                for t in ("lstm1", "lstm2", "l4_pre"):
                    try:
                        tv = m.tap(t)
                        dd = np.abs(tv - reftaps[i][t]).reshape(1024, -1).max(1)
                        bad = np.where(dd > 1e-3)[0]
                        msg += f" | {t}: {len(bad)} bad rows {bad[:6]}"
                    except Exception as e:  # noqa: BLE001
                        msg += f" | {t}: {e}"
                print(msg, flush=True)
    print("multi-stream worst diff", worst, "opts", opts, flush=True)


def run_probe():
    from clair3_b200._ffi import check, ffi, lib
    r = np.random.default_rng(0)
    for n in (16, 32, 64, 128, 256):
        a = r.standard_normal((128, 16)).astype(np.float32)
        b = r.standard_normal((n, 16)).astype(np.float32)
        d = np.zeros((128, n), dtype=np.float32)
        tm = np.zeros(10, dtype=np.int64)
        reps = 64
        check(lib().c3b_debug_ts_probe(ffi.cast("float *", a.ctypes.data), ffi.cast("float *", b.ctypes.data), n, reps,
                                       ffi.cast("float *", d.ctypes.data), ffi.cast("int64_t *", tm.ctypes.data)))
        ref = a.astype(np.float16).astype(np.float64) @ b.astype(np.float16).astype(np.float64).T
        print(f"probe N={n}: TS-form max err {np.abs(d - ref).max():.3e} (|ref| max {np.abs(ref).max():.2f});  per MMA: "
              f"SS {tm[1]/reps:.1f} | TS {tm[3]/reps:.1f} | SS 4-acc round-robin {tm[5]/reps:.1f} | SS 2-acc {tm[7]/reps:.1f} | "
              f"TS 4-acc {tm[9]/reps:.1f} cycles (issue-only: {tm[0]/reps:.1f}/{tm[2]/reps:.1f}/{tm[4]/reps:.1f}/{tm[6]/reps:.1f}/{tm[8]/reps:.1f})", flush=True)


def run_mmaprobe():
    """Cycles per tcgen05.mma when the A / B tile and the accumulator change from one MMA to the next."""
    from clair3_b200._ffi import check, ffi, lib
    all_modes = [
        ("same A, same B, same D", (0, 0, 0, 1, 1)),
        ("A rotates", (8192, 0, 0, 1, 1)),
        ("B rotates", (0, 0, 8192, 1, 1)),
        ("A+B rotate, 1 acc (MT=1)", (8192, 0, 8192, 1, 1)),
        ("A+B rotate, 2 acc alternate", (8192, 0, 8192, 1, 2)),
        ("A rotates misaligned 16B", (8192, 16, 0, 1, 1)),
        ("MT=1 misaligned 48B", (8192, 48, 8192, 1, 1)),
        ("MT=2 (B per 2, 2 acc)", (8192, 0, 8192, 2, 2)),
        ("MT=2 misaligned", (8192, 16, 8192, 2, 2)),
        ("MT=4 (B per 4, 4 acc)", (8192, 0, 8192, 4, 4)),
        ("MT=4 misaligned", (8192, 16, 8192, 4, 4)),
        ("MT=8 (B per 8, 8 acc)", (8192, 0, 8192, 8, 8)),
        ("same A, B rotates, 4 acc", (0, 0, 8192, 1, 4)),
        ("same A+B, 4 acc", (0, 0, 0, 1, 4)),
    ]
    reps = 256
    for n in (64, 128, 256):
        modes = [(nm, md) for nm, md in all_modes if md[4] * n <= 512]
        arr = np.array([md for _, md in modes], dtype=np.int32)
        tm = np.zeros(2 * len(modes), dtype=np.int64)
        check(lib().c3b_debug_mma_probe(n, reps, len(modes), ffi.cast("int *", arr.ctypes.data), ffi.cast("int64_t *", tm.ctypes.data)))
        for i, (nm, _) in enumerate(modes):
            print(f"N={n:3d} {nm:32s}: {tm[2*i+1]/reps:6.1f} cycles/MMA (issue {tm[2*i]/reps:5.1f})", flush=True)


def run_itrace():
    """Cycle stamps of CTA 0 of the LSTM2 input projection (igemm): MMA thread and one epilogue thread per tile."""
    from clair3_b200 import synth
    from clair3_b200._ffi import check, ffi, lib
    from clair3_b200.model import Clair3_P
    sd = synth.pileup_state_dict(False, seed=0)
    x = synth.pileup_inputs(1024, seed=0)
    m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
    m.set_option("lstm_tile", 64)
    m.set_option("lstm_trace", 30)
    m.to(torch.device("cuda"))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    xd = torch.from_numpy(x).cuda()
    for _ in range(3):
        m(xd)
    buf = np.zeros(2 * 33 * 4, dtype=np.int64)
    check(lib().c3b_debug_lstm_trace(m._handle, ffi.cast("int64_t *", buf.ctypes.data)))
    tr = buf[:64].reshape(8, 8)
    t0 = tr[0, 0]
    for li in range(8):
        r = tr[li]
        if r[0] == 0:
            continue
        print(f"tile {li}: MMA: start {r[0]-t0:7d} tmem_wait {r[1]-r[0]:6d} first_stage_wait {r[2]-r[1]:6d} issue+run {r[3]-r[2]:6d} | "
              f"epilogue: start {r[4]-t0:7d} wait_full {r[5]-r[4]:6d} work {r[6]-r[5]:6d}", flush=True)


def run_tmemprobe():
    from clair3_b200._ffi import check, ffi, lib
    reps = 256
    tm = np.zeros(6, dtype=np.int64)
    check(lib().c3b_debug_tmem_probe(reps, ffi.cast("int64_t *", tm.ctypes.data)))
    for ph, nm in enumerate(("tensor pipe idle", "MMA stream running")):
        print(f"{nm:20s}: ld16+wait {tm[ph*3]/reps:6.1f} | 4 x ld16 per wait {tm[ph*3+1]/reps:6.1f} per ld | "
              f"full epilogue chunk {tm[ph*3+2]/reps:6.1f} cycles", flush=True)


def run_trace(opts):
    """Per-step cycle breakdown of the persistent LSTM kernels (CTA 0, thread 0)."""
    from clair3_b200 import synth
    from clair3_b200._ffi import check, ffi, lib
    from clair3_b200.model import Clair3_P
    sd = synth.pileup_state_dict(False, seed=0)
    x = synth.pileup_inputs(1024, seed=0)
    for tile in ([opts["lstm_tile"]] if "lstm_tile" in opts else [16, 32, 64]):
        m = Clair3_P(add_indel_length=False, predict=True, input_channels=18)
        m.set_option("lstm_tile", tile)
        m.set_option("lstm_trace", 1)
        if "lstm_mufu16" in opts:
            m.set_option("lstm_mufu16", opts["lstm_mufu16"])
        m.to(torch.device("cuda"))
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
        xd = torch.from_numpy(x).cuda()
        for _ in range(3):
            m(xd)
        buf = np.zeros(2 * 33 * 4, dtype=np.int64)
        check(lib().c3b_debug_lstm_trace(m._handle, ffi.cast("int64_t *", buf.ctypes.data)))
        tr = buf.reshape(2, 33, 4)
        for layer in range(2):
            t = tr[layer]
            issue = (t[:, 1] - t[:, 0])[2:].mean()
            mma_wait = (t[:, 2] - t[:, 1])[2:].mean()
            epi = (t[:, 3] - t[:, 2])[2:].mean()
            gap = (t[1:, 0] - t[:-1, 3])[2:].mean()
            total = (t[-1, 3] - t[0, 0]) / 33.0
            print(f"tile {tile} LSTM{layer+1}: per step {total:.0f} cyc = issue {issue:.0f} + wait-for-accumulator {mma_wait:.0f} "
                  f"+ epilogue {epi:.0f} + stage/sync gap {gap:.0f}", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    cases = sys.argv[2:] or (GOLDEN_PILEUP + GOLDEN_FA)
    opts = {}
    cases2 = []
    for c in cases:
        if "=" in c:
            k, v = c.split("=")
            opts[k] = int(v)
        else:
            cases2.append(c)
    cases = cases2 or (GOLDEN_PILEUP + GOLDEN_FA)
    if mode == "ptrace":
        for c in ([int(d) for d in str(opts["convs"])] if "convs" in opts else [1]):
            run_ptrace(c)
    elif mode == "stress":
        run_stress(opts)
    elif mode == "probe":
        run_probe()
    elif mode == "mmaprobe":
        run_mmaprobe()
    elif mode == "itrace":
        run_itrace()
    elif mode == "tmemprobe":
        run_tmemprobe()
    elif mode == "trace":
        run_trace(opts)
    else:
        for c in cases:
            try:
                run_case(c, 1 if mode == "fp32" else 0, opts)
            except Exception as e:  # noqa: BLE001
                print(f"[{c}] ERROR: {e}", flush=True)
