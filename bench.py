#!/usr/bin/env python
"""Benchmark of the RSRGAN GAN-training hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg4|cfg5|cfgP|cfgR]

metric : GAN train frames/sec (257-d LPS -> 40-d MFCC)
step   : one batch schedule of scripts/train_gan_rnn_placeholder.py:72-101 = 1 D update + 2 G updates
         on one synthetic minibatch (B utterances x T frames per GPU; weak scaling, B per GPU fixed)
value  : frames/sec with the minibatches already resident in HBM
e2e    : same, through GAN_RNN.train_batch() with HOST (pinned) buffers: H2D of the batch and D2H of the
         losses inside the timed region
N > 1  : one process per GPU (torchrun), one NCCL all-reduce of the flat gradient buffer per update;
         timed on the device, max over ranks.

`--impl reference` times the CPU restatement of the reference (oracle/cpu_baseline.py, kind "port":
the reference's TF-1.4 path cannot run here) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]/[2]: 2-layer LSTM-512 G (+ projection 256) + 3-hidden-layer DNN D
    "cfg2": dict(g_type="lstm", d_type="dnn", g_cell=512, g_proj=256, g_layers=2, B=128, T=100,
                 name="gan_rnn_placeholder: 2xLSTMP(512->256) G + discriminator_dnn(1024x4), B=128 x T=100 per GPU"),
    # reference-native sizes of the shipped driver family (models/lstm.py:43-45 + discriminator_lstm)
    "cfgP": dict(g_type="lstm", d_type="lstm", g_cell=760, g_proj=280, g_layers=3, B=8, T=100,
                 name="gan_rnn_placeholder ref-native: lstm G (3xLSTMP 760->280) + discriminator_lstm, B=8 x T=100"),
    # BASELINE.json configs[4]: res_lstm_l 4 x 1024 (P = 257) + discriminator_lstm, T = 200, B = 64 per GPU
    "cfg5": dict(g_type="res_lstm_l", d_type="lstm", g_cell=1024, g_proj=257, g_layers=4, B=64, T=200,
                 name="gan_rnn_placeholder: res_lstm_l G (4xLSTMP 1024->257) + discriminator_lstm, B=64 x T=200 per GPU"),
    # BASELINE.json configs[3]: RCED convolutional generator (splice = 1) + discriminator_dnn, 256 frames per GPU
    "cfg4": dict(g_type="rced", d_type="dnn", g_cell=None, g_proj=None, g_layers=None, B=256, T=1,
                 name="frame GAN: RCED generator (9 x conv1d SAME + FC) + discriminator_dnn(1024x4), 256 frames per GPU"),
    "cfgR": dict(g_type="res_lstm_l", d_type="lstm", g_cell=760, g_proj=257, g_layers=4, B=8, T=100,
                 name="run_gan_rnn_placeholder.sh: res_lstm_l G (4xLSTMP 760->257) + discriminator_lstm, B=8 x T=100"),
}


def flops_per_frame(cfg):
    """7 F_G + 10 F_D algorithmic forward-equivalent flops per input frame per schedule."""
    def lstmp(i, c, p):
        return 2 * (i + p) * 4 * c + 2 * c * p
    if cfg["g_type"] == "rced":
        ch, wd = (1, 12, 16, 20, 24, 32, 24, 20, 16, 12), (13, 11, 9, 7, 7, 7, 9, 11, 13)     # models/rced.py:92-93
        fg = sum(2 * 257 * w * ch[i] * ch[i + 1] for i, w in enumerate(wd)) + 2 * 257 * 12 * 40
    elif cfg["g_type"] == "lstm":
        p, c = cfg["g_proj"], cfg["g_cell"]
        fg = 2 * 257 * p + cfg["g_layers"] * lstmp(p, c, p) + 2 * p * 40
    else:
        fg = cfg["g_layers"] * lstmp(257, cfg["g_cell"], 257) + 2 * 257 * 40
    if cfg["d_type"] == "dnn":
        fd = 2 * 40 * 1024 + 3 * 2 * 1024 * 1024 + 2 * 1024
    else:
        fd = lstmp(40, 256, 40) + lstmp(40, 256, 40) + 2 * 40
    if cfg.get("_parts"):
        return fg, fd
    return 7 * fg + 10 * fd


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop = index, [], threading.Event()

    def run(self):
        try:                                   # in-process NVML: a sample every few ms instead of one nvidia-smi fork
            import pynvml
            pynvml.nvmlInit()
            hd = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(hd, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(hd, pynvml.NVML_CLOCK_SM)
                rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(hd)
                self.rows.append([str(sm), str(mx)] + ["Active" if rs & bits[n] else "Not Active"
                                                       for n in ("hw_slowdown", "hw_thermal_slowdown",
                                                                 "sw_thermal_slowdown", "sw_power_cap")])
                self.stop.wait(0.01)
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


NCU_SUMMARY = {   # C-ABI call -> committed `ncu --set full` summary of its kernel (profiles/, made by scripts in DESIGN.md 6)
    "rsr_gemm": "r2_gemm_full_summary.csv",                        # 90 launches of one cfg-2 schedule, final tree
    "rsr_lstmp_rec_bwd": "r2_recbwd_pair_v2_full_summary.csv",     # CTA-pair kernels (cfg-2: Cp = 512)
    "rsr_lstmp_rec_fwd": "r1_recfwd_full_summary.csv",
    "rsr_lstmp_fused_fwd": "r2_recfwd_pair_full_summary.csv",
    "rsr_lstmp_wave_fwd": "r2_wave_fwd_full_summary.csv",          # layer-wavefront launches (cfg-2 forward, cfg-P backward)
    "rsr_lstmp_wave_bwd": "r2_wave_bwd_full_summary.csv",
}


def ncu_traffic(call):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches) from the
    committed ncu --set full capture of this kernel, or None."""
    import csv
    path = os.path.join(ROOT, "profiles", NCU_SUMMARY.get(call, "-"))
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for r in rows[2:]:
        try:
            b = sum(float(r[hdr.index(k)]) * mult[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        except (ValueError, KeyError):
            try:        # section captures carry the total only (dram__bytes.sum.per_second x duration, see profiles/README.md)
                b = float(r[hdr.index("dram__bytes.sum")]) * mult[units[hdr.index("dram__bytes.sum")]]
            except (ValueError, KeyError):
                continue
        tot, n = tot + b, n + 1
    return tot / n if n else None


def cpu_sample_utterances(cfg, cb):
    """Utterances per CPU schedule: the WHOLE minibatch of the config when one schedule of it takes about four seconds or
    less on this host (probed with one schedule of ~3200 frames), else as many utterances as fit that time, never fewer than
    the ~3200-frame probe.  (cfg-2 on the 16-core GPU box: 128 of 128 utterances, ~2 s per schedule.)"""
    B, T = cfg["B"], cfg["T"]
    Bs0 = max(1, min(B, 3200 // T))
    if Bs0 >= B:
        return B
    v, _, _ = cb.time_schedule(cfg, Bs0, T, steps=1, warmup=1)
    return int(min(B, max(Bs0, v * 4.0 / T)))


def sample_text(Bs, cfg):
    if Bs == cfg["B"]:
        return "the whole minibatch: %d utterances x %d frames per step" % (Bs, cfg["T"])
    return "%d of %d utterances x %d frames per step" % (Bs, cfg["B"], cfg["T"])


def run_reference(a, cfg):
    """CPU arm: oracle/cpu_baseline.py (port of the reference schedule) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    Bs = cpu_sample_utterances(cfg, cb)            # the whole minibatch when the host is fast enough, else a bounded sample
    gan = cb.CpuGan(cfg, 1234)
    import torch
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(Bs, cfg["T"], 257, generator=g)
    y = torch.randn(Bs, cfg["T"], 40, generator=g)
    ln = torch.full((Bs,), cfg["T"], dtype=torch.int64)
    for _ in range(a.warmup):
        gan.schedule(x, y, ln)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        gan.schedule(x, y, ln)
    dt = (time.perf_counter() - t0) / max(a.steps, 1)
    v = Bs * cfg["T"] / dt
    sample = sample_text(Bs, cfg) + " (same networks, same schedule), torch-CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": "gan_train_frames_per_sec", "value": v, "unit": "frames/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "sample": sample, "same_config": bool(Bs == cfg["B"])},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--dtype", default="f16", help="16-bit tensor-core operand type: f16 | bf16 (fp32 accumulate/state)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    if a.impl == "reference":
        return run_reference(a, cfg)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: while the benchmark runs, file descriptor 1 points at stderr, so that
    # banners written by native libraries (the "NCCL version ..." line appears on stdout at communicator creation)
    # cannot precede it; the descriptor is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rsrgan_b200.gan_rnn import GAN_RNN

    B, T = cfg["B"], cfg["T"]
    args = Namespace(g_type=cfg["g_type"], d_type=cfg["d_type"], batch_size=B, num_gpu=world,
                     **{k: cfg[k] for k in ("g_cell", "g_proj", "g_layers") if cfg[k] is not None},
                     init_mse_weight=10.0, init_disc_noise_std=0.05, l2_scale=0.0, dtype=a.dtype, seed=1234,
                     # run_gan_rnn_placeholder.sh:127-128, times num_gpu (train...py:458-459)
                     g_learning_rate=8e-5 * world, d_learning_rate=1e-3 * world)
    model = GAN_RNN(None, args, ["/gpu:%d" % local])
    h = model.h

    # synthetic minibatches: x, y ~ N(0,1), full lengths; a ring of distinct batches
    rng = np.random.default_rng(1234 + rank)
    NB = 4
    host = [(rng.standard_normal((B, T, 257), dtype=np.float32), rng.standard_normal((B, T, 40), dtype=np.float32),
             np.full(B, T, np.float32)) for _ in range(NB)]
    pinned = [tuple(torch.from_numpy(v).pin_memory() for v in b) for b in host]
    resident = [tuple(v.cuda(non_blocking=True) for v in b) for b in pinned]
    resident = [(x, y, ln.to(torch.int32)) for x, y, ln in resident]
    torch.cuda.synchronize()

    def step_resident(i):
        x, y, ln = resident[i % NB]
        model.train_batch(x, y, ln, sync=False)

    def step_e2e(i):
        # every step: H2D of that step's batch from pinned host memory and D2H of its losses.  The copy of batch
        # i + 1 is started (GAN_RNN.prefetch, a copy stream) before the schedule of batch i is waited for -- the
        # double buffering any input pipeline does; the first batch of the loop is copied in line.
        x, y, ln = pinned[i % NB]
        nx = pinned[(i + 1) % NB]
        if getattr(step_e2e, "primed", None) != i:
            model.prefetch(x, y, ln)
        out = model.train_batch(x, y, ln, sync=False)
        model.prefetch(*nx)
        step_e2e.primed = i + 1
        d_vals, g = out
        return d_vals.tolist(), g.tolist()                   # D2H read of the step's losses (synchronises)

    def timed(fn, steps, warmup, sampler=None):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = h.launches
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.stop.set()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), h.launches - l0

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = timed(step_resident, a.steps, max(a.warmup, 3), sampler)
    value = B * T * world * a.steps / (ms * 1e-3)
    ms_e2e, _ = timed(step_e2e, a.steps, 3)
    e2e = B * T * world * a.steps / (ms_e2e * 1e-3)

    # kernel shares: a few eager steps with every C-ABI call bracketed by CUDA events on its stream
    h.timing = []
    nprof = 3
    for i in range(nprof):
        step_resident(i)
    shares = h.timing_summary()
    h.timing = None
    tot = sum(v[1] for v in shares.values()) or 1.0
    top = max(shares.items(), key=lambda kv: kv[1][1])
    sus, burst, hbm, how = peaks()
    name, (cnt, tms, work) = top
    def roof(name, cnt, tms, work):
        ach = work / (tms * 1e-3) / 1e12 if tms > 0 else 0.0
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": sus, "unit": "TFLOP/s",
                "frac": ach / sus, "peak_burst": burst, "frac_burst": ach / burst, "traffic": ncu_traffic(name),
                "traffic_source": "profiles/" + NCU_SUMMARY[name] + " (ncu capture of one schedule, mean over the captured launches)"
                if name in NCU_SUMMARY else None,
                "peak_source": how + " (bf16 sustained = torch.matmul back to back for 4 s; peak_burst / frac_burst = its best single call, the "
                               "figure that matches the boost clocks of this short timed region; fp16 runs at the same rate)",
                "launches_per_step": cnt / nprof, "avg_launch_ms": tms / cnt, "share_of_step": tms / tot,
                "algorithmic_flops_per_launch": work / cnt}
    roofline = roof(name, cnt, tms, work)
    # the fused LSTM-gate kernels named by north_star, reported whatever their share (latency / DSMEM bound)
    lstm_roofs = [roof(k, *shares[k]) for k in ("rsr_lstmp_wave_fwd", "rsr_lstmp_fused_fwd", "rsr_lstmp_rec_fwd", "rsr_lstmp_rec_bwd")
                  if k in shares]
    fpf = flops_per_frame(cfg)
    step_tflops = value / world * fpf / 1e12
    # what the schedule EXECUTES: G(x) of the D update is the forward the first G update reuses when the generator has no
    # dropout (same weights, same minibatch, same numbers: GAN_RNN._schedule), i.e. 6 F_G instead of the reference's 7
    fg1, fd1 = flops_per_frame(dict(cfg, _parts=True))
    shared = getattr(model, "G", None) is not None and model.G.keep_prob >= 1.0 and getattr(model, "gen_updates", 0) > 0
    fpf_exec = (6 if shared else 7) * fg1 + 10 * fd1

    out = {
        "metric": "gan_train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": cfg["name"], "schedule": "1 D update + 2 G updates per minibatch",
                   "frames_per_step_per_gpu": B * T, "parallelism": "dp%d" % world,
                   "l2": "working set per step (saved gate activations, %.0f MB) exceeds the 126 MB L2; 4 distinct minibatches rotate"
                         % (model.G.ws.nbytes() / 1e6)},
        "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in pinned[0])),
                "d2h_bytes_per_step": 2 * 8 * 4},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_lstm_kernels": lstm_roofs,
        "step_roofline": {"algorithmic_mflop_per_frame": fpf / 1e6, "achieved_tflops_per_gpu": step_tflops,
                          "frac_of_sustained_bf16": step_tflops / sus, "frac_of_burst_bf16": step_tflops / burst,
                          "executed_mflop_per_frame": fpf_exec / 1e6,
                          "executed_tflops_per_gpu": value / world * fpf_exec / 1e12,
                          "note": "algorithmic = 7 F_G + 10 F_D of the reference schedule (SURVEY 8d); executed = with the generator "
                                  "forward shared between the D update and the first G update (identical numbers)"},
        "kernel_shares": {k: {"calls_per_step": v[0] / nprof, "ms_per_step": v[1] / nprof, "share": v[1] / tot}
                          for k, v in sorted(shares.items(), key=lambda kv: -kv[1][1])},
    }
    if world > 1:
        # data-parallel invariant: every rank applied the same averaged gradients -> bit-identical weights
        chk = torch.stack([model.G.P.theta.double().sum(), model.D.P.theta.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["ranks_in_sync"] = bool(torch.equal(lo, hi))
        # gradient averaging: one kernel over NVLink peer memory inside the schedule's CUDA graph, or NCCL between segments
        out["allreduce"] = "peer" if model.peer is not None else "nccl"
        if model.peer is not None:
            out["allreduce_error"] = int(model.peer.error())
    if rank == 0:
        out["clocks"] = sampler.summary()
        if world == 1 and not a.no_cpu_baseline:
            from oracle import cpu_baseline as cb
            Bs = cpu_sample_utterances(cfg, cb)
            v, dt, cores = cb.time_schedule(cfg, Bs, T, steps=2, warmup=0)
            out["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": sample_text(Bs, cfg) + ", 2 schedules (median), torch-CPU fp32 restatement "
                                                                    "of the reference (TF-1.4 unavailable)"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)                                        # anything printed during teardown goes to stderr again
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
