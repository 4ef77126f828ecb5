#!/usr/bin/env python
"""Benchmark of the hot path: es_en_20h-shaped training steps (fwd + bwd + optimizer) over bucketed,
Fisher-shaped synthetic batches -> train input frames / second (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU
                                                             # (numpy oracle; real Chainer is not installable)
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for what each field means.
"""
import argparse
import json
import os
import sys as _sys

# The CPU arms (`--impl reference`, `cpu_baseline`) use ALL host cores: torchrun exports OMP_NUM_THREADS=1, which silently
# single-threads OpenBLAS (round-1 N>1 reference numbers were 2x too slow for that reason).  Must happen before numpy loads.
if "reference" in _sys.argv:
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)
import random
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

D, V, BATCH = 40, 1098, 32
TRAIN_EXTRAS = {"random_out": 0, "speech_noise": 0.25, "teach_ratio": 0.8}
DATA_CFG = {"buckets_num": 20, "buckets_width": 80, "train_scale": 1, "max_pred": 175, "zero_input": 0.1,
            "dec_key": "bpe_w", "enc_key": "sp"}
OPT_CFG = {"type": 0, "lr": 0.001, "l2": 0.0001, "grad_clip": 2, "grad_noise_eta": 0, "freeze": []}
DROPOUT = (0.3, 0.3, 0.0)          # experiments/es_en_20h/model_cfg.json
SET = "fisher_train"


def model_cfg():
    from ast_b200.config import es_en_20h_model_cfg
    return es_en_20h_model_cfg(vocab=V, dropout=DROPOUT)


def make_plan(world, n_batches, seed=1234):
    """Global batch plan of the synthetic epoch (same on every rank and for both impls)."""
    from ast_b200.dataloader import SyntheticDataLoader, plan_batches
    random.seed("seed-ast-20h")                          # train_cfg.json seed, nn.py:54
    np.random.seed(seed)
    loader = SyntheticDataLoader.fisher_shaped(DATA_CFG, None, 0, D, V, n_utts=17306 * max(1, world), seed=seed, set_key=SET)
    plan = plan_batches(loader.buckets[SET], BATCH * world)
    plan = [p for p in plan if len(p[0]) == BATCH * world][:n_batches]
    return loader, plan


def host_batch(loader, utts):
    """Host-side batch exactly as the loader would build it (features, keep masks, labels)."""
    from ast_b200.dataloader import drop_frame_mask
    max_sp = (DATA_CFG["buckets_num"] + 1) * DATA_CFG["buckets_width"]
    feats = [loader._load_utt(u, SET)[:max_sp] for u in utts]
    keep = [drop_frame_mask(len(f), DATA_CFG["zero_input"]) for f in feats]
    ys = [np.asarray([1] + loader._labels(u, SET)[:DATA_CFG["max_pred"] - 2] + [2], dtype=np.int32) for u in utts]
    L = max(len(v) for v in ys)
    y = np.zeros((len(ys), L), dtype=np.int32)
    for i, v in enumerate(ys):
        y[i, :len(v)] = v
    frames = int(sum(len(f) for f in feats))
    return feats, keep, y, frames


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        mx = [float(r[2]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def blas_threads_all_cores():
    """Context manager: OpenBLAS / OpenMP pools at os.cpu_count() for the CPU legs (whatever the launcher exported)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def blas_thread_count():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return None


def step_flops(B, T, L, feat_dim=None, vocab=None):
    """Algorithmic FLOPs of one training step from SURVEY 8(d)'s forward-MAC formulas (FLOPs = 2 MACs, fwd+bwd = 3 x fwd),
    split by stage.  S = L - 1 decode steps actually run (zip(y, y[1:]), seq2seq.py:423)."""
    Dd = D if feat_dim is None else feat_dim
    Vv = V if vocab is None else vocab
    T1 = (T - 1) // 2 + 1
    Tp = (T1 - 1) // 2 + 1
    Fp = (Dd - 13) // 13 + 1
    R, h, H, A, E, S = 512 * Fp, 256, 512, 512, 128, L - 1
    cnn = B * 128 * T1 * Fp * 117 + B * 512 * Tp * Fp * 1152
    enc = 2 * B * Tp * 4 * h * (R + h + h) + 2 * B * Tp * 3 * 4 * h * h
    dec = B * S * 4 * H * (E + A + H + H) + B * S * 3 * 4 * H * H + B * S * (H * H + 2 * Tp * H) + B * S * 2 * H * A + B * S * A * Vv
    return {"cnn": 2.0 * cnn, "enc": 2.0 * enc, "dec": 2.0 * dec, "fwd": 2.0 * (cnn + enc + dec), "step": 6.0 * (cnn + enc + dec), "Tp": Tp}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j["bf16_tflops"], j.get("bf16_tflops_sustained", j["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (numpy oracle, OpenBLAS on all cores)
# ------------------------------------------------------------------------------------------------
def cpu_step_rate(loader, plan, n_steps, budget_s=25.0):
    """frames/s of fwd+bwd+update for up to n_steps batches (bounded by budget_s of CPU work)."""
    from oracle import ast_oracle as O
    cfg = model_cfg()
    cfg["dropout"] = {"embed": 0.0, "rnn": 0.0, "out": 0.0}     # mask multiplies are a rounding error of the CPU time
    P = O.init_params(cfg, D, seed=0)
    om = O.OracleModel(cfg, P, dtype=np.float32)
    opt = O.OracleAMSGrad(om.p, lr=OPT_CFG["lr"], l2=OPT_CFG["l2"], grad_clip=OPT_CFG["grad_clip"])
    frames, t_total, done = 0, 0.0, 0
    desc = []
    for utts, _ in plan[:n_steps]:
        feats, keep, y, fr = host_batch(loader, utts[:BATCH])
        X = O.pad_sequence([f * k[:, None] for f, k in zip(feats, keep)], 0)
        bits = O.teacher_forcing_bits(y.shape[1], TRAIN_EXTRAS["teach_ratio"])
        t0 = time.perf_counter()
        om.forward_loss(X, y, tf_bits=bits)
        g = om.backward()
        opt.update(om.p, g)
        t_total += time.perf_counter() - t0
        frames += fr
        done += 1
        desc.append(f"B{X.shape[0]}xT{X.shape[1]}xL{y.shape[1]}")
        if t_total > budget_s:
            break
    return frames / t_total, done, t_total, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    loader, plan = make_plan(1, max(args.steps + args.warmup, 2))
    with blas_threads_all_cores():
        rate, done, t_total, desc = cpu_step_rate(loader, plan[args.warmup:], args.steps, budget_s=150.0)
    line = {"impl": "reference", "metric": "train_frames_per_sec", "value": rate, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": done, "warmup": 0, "ms_per_step": 1e3 * t_total / max(done, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "es_en_20h bucketed training steps (B=32, Fisher-shaped lengths, D=40, V=1098): fwd+bwd+AMSGrad",
                       "steps_run": desc},
            "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{done} training steps of the same batch plan; numpy restatement of the reference "
                                       "(real Chainer/CuPy is not installable), OpenBLAS on all host cores",
                             "blas_threads": blas_thread_count()},
            "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def roofline_dominant(engine, torch, peaks):
    """Time the largest single-pass tensor-core launch of the step in isolation at its workload shape with CUDA events and
    report achieved / measured peak: the grouped 2-CTA tcgen05 GEMM that computes the two layer-0 input projections of the
    encoder (2 problems of M = T'B = 5120 rows of a B32 x T640 batch, N = 4h = 1024, K = 1536) - the shape of the ncu
    --set full capture under profiles/.  `other_kernels` carries the rest of the GEMM family at the step's other large shapes."""
    import ctypes as C
    from ast_b200._lib import ptr, check
    hbm, tf_burst, tf_sus, how = peaks
    lib = engine.lib
    dev = engine.device
    peak = tf_burst / 2.0                                         # TF32 dense = 1/2 of the measured bf16 figure
    tc = bool(engine.get_option("tc_gemm")) and not engine.get_option("exact")
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, n=8, skip=3):
        ts = []
        for it in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= skip:
                ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))

    if not tc:
        M, N, K = 5120, 1024, 1536
        A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); Cc = torch.empty(M, N, device=dev)
        ms = timed(lambda: check(lib.ast_gemm(0, 0, 1, M, N, K, 1.0, ptr(A), K, ptr(W), K, 0.0, ptr(Cc), N, None, st), "ast_gemm"))
        ach = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        return {"bound": "tensor", "kernel": "sgemm_kernel (fp32 SIMT)", "shape": f"M{M} N{N} K{K}", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "ms_per_launch": ms, "algorithmic_bytes": 4.0 * (M * K + K * N + M * N), "traffic": None}
    G, M, N, K = 2, 5120, 1024, 1536
    A = torch.randn(G, M, K, device=dev); W = torch.randn(G, N, K, device=dev); Cc = torch.empty(G, M, N, device=dev)
    ms = timed(lambda: check(lib.ast_gemm_grouped(G, 0, 1, M, N, K, ptr(A), M * K, K, ptr(W), N * K, K, ptr(Cc), M * N, N, 0, st), "ast_gemm_grouped"))
    flops = 2.0 * G * M * N * K
    achieved = flops / (ms * 1e-3) / 1e12
    out = {"bound": "tensor", "kernel": "gemm_tc2_kernel<NT> (tcgen05 TF32, cta_group::2, grouped launch of 2 problems)",
           "shape": f"2 x (M{M} N{N} K{K}): the two encoder layer-0 input projections of a B32 x T640 batch", "achieved": achieved, "peak": peak,
           "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": f"{how} bf16 burst / 2 (TF32 dense is half of bf16)", "ms_per_launch": ms,
           "algorithmic_bytes": 4.0 * G * (M * K + K * N + M * N), "traffic": None, "traffic_unit": "bytes/launch"}
    out.update(NCU_GEMM)
    others = []
    for (kind, ta, tb, M2, N2, K2, what) in (("2cta", 0, 0, 5120, 1536, 1024, "encoder layer-0 data gradient (one direction)"),
                                             ("2cta_splitk", 1, 0, 1024, 1536, 5120, "encoder layer-0 upward weight gradient (split-K)"),
                                             ("1cta", 0, 0, 15744, 128, 2560, "CNN_1 data gradient, even rows (transposed convolution, overlapping-rows operand)"),
                                             ("1cta", 0, 0, 15744, 1152, 512, "CNN_1 data gradient in its former im2col-gradient form"),
                                             ("2cta", 0, 1, 8192, 4096, 4096, "kernel ceiling: a GEMM large enough to fill the pipeline")):
        overlap = "overlapping-rows" in what        # row j of the operand = rows j .. j+4 of a (M2 + 8) x 512 buffer: lda = 512 < K
        A2 = torch.randn(M2 + 8, 512, device=dev) if overlap else torch.randn((K2, M2) if ta else (M2, K2), device=dev)
        W2 = torch.randn((N2, K2) if tb else (K2, N2), device=dev)
        C2 = torch.empty(M2, N2, device=dev)
        which = {"2cta": -2, "2cta_splitk": -3, "1cta": 1}[kind]
        ms2 = timed(lambda: check(lib.ast_gemm(which, ta, tb, M2, N2, K2, 1.0, ptr(A2), A2.shape[1], ptr(W2), W2.shape[1], 0.0, ptr(C2), N2, None, st),
                                  "ast_gemm"), n=7)
        ach = 2.0 * M2 * N2 * K2 / (ms2 * 1e-3) / 1e12
        others.append({"kernel": {"2cta": "gemm_tc2_kernel (cta_group::2)", "2cta_splitk": "gemm_tc2_kernel (cta_group::2, split-K)",
                                  "1cta": "gemm_tc_kernel (1-CTA persistent)"}[kind],
                       "shape": f"M{M2} N{N2} K{K2} ({'T' if ta else 'N'}{'T' if tb else 'N'}): {what}",
                       "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "ms_per_launch": ms2})
    out["other_kernels"] = others
    out["other_kernels_ncu"] = NCU_GEMM2
    return out


# dram__bytes_read.sum + dram__bytes_write.sum and pipe counters of this kernel at this shape from the ncu --set full capture
# (tools/ncu_capture_roofline.sh, extract under profiles/); filled in from the round's capture
NCU_GEMM = {"traffic": 76.01e6 + 9.36e6,     # the 42 MB output is only partly written back within the kernel's lifetime (126 MB L2)
            "ncu": {"gpu_time_us_cold": 63.4, "tensor_pipe_active_pct": 68.6, "tensor_pipe_elapsed_pct": 49.9,
                    "l2_sector_pct_of_peak": 29.5, "l2_hit_rate_pct": 68.4, "smem_fill_bytes": 503.7e6,
                    "source": "profiles/r01_ncu_full_extract_v28_roofline.txt"}}


NCU_GEMM2 = {"M5120_N1536_K1024": {"gpu_time_us_cold": 36.3, "tensor_pipe_active_pct": 54.9, "tensor_pipe_elapsed_pct": 41.8, "dram_bytes": 27.3e6 + 0.5e6},
             "M8192_N4096_K4096": {"gpu_time_us_cold": 346.2, "tensor_pipe_active_pct": 95.2, "tensor_pipe_elapsed_pct": 88.8, "dram_bytes": 623.3e6 + 117.7e6},
             "M15744_N128_K2560_1cta": {"gpu_time_us_cold": 32.4, "tensor_pipe_active_pct": 42.7, "tensor_pipe_elapsed_pct": 30.1, "dram_bytes": 33.6e6 + 0.01e6},
             "M15744_N1152_K512_1cta": {"gpu_time_us_cold": 56.5, "tensor_pipe_active_pct": 37.6, "tensor_pipe_elapsed_pct": 31.3, "dram_bytes": 34.64e6 + 18.55e6},
             "source": "profiles/r01_ncu_full_extract_v26_gemm.txt, profiles/r01_ncu_full_extract_v28_roofline.txt"}


BEAM_BYTES_PER_STEP = 33e6      # decoder weights 31.6 MB fp32 + enc_states once ("achieved GB/s over steps x 33 MB")


def beam_roofline(search_steps, seconds):
    """Weight-streaming view of beam decoding: every search step reads the decoder weights once; the encoder pass
    of each utterance is inside `seconds` and is not credited any bytes."""
    hbm = measured_peaks()[0]
    ach = search_steps * BEAM_BYTES_PER_STEP / seconds / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "algorithmic_bytes_per_search_step": BEAM_BYTES_PER_STEP, "search_steps": int(search_steps)}


def beam_rate(model, torch, T, n_utts, stop_limit, N=10, K=10):
    """beam-10 decode utterances / second (BASELINE metric ii), fp32-faithful mode, host feature buffers
    in, hypotheses out, on FRESH random-init weights (a model that has just fitted the synthetic unigram law emits EOS
    at once, which would time a 2-step search).  Random-init hypotheses never end in EOS, so every search runs exactly
    `stop_limit` steps: 175 (= max_pred, beam.py:104) is the worst case, 40 a typical Fisher hypothesis length."""
    from ast_b200.nn import beam_result_to_entries
    e = model._engine
    rng = np.random.default_rng(7)
    utts = [rng.standard_normal((1, T, D), dtype=np.float32) for _ in range(n_utts + 1)]
    steps = 0
    for i, x in enumerate(utts):
        if i == 1:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        r = e.beam_search(x, stop_limit, N, K)
        ent = beam_result_to_entries(r)
        if i >= 1:
            steps += r["n_steps"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"utts_per_s": n_utts / dt, "T": T, "stop_limit": stop_limit, "avg_steps": steps / n_utts, "n_utts": n_utts,
            "N": N, "K": K, "us_per_beam_step": 1e6 * dt / max(steps, 1), "roofline": beam_roofline(steps, dt)}


def beam_rate_pool(model, torch, T, n_utts, stop_limit, n_streams, N=10, K=10):
    """Throughput mode: the same searches, `n_streams` independent utterances in flight (ast_b200.beam.BeamPool)."""
    from ast_b200.beam import BeamPool
    from ast_b200.nn import beam_result_to_entries
    pool = BeamPool(model._engine, n=n_streams)
    rng = np.random.default_rng(7)
    utts = [rng.standard_normal((1, T, D), dtype=np.float32) for _ in range(n_utts)]
    pool.decode(utts[:n_streams], stop_limit, N, K, convert=beam_result_to_entries)      # warm every replica
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pool.decode(utts, stop_limit, N, K, convert=beam_result_to_entries)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert all(r is not None for r in res)
    return {"utts_per_s": n_utts / dt, "T": T, "stop_limit": stop_limit, "n_utts": n_utts, "N": N, "K": K, "utterances_in_flight": n_streams,
            "roofline": beam_roofline(n_utts * stop_limit, dt)}


def beam_rate_batch(model, torch, T, n_utts, stop_limit, G, N=10, K=10):
    """Throughput mode, second generation: G utterances searched in LOCK-STEP by one engine (ast_beam_search_batch): one pass over
    the decoder weights per step serves all G x N rows, equal-length utterances share an encoder batch.  Host feature buffers in,
    hypotheses (host lists) out."""
    from ast_b200.nn import beam_result_to_entries
    e = model._engine
    rng = np.random.default_rng(7)
    utts = [rng.standard_normal((1, T, D), dtype=np.float32) for _ in range(n_utts)]
    [beam_result_to_entries(r) for r in e.beam_search_batch(utts[:G], stop_limit, N, K)]     # warm (workspace, kernels, pinned staging)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    steps = 0
    for i in range(0, n_utts, G):
        res = e.beam_search_batch(utts[i:i + G], stop_limit, N, K)
        ent = [beam_result_to_entries(r) for r in res]
        steps += sum(r["n_steps"] for r in res)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    passes = steps / G                                                               # decoder-weight passes actually made
    return {"utts_per_s": n_utts / dt, "T": T, "stop_limit": stop_limit, "n_utts": n_utts, "N": N, "K": K, "utterances_per_search": G,
            "us_per_lockstep_step": 1e6 * dt / max(passes, 1), "roofline": beam_roofline(steps, dt),
            "roofline_note": "achieved = (utterance-steps x 33 MB) / time as SURVEY 8(d) defines it; the lock-step search reads the weights "
                             "once per step for all utterances, so figures above the HBM peak are possible and mean reuse, not bandwidth"}


def beam_cpu_rate(T, stop_limit, n_utts=1, N=10, K=10):
    """The reference's beam search (nn.py:235-322 restated in the numpy oracle) on the host cores, same inputs."""
    from oracle import ast_oracle as O
    cfg = model_cfg()
    P = O.init_params(cfg, D, seed=0)
    om = O.OracleModel(cfg, P, dtype=np.float32)
    rng = np.random.default_rng(7)
    utts = [rng.standard_normal((1, T, D), dtype=np.float32) for _ in range(n_utts)]
    t0 = time.perf_counter()
    for x in utts:
        om.decode_beam(x, stop_limit, N, K)
    dt = time.perf_counter() - t0
    return {"utts_per_s": n_utts / dt, "T": T, "stop_limit": stop_limit, "n_utts": n_utts, "cores": os.cpu_count(), "kind": "port"}



def tf32_peak_measured(torch, dev):
    """Dense TF32 peak measured on THIS GPU the way the driver measured the bf16 one (MEASURED_PEAKS.json `how`): cuBLAS through
    torch.matmul, fp32 operands with TF32 tensor-core math allowed, 8192^3, best of 10 with CUDA events.  Library call used as
    the yardstick only - nothing on the hot path calls cuBLAS."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def stage_profile(e, opt, resident, torch, peak, n=6):
    """Where a step's time goes, measured live with the library's stage events (CUDA events recorded on the caller's stream
    at stage boundaries, `stage_timing` option) on the MEDIAN-length batch of the timed region: per stage milliseconds,
    algorithmic GFLOP, achieved TFLOP/s and its fraction of the TF32 peak; for the two latency-bound stage
    families the time per sequence step of their persistent kernel."""
    order = sorted(range(len(resident)), key=lambda i: resident[i][0].shape[1])
    Xd, yd, bits, _ = resident[order[len(order) // 2]]
    B, T, _ = Xd.shape
    L = yd.shape[1]
    fl = step_flops(B, T, L)
    e.set_option("stage_timing", 1)
    acc = {}
    # rank 0 alone runs this profile: the data-parallel all-reduce hook must not run (the other ranks are not in the collective)
    hook, opt.pre_update = opt.pre_update, None
    try:
        for it in range(n + 2):
            e.forward_loss(Xd, yd, use_true=bits, noise_sigma=TRAIN_EXTRAS["speech_noise"])
            e.backward()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); opt.update(); ev1.record()
            torch.cuda.synchronize()
            if it >= 2:
                for name, ms in e.stage_times():
                    acc.setdefault(name, []).append(ms)
                acc.setdefault("optimizer", []).append(ev0.elapsed_time(ev1))
    finally:
        e.set_option("stage_timing", 0)
        opt.pre_update = hook
    med = {k: float(np.median(v)) for k, v in acc.items()}
    Tp, S = fl["Tp"], L - 1
    # (stage mark that ENDS the stage, label, FLOPs, dominant kernel(s), sequence steps of the persistent kernel)
    table = [("fwd:cnn_done", "CNN forward", fl["cnn"], "gemm_tc3 (3xTF32 tcgen05 implicit GEMM) + BN statistics / apply", None),
             ("fwd:encoder_done", "encoder forward", fl["enc"], "lstm_seq_fwd_tc_kernel x3 (persistent wavefront) + gemm_tc2 projections", Tp),
             ("fwd:decoder_done", "decoder forward", fl["dec"], "dec_seq2_fwd_kernel (one cooperative launch) + batched logits GEMM / CE", S),
             ("bwd:decoder_done", "decoder backward", 2 * fl["dec"], "dec_seq2_bwd_kernel (one cooperative launch) + weight-gradient GEMMs (side stream)", S),
             ("bwd:encoder_done", "encoder backward", 2 * fl["enc"], "lstm_seq_bwd_tc_kernel x3 (persistent wavefront) + gated dx GEMMs", Tp),
             ("bwd:cnn_done", "CNN backward", 2 * fl["cnn"], "gemm_tc (transposed-convolution dx, split-K dW) + BN backward", None),
             ("bwd:side_stream_joined", "join of side-stream weight gradients", 0.0, "gemm_tc2 split-K (grouped)", None),
             ("optimizer", "WD + global-norm clip + AMSGrad", 0.0, "opt_sqnorm + opt_amsgrad (HBM-bound: 9 x 4 B per parameter)", None)]
    out, total = [], 0.0
    for key, label, flops, kern, steps in table:
        if key not in med:
            continue
        ms = med[key]
        total += ms
        ent = {"stage": label, "ms": ms, "gflop": flops / 1e9, "dominant_kernels": kern}
        if flops > 0 and ms > 0:
            ent["achieved_tflops"] = flops / (ms * 1e-3) / 1e12
            ent["frac_of_tf32_peak"] = ent["achieved_tflops"] / peak
        if steps:
            ent["sequence_steps"] = steps
            ent["us_per_sequence_step"] = 1e3 * ms / steps
        out.append(ent)
    for ent in out:
        ent["share_of_step"] = ent["ms"] / total if total > 0 else None
    return {"batch": f"B{B} x T{T} x L{L} (median-length batch of the timed region)", "ms_total": total,
            "step_gflop": fl["step"] / 1e9, "stages": out}


def parity_block(torch, precision):
    """Same-run parity: the measured configuration - shipped geometry, dropout .3/.3, speech_noise .25 (explicit
    tensor), teach_ratio .8, the precision mode of this run - on a batch small enough for the float64 oracle, which is fed the
    device's own dropout masks (oracle/device_rng.py); then greedy and beam-10 hypotheses in the fp32-faithful decode mode against
    the float32 oracle.  The oracle is the CHECKER here (part of the cpu_baseline leg), never the thing timed."""
    from ast_b200.engine import Engine
    from ast_b200.nn import beam_result_to_entries
    from oracle import ast_oracle as O
    from oracle import device_rng as R
    cfg = model_cfg()
    B, T, Lmin, Lmax, seed = 6, 200, 8, 12, 4321
    from oracle.ref_golden_common import golden_params
    P = {k: (v.astype(np.float32) if v.dtype.kind == "f" else v) for k, v in golden_params(cfg, D, 3).items()}
    X, y, _ = O.synth_batch(B, T, D, V, Lmin, Lmax, seed=5, Tmin=T - 60)
    L = y.shape[1]
    rng = np.random.default_rng(6)
    noise = rng.normal(1.0, TRAIN_EXTRAS["speech_noise"], size=X.shape).astype(np.float32)
    bits = [True if not (0 < i < L - 2) else bool(rng.random() < TRAIN_EXTRAS["teach_ratio"]) for i in range(L - 1)]
    e = Engine(cfg, D, torch.cuda.current_device())
    for k in e.info:
        e.view(k).copy_(torch.as_tensor(P[k], device=e.device))
    e.weights_changed()
    e.set_option("exact", 0 if precision == "tf32" else 1)
    e.set_option("tc_gemm", 1 if precision == "tf32" else 0)
    e.set_option("seed", seed)
    loss = float(e.forward_loss(X, y, use_true=bits, noise=noise))
    e.backward()
    torch.cuda.synchronize()
    om = O.OracleModel(cfg, P, dtype=np.float64)
    om.dropout_masks = {k: v.astype(np.float64) for k, v in
                        R.training_masks(seed, 1, B, e.Tp, L - 1, 256, 512, 128, 3, DROPOUT[1], DROPOUT[0]).items()}
    want = float(om.forward_loss(X, y, tf_bits=bits, noise=noise))
    g = om.backward()
    gmax = gl2 = 0.0
    worst = ""
    for k in e.info:
        a, b = e.view(k, grad=True).cpu().numpy().astype(np.float64), g[k]
        r1 = float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
        r2 = float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
        if r1 > gmax:
            gmax, worst = r1, k
        gl2 = max(gl2, r2)
    # decode: fp32-faithful mode, EOS-boosted bias so that hypotheses end (and finished ones are carried along)
    P32 = dict(P)
    P32["out/b"] = P["out/b"].copy()
    P32["out/b"][O.EOS_ID] += 1.4
    om32 = O.OracleModel(cfg, P32, dtype=np.float32)
    e.view("out/b").copy_(torch.as_tensor(P32["out/b"], device=e.device))
    e.bn_state.zero_()
    for k in ("CNN_0_bn/avg_var", "CNN_1_bn/avg_var"):
        e.bn_view(k).fill_(1.0)
    e.weights_changed()
    e.set_option("exact", 1)
    e.set_option("tc_gemm", 0)
    pw = om32.predict(X, O.GO_ID, O.EOS_ID, 16)
    pg = e.predict(X, O.GO_ID, O.EOS_ID, 16).cpu().numpy()
    greedy_ok = bool(pw.shape == pg.shape and (pw == pg).all())
    nb_w = om32.decode_beam(X[0:1], 24, 10, 10)
    nb_g = beam_result_to_entries(e.beam_search(X[0:1], 24, 10, 10, O.GO_ID, O.EOS_ID))
    beam_ok = [list(map(int, h["hyp"])) for h in nb_w] == [h["hyp"] for h in nb_g]
    score_err = float(max(abs(float(a["score"]) - float(b["score"])) for a, b in zip(nb_w, nb_g)))
    beam_detail = None
    if not beam_ok:      # say what differs: a permutation among (near-)equal scores, or different token sequences
        sw, sg = sorted(tuple(map(int, h["hyp"])) for h in nb_w), sorted(tuple(h["hyp"]) for h in nb_g)
        beam_detail = {"same_set_of_hypotheses": sw == sg, "oracle_scores": [float(h["score"]) for h in nb_w],
                       "device_scores": [float(h["score"]) for h in nb_g]}
    return {"against": "numpy restatement of the reference (oracle/, pinned to the reference's own source by tests/golden/ref_*.npz)",
            "config": f"shipped geometry, B{B} x T{T} x L{L}, dropout .3/.3 (device masks injected into the oracle), speech_noise .25, "
                      f"teach_ratio .8, {precision} training mode; decode in the fp32-faithful mode",
            "loss_rel": abs(loss - want) / abs(want), "grad_rel_max": gmax, "grad_rel_max_tensor": worst, "grad_l2_rel_max": gl2,
            "greedy_identical": greedy_ok, "beam_identical": bool(beam_ok), "beam_score_abs_err": score_err,
            "beam_hyp_lens": [len(h["hyp"]) for h in nb_g], "beam_detail": beam_detail,
            "tolerances": {"loss_rel": 1e-3, "grad_rel": 1e-2}}


def sub_benchmark(torch, dev, name, feat_dim, vocab, shapes, precision, steps=6, warmup=3):
    """A short device-resident measurement of another configuration of BASELINE.json (same step: fwd + bwd + WD/clip/AMSGrad, shipped
    dropout / noise / teach_ratio, L2 flushed between steps): `shapes` = [(B, lengths, target lengths)] synthetic batches."""
    from ast_b200.config import es_en_20h_model_cfg
    from ast_b200.nn import Adam, GradientClipping, WeightDecay
    from ast_b200.seq2seq import SpeechEncoderDecoder, draw_use_true
    model = SpeechEncoderDecoder(dev.index, es_en_20h_model_cfg(vocab=vocab, dropout=DROPOUT), feat_dim=feat_dim)
    model.init_params(seed=0)
    e = model._engine
    e.set_option("exact", 0 if precision == "tf32" else 1)
    e.set_option("tc_gemm", 1 if precision == "tf32" else 0)
    opt = Adam(alpha=OPT_CFG["lr"]).setup(model)
    opt.add_hook(WeightDecay(OPT_CFG["l2"])); opt.add_hook(GradientClipping(OPT_CFG["grad_clip"]))
    rng = np.random.default_rng(17)
    batches = []
    for B, lens, tlens in shapes:
        T, L = int(max(lens)), int(max(tlens))
        X = np.zeros((B, T, feat_dim), np.float32)
        y = np.zeros((B, L), np.int32)
        for b in range(B):
            X[b, :lens[b]] = rng.standard_normal((lens[b], feat_dim), dtype=np.float32)
            y[b, :tlens[b]] = [1] + rng.integers(4, vocab, tlens[b] - 2).tolist() + [2]
        bits = np.asarray(draw_use_true(L, TRAIN_EXTRAS["teach_ratio"]), dtype=np.uint8)
        batches.append((torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), torch.from_numpy(bits).to(dev), int(sum(lens)), (B, T, L)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(b):
        e.forward_loss(b[0], b[1], use_true=b[2], noise_sigma=TRAIN_EXTRAS["speech_noise"])
        e.backward()
        opt.update()
    for i in range(warmup):
        step(batches[i % len(batches)])
    torch.cuda.synchronize()
    ms = frames = flops = 0.0
    for i in range(steps):
        b = batches[i % len(batches)]
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(b); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1); frames += b[3]; flops += step_flops(*b[4], feat_dim=feat_dim, vocab=vocab)["step"]
    del model, e
    return {"config": name, "precision": precision, "frames_per_s": frames / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
            "shapes": [f"B{b[4][0]}xT{b[4][1]}xL{b[4][2]}" for b in batches], "achieved_tflops": flops / (ms * 1e-3) / 1e12}


def run_ours(args):
    import torch
    from ast_b200 import dist as adist
    from ast_b200 import _lib
    from ast_b200.dataloader import DevicePacker
    from ast_b200.nn import Adam, GradientClipping, WeightDecay
    from ast_b200.seq2seq import SpeechEncoderDecoder, config as train_config, draw_use_true

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    rank, local_rank, world = adist.init_process_group()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    K, W = args.steps, max(args.warmup, 3)
    loader, gplan = make_plan(world, K + W)
    plan = adist.shard_batch_plan(gplan, rank, world)
    cfg = model_cfg()
    model = SpeechEncoderDecoder(local_rank, cfg, feat_dim=D)
    model._seed = 0
    model.init_params(seed=0)
    e = model._engine
    e.set_option("exact", 0 if args.precision == "tf32" else 1)
    e.set_option("tc_gemm", 1 if args.precision == "tf32" else 0)
    for kv in args.opt:
        k, v = kv.split("=")
        e.set_option(k, float(v))
    adist.broadcast_params_(e)
    opt = Adam(alpha=OPT_CFG["lr"]).setup(model)
    opt.add_hook(WeightDecay(OPT_CFG["l2"]))
    opt.add_hook(GradientClipping(OPT_CFG["grad_clip"]))
    adist.GradAllReduce(e, opt, world, overlap=not args.no_allreduce_overlap)
    packer = DevicePacker(dev)
    train_config.train = True
    lib = _lib.load()

    # ---- host batches (features, masks, labels, scheduled-sampling bits) -----------------------------
    random.seed(1000 + rank)
    host = []
    for utts, _, _ in plan:
        feats, keep, y, frames = host_batch(loader, utts)
        bits = np.asarray(draw_use_true(y.shape[1], TRAIN_EXTRAS["teach_ratio"]), dtype=np.uint8)
        host.append((feats, keep, y, bits, frames))
    max_sp = (DATA_CFG["buckets_num"] + 1) * DATA_CFG["buckets_width"]

    def step_resident(Xd, yd, bits):
        loss = e.forward_loss(Xd, yd, use_true=bits, noise_sigma=TRAIN_EXTRAS["speech_noise"])
        e.backward()
        opt.update()
        return loss

    # ---- (1) device-resident: inputs already in HBM when the timed region starts --------------------
    resident = [(packer.pack(f, max_sp, k), torch.from_numpy(y).to(dev), torch.from_numpy(bits).to(dev), fr) for f, k, y, bits, fr in host]
    maxT = max(x[0].shape[1] for x in resident); maxL = max(x[1].shape[1] for x in resident)
    e.ensure_workspace(B=BATCH, T=maxT, L=maxL, N=16, steps=1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    for i in range(W):
        step_resident(*resident[i][:3])
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    lib.ast_launch_count(1)
    evs = []
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(W, W + K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident(*resident[i][:3])
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.ast_launch_count(0))
    if world > 1:
        torch.distributed.barrier()
    clk = clocks.stop() if rank == 0 else None
    ms_dev = sum(a.elapsed_time(b) for a, b in evs)
    if rank == 0 and os.environ.get("AST_BENCH_VERBOSE"):
        for (a, b), x in zip(evs, resident[W:W + K]):
            print(f"  step B{x[0].shape[0]} x T{x[0].shape[1]} x L{x[1].shape[1]}: {a.elapsed_time(b):.3f} ms", file=sys.stderr)
    frames = sum(x[3] for x in resident[W:W + K])
    t_max = adist.max_over_ranks(ms_dev * 1e-3, dev)
    frames_all = adist.sum_over_ranks(frames, dev)
    value = frames_all / t_max

    # ---- (2) end to end through the public API: host buffers, pinned H2D pack, loss read-back --------
    random.seed(2000 + rank)
    for i in range(min(2, W)):
        f, k, y, bits, fr = host[i]
        float(step_resident(packer.pack(f, max_sp, k), torch.from_numpy(y).to(dev), bits))
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    from ast_b200.engine import AsyncScalar
    reader = AsyncScalar(dev)
    h2d = d2h = 0
    t0 = time.perf_counter()
    for i in range(W, W + K):
        f, k, y, bits, fr = host[i]
        Xd, yd, bd = packer.pack(f, max_sp, k, labels=y, bits=bits)     # one pinned staging buffer, one H2D copy, one kernel
        loss = step_resident(Xd, yd, bd)
        h2d += sum(x.nbytes for x in f) + sum(m.nbytes for m in k) + y.nbytes + bits.nbytes
        reader.push(loss)                                   # D2H of the loss into pinned memory, event behind it
        if len(reader) > 1:
            reader.pop(); d2h += 4                          # loss read one step late (asynchronous logging, nn.py:189)
    while len(reader):
        reader.pop(); d2h += 4
    torch.cuda.synchronize()
    t_e2e = adist.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = frames_all / t_e2e

    if rank != 0:
        return
    peaks = measured_peaks()
    hbm_peak = peaks[0]
    tf32_cublas = tf32_peak_measured(torch, dev)
    gemm_roof = roofline_dominant(e, torch, peaks)
    own_ceiling = max([k["achieved"] for k in gemm_roof.get("other_kernels", []) if "kernel ceiling" in k["shape"]] or [0.0])
    tf32_peak = max(tf32_cublas, own_ceiling)
    for ent in [gemm_roof] + gemm_roof.get("other_kernels", []):
        ent["peak"] = tf32_peak
        ent["frac"] = ent["achieved"] / tf32_peak
    gemm_roof["peak_source"] = "measured in this run: max(cuBLAS TF32 8192^3 via torch.matmul, this repo's cta_group::2 kernel at 8192x4096x4096)"
    flops_timed = sum(step_flops(x[0].shape[0], x[0].shape[1], x[1].shape[1])["step"] for x in resident[W:W + K])
    step_tflops = flops_timed / (ms_dev * 1e-3) / 1e12
    stages = stage_profile(e, opt, resident[W:W + K], torch, tf32_peak) if args.precision == "tf32" else None
    roof = {"bound": "tensor", "kernel": "whole training step (every kernel of the timed region; the step is dominated by the latency-bound "
                                         "persistent recurrence / decoder kernels, see stages)",
            "achieved": step_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": step_tflops / tf32_peak, "traffic": None,
            "algorithmic_gflop_timed_region": flops_timed / 1e9,
            "algorithmic_work": "SURVEY 8(d) forward-MAC formulas x 2 FLOP x 3 (fwd + bwd) per batch (B, padded T, L), summed over the timed batches",
            "peak_source": f"TF32 dense, measured in this run: cuBLAS (torch.matmul, allow_tf32) 8192^3 best of 10 = {tf32_cublas:.0f} TFLOP/s; "
                           f"this repo's tcgen05 cta_group::2 kernel at 8192x4096x4096 = {own_ceiling:.0f}; MEASURED_PEAKS.json ({peaks[3]}) has bf16 only "
                           f"({peaks[1]:.0f} burst / {peaks[2]:.0f} sustained)",
            "tf32_peak_cublas_measured": tf32_cublas, "tf32_peak_own_kernel_measured": own_ceiling,
            "stages": stages, "largest_gemm": gemm_roof}
    cpu_rate, cpu_done, cpu_t, cpu_desc = (None, 0, 0.0, [])
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        with blas_threads_all_cores():
            cpu_rate, cpu_done, cpu_t, cpu_desc = cpu_step_rate(loader, gplan[W:], 3, budget_s=20.0)
            parity = parity_block(torch, args.precision)
    subs = None
    if world == 1 and not args.no_sub:
        rs = np.random.default_rng(99)
        c1 = [(16, [1000] + rs.integers(921, 1001, 15).tolist(), rs.integers(20, 41, 16).tolist()) for _ in range(3)]
        c3 = [(32, sorted(rs.integers(lo, lo + 80, 32).tolist()), rs.integers(30, 120, 32).tolist()) for lo in (400, 800, 1200)]
        mid = [(x[0].shape[0], [x[0].shape[1]] * x[0].shape[0], [x[1].shape[1]] * x[0].shape[0]) for x in resident[W:W + 6]]
        subs = [sub_benchmark(torch, dev, "C1: one step, B16 x T~1000 x D40, L 20..40 (configs[0]'s shape on the GPU)", 40, V, c1, args.precision),
                sub_benchmark(torch, dev, "C3: asr_gpfr-shaped (D=13 MFCC -> F'=1, read-speech lengths 400..1280, B=32)", 13, V, c3, args.precision),
                sub_benchmark(torch, dev, "C2 batches in the fp32-faithful mode (3xTF32 / fp32 FMA everywhere): the mode greedy / beam identity is asserted in",
                              40, V, mid, "f32" if args.precision == "tf32" else "tf32")]
    line = {
        "metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32 tensor-core GEMMs (encoder recurrences: fp16 operands, same 11-bit significand), fp32 accumulate/state" if args.precision == "tf32" else "f32 (3xTF32 / fp32 FMA, fp32-faithful)",
        "data": "synthetic",
        "config": {"workload": "es_en_20h bucketed training steps (configs[1]): B=32/GPU, Fisher-shaped lengths (20x80-frame buckets, "
                               "truncated at 1680), D=40 fbank, V=1098, 2xCNN + 2x3-layer LSTM encoder + 3-layer attention decoder; "
                               "fwd + bwd + WD/clip/AMSGrad; dropout .3/.3, speech_noise .25, teach_ratio .8, zero_input .1",
                   "global_batch": BATCH * world, "parallelism": f"dp{world}", "grad_allreduce": ("3 buckets overlapped with backward" if not args.no_allreduce_overlap else "3 buckets after backward") if world > 1 else "none", "l2_flush_between_steps": True,
                   "frames_counted": "true (unpadded, post-truncation) input frames",
                   "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
                   # which encoder schedule this box allowed (persistent spin-wait wavefront needs 12 co-resident 8-CTA clusters;
                   # a GPU whose GPCs cannot hold them falls back to per-chunk launches, ~0.4 ms per step slower)
                   "encoder_schedule": {"persistent_wavefront": bool(e.get_option("enc_persist_active") == 1),
                                        "max_clusters_fwd": int(e.get_option("enc_max_clusters_fwd")),
                                        "max_clusters_bwd": int(e.get_option("enc_max_clusters_bwd")),
                                        "queues_ok": int(e.get_option("queues_ok")), "live_models": int(e.get_option("live_models")),
                                        "eager_loading": int(e.get_option("eager_loading")), "sms": int(e.get_option("num_sms"))}},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
    }
    if parity is not None:
        line["parity"] = parity
    if subs is not None:
        line["sub_benchmarks"] = subs
    if world == 1 and not args.no_beam:
        e.set_option("exact", 1)
        train_config.train = False
        model.init_params(seed=0)                     # fresh weights + BatchNorm running statistics
        e.bn_state.zero_()
        for k in ("CNN_0_bn/avg_var", "CNN_1_bn/avg_var"):
            e.bn_view(k).fill_(1.0)
        e.weights_changed()
        line["beam"] = {"metric": "beam10_decode_utts_per_sec", "mode": "exact fp32-faithful, batch-size-1 utterances (beam.py:111), "
                        "random-init weights: every search runs stop_limit steps",
                        "T1000_175steps_worst_case": beam_rate(model, torch, 1000, 4, 175),
                        "T1000_40steps": beam_rate(model, torch, 1000, 12, 40),
                        "T3000_40steps": beam_rate(model, torch, 3000, 6, 40),
                        "T1000_175steps_8_in_flight": beam_rate_pool(model, torch, 1000, 24, 175, 8),
                        "T1000_40steps_8_in_flight": beam_rate_pool(model, torch, 1000, 32, 40, 8),
                        "T3000_40steps_8_in_flight": beam_rate_pool(model, torch, 3000, 24, 40, 8),
                        "T1000_175steps_lockstep32": beam_rate_batch(model, torch, 1000, 64, 175, 32),
                        "T1000_40steps_lockstep32": beam_rate_batch(model, torch, 1000, 96, 40, 32),
                        "T3000_40steps_lockstep32": beam_rate_batch(model, torch, 3000, 64, 40, 32),
                        "in_flight_note": "ast_b200.beam.BeamPool: independent utterances decoded concurrently by engine replicas "
                                          "(own streams / host threads), hypotheses identical to the sequential loop"}
        if not args.no_cpu_baseline:
            line["beam"]["cpu_baseline_T1000_40steps"] = beam_cpu_rate(1000, 40, 2)
    if cpu_rate is not None:
        line["cpu_baseline"] = {"value": cpu_rate, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{cpu_done} training steps ({', '.join(cpu_desc)}) of the same batch plan in {cpu_t:.1f}s; "
                                          "numpy restatement of the reference (Chainer/CuPy not installable), OpenBLAS all cores"}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32", choices=["f32", "tf32"],
                    help="tf32: tcgen05 TF32 GEMMs + single-pass TF32 / FP16-operand recurrences (training mode, inside the parity "
                         "tolerances); f32: fp32-faithful everywhere (the decode / hypothesis-identity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the C1 / C3 / other-precision sub-benchmarks")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable; experiments), e.g. --opt enc_pchunk=8")
    ap.add_argument("--no-allreduce-overlap", action="store_true",
                    help="N>1: all-reduce the gradient buckets on the compute stream after backward instead of overlapped with it")
    args = ap.parse_args()
    # keep stdout clean for the ONE JSON line: library banners (e.g. NCCL's version line) go to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
