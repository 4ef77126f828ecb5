"""GPU bring-up diagnostics: compares every stage of the CUDA path with the oracle and prints the
max abs / rel error per intermediate, so a single gpurun call pinpoints a failing kernel.
Usage (GPU box):  python tools/stage_check.py [--full]
"""
import argparse
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ast_b200 import _lib                      # noqa: E402
from ast_b200._lib import check, ptr           # noqa: E402
from ast_b200.engine import Engine             # noqa: E402
from oracle import ast_oracle as O             # noqa: E402

dev = torch.device("cuda", 0)
RES = []


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf"), float("inf")
    d = np.abs(a - b).max() if a.size else 0.0
    return d, d / (np.abs(b).max() + 1e-30)


def report(name, got, want, tol=1e-3):
    d, r = rel(got, want)
    ok = r <= tol
    RES.append((name, ok))
    print(f"[{'ok' if ok else 'FAIL'}] {name:40s} maxabs {d:.3e}  rel {r:.3e}  (tol {tol:g})", flush=True)
    return ok


def section(fn):
    def wrapped(*a, **k):
        print(f"\n=== {fn.__name__} ===", flush=True)
        try:
            t = time.time()
            fn(*a, **k)
            torch.cuda.synchronize()
            print(f"    ({time.time() - t:.2f}s)", flush=True)
        except Exception:
            RES.append((fn.__name__, False))
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("CUDA context is broken:", e, flush=True)
                summary()
                sys.exit(2)
    return wrapped


def stream():
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@section
def check_gemm():
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for (ta, tb, M, N, K) in [(0, 1, 300, 200, 120), (0, 1, 257, 129, 117), (0, 0, 130, 260, 72), (1, 0, 117, 90, 1000),
                              (1, 1, 64, 64, 64), (0, 1, 4000, 1024, 256)]:
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        Bm = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        C0 = rng.standard_normal((M, N)).astype(np.float32)
        want = (A.T if ta else A).astype(np.float64) @ (Bm.T if tb else Bm).astype(np.float64) * 1.0 + 0.5 * C0 + bias
        dA, dB, db, dC = (torch.as_tensor(x, device=dev) for x in (A, Bm, bias, C0))
        check(lib.ast_gemm(0, ta, tb, M, N, K, 1.0, ptr(dA), A.shape[1], ptr(dB), Bm.shape[1], 0.5, ptr(dC), N, ptr(db), stream()))
        report(f"sgemm ta={ta} tb={tb} {M}x{N}x{K}", dC.cpu().numpy(), want, 1e-5)


def lstm_ref(G, Wl, T, B, h):
    Hs = np.zeros((T + 1, B, h), np.float64); Cs = np.zeros((T + 1, B, h), np.float64)
    act = np.zeros((T, B, 4 * h), np.float64)
    for t in range(T):
        g = G[t] + Hs[t] @ Wl.T
        c, hh, (a, i, f, o) = O.lstm_cell(Cs[t], g)
        Hs[t + 1], Cs[t + 1] = hh, c
        act[t] = np.stack((a, i, f, o), axis=2).reshape(B, 4 * h)
    return Hs, Cs, act


@section
def check_lstm_seq():
    lib = _lib.load()
    rng = np.random.default_rng(1)
    for (T, B, h) in [(7, 16, 256), (5, 3, 128), (6, 32, 256), (4, 20, 64)]:
        G = rng.standard_normal((T, B, 4 * h)).astype(np.float32)
        Wl = (rng.standard_normal((4 * h, h)) / np.sqrt(h)).astype(np.float32)
        Hs, Cs, act = lstm_ref(G.astype(np.float64), Wl.astype(np.float64), T, B, h)
        dG = torch.as_tensor(G, device=dev); dW = torch.as_tensor(Wl, device=dev)
        dH = torch.zeros(T + 1, B, h, device=dev); dC = torch.zeros(T + 1, B, h, device=dev)
        out = torch.zeros(T, B, h, device=dev)
        check(lib.ast_lstm_seq(0, ptr(dG), ptr(dW), ptr(dH), ptr(dC), ptr(out), T, B, h, None, None, 1, stream()), "lstm fwd")
        torch.cuda.synchronize()
        report(f"lstm_seq fwd H  T{T} B{B} h{h}", dH.cpu().numpy(), Hs, 2e-5)
        report(f"lstm_seq fwd C  T{T} B{B} h{h}", dC.cpu().numpy(), Cs, 2e-5)
        report(f"lstm_seq fwd act T{T} B{B} h{h}", dG.cpu().numpy(), act, 2e-5)
        report(f"lstm_seq fwd out T{T} B{B} h{h}", out.cpu().numpy(), Hs[1:], 2e-5)
        # backward
        dout = rng.standard_normal((T, B, h)); dhf = rng.standard_normal((B, h)); dcf = rng.standard_normal((B, h))
        dGref = np.zeros((T, B, 4 * h)); dh = dhf.copy(); dc = dcf.copy()
        for t in reversed(range(T)):
            a4 = act[t].reshape(B, h, 4)
            dg, dc = O.lstm_cell_bwd(dout[t] + dh, dc, Cs[t], Cs[t + 1], (a4[:, :, 0], a4[:, :, 1], a4[:, :, 2], a4[:, :, 3]))
            dGref[t] = dg
            dh = dg @ Wl.astype(np.float64)
        ddout = torch.as_tensor(dout.astype(np.float32), device=dev)
        ddh = torch.as_tensor(dhf.astype(np.float32), device=dev); ddc = torch.as_tensor(dcf.astype(np.float32), device=dev)
        check(lib.ast_lstm_seq(1, ptr(dG), ptr(dW), ptr(dH), ptr(dC), ptr(ddout), T, B, h, ptr(ddh), ptr(ddc), 1, stream()), "lstm bwd")
        torch.cuda.synchronize()
        report(f"lstm_seq bwd dG T{T} B{B} h{h}", dG.cpu().numpy(), dGref, 5e-5)


@section
def check_softmax_ce():
    lib = _lib.load()
    rng = np.random.default_rng(2)
    B, V = 16, 1098
    ld = 1104
    z = (3 * rng.standard_normal((B, V))).astype(np.float32)
    t = rng.integers(0, V, B).astype(np.int32); t[3] = 0; t[7] = 0
    w = np.ones(V, np.float32); w[0] = 0
    loss, dz = O.softmax_cross_entropy(z.astype(np.float64), t.astype(np.int64), w.astype(np.float64))
    zp = np.zeros((B, ld), np.float32); zp[:, :V] = z
    dzp = torch.as_tensor(zp, device=dev); dt = torch.as_tensor(t, device=dev)
    rl = torch.zeros(B, device=dev); am = torch.zeros(B, dtype=torch.int32, device=dev)
    check(lib.ast_softmax_ce(ptr(dzp), ld, ptr(dt), B, V, ptr(rl), ptr(am), stream()))
    report("softmax_ce loss", rl.sum().item(), loss, 1e-5)
    report("softmax_ce dz", dzp.cpu().numpy()[:, :V], dz, 1e-5)
    report("softmax_ce argmax", am.cpu().numpy(), z.argmax(1), 0)


def load_params(e, P):
    for k in e.info:
        e.view(k).copy_(torch.as_tensor(P[k], device=dev))
    e.weights_changed()


def model_case(B, T, D, V, Lmin, Lmax, seed, bits_mode, label, tol_loss=1e-4, tol_grad=2e-3):
    cfg = O.default_model_cfg(vocab=V)
    P = O.init_params(cfg, D, seed=seed)
    rng = np.random.default_rng(seed + 100)
    for k in P:
        if k.endswith(("gamma", "beta", "/b")):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    X, y, lens = O.synth_batch(B, T, D, V, Lmin, Lmax, seed=seed + 1, Tmin=max(T - 79, 1))
    L = y.shape[1]
    bits = None
    if bits_mode:
        bits = [bool(b) for b in np.random.default_rng(5).random(L - 1) < 0.6]
    om = O.OracleModel(cfg, P, dtype=np.float64)
    t0 = time.time(); loss = om.forward_loss(X, y, tf_bits=bits); g = om.backward(); t_or = time.time() - t0
    e = Engine(cfg, D, 0)
    load_params(e, P)
    e.set_option("exact", 1)
    lg = e.forward_loss(X, y, use_true=None if bits is None else [1 if (b or i == 0 or i >= L - 2) else 0 for i, b in enumerate(bits)])
    torch.cuda.synchronize()
    print(f"  [{label}] oracle(f64) {t_or:.1f}s  loss oracle {float(loss):.6f} gpu {float(lg):.6f}", flush=True)
    # stage by stage
    Tp = e.Tp
    cache = om._cnn_cache
    Fp = O.cnn_shapes(cfg, T, D)[0][11]
    T1 = O.cnn_shapes(cfg, T, D)[0][10]
    xhat0, inv0 = cache[0][2], cache[0][3]
    mu0 = None
    raw0 = e.debug_fetch("raw0").cpu().numpy().reshape(B, Fp, T1, -1)
    cols0, W0 = cache[0][1], om.p["CNN_0/W"]
    want_raw0 = (cols0 @ W0.reshape(W0.shape[0], -1).T).reshape(B, T1, Fp, -1).transpose(0, 2, 1, 3)
    report(f"{label} raw0 (conv0)", raw0, want_raw0, 1e-4)
    cols1, W1 = cache[1][1], om.p["CNN_1/W"]
    raw1 = e.debug_fetch("raw1").cpu().numpy()
    Rs = raw1.size // (B * Fp * 512)
    raw1 = raw1.reshape(B, Fp, Rs, 512)[:, :, :Tp]
    want_raw1 = (cols1 @ W1.reshape(W1.shape[0], -1).T).reshape(B, Tp, Fp, -1).transpose(0, 2, 1, 3)
    report(f"{label} raw1 (conv1)", raw1, want_raw1, 1e-4)
    rnn_in = e.debug_fetch("rnn_in").cpu().numpy().reshape(Tp, B, -1)
    a1 = np.maximum(om.p["CNN_1_bn/gamma"] * cache[1][2] + om.p["CNN_1_bn/beta"], 0).reshape(B, Tp, Fp, -1)
    want_rnn = a1.transpose(1, 0, 3, 2).reshape(Tp, B, -1)
    report(f"{label} rnn_in (bn+relu+relayout)", rnn_in, want_rnn, 1e-4)
    report(f"{label} bn running stats", e.bn_state.cpu().numpy(),
           np.concatenate([om.p["CNN_0_bn/avg_mean"], om.p["CNN_0_bn/avg_var"], om.p["CNN_1_bn/avg_mean"], om.p["CNN_1_bn/avg_var"]]), 1e-4)
    for l in range(3):
        for d, stack in enumerate(("enc", "rev_enc")):
            Hs = e.debug_fetch(f"H_{l}{d}").cpu().numpy().reshape(Tp + 1, B, -1)
            want = np.stack([om._enc_cache["fwd" if d == 0 else "rev"][i][l][0][3] for i in range(Tp)])  # c
            Cs = e.debug_fetch(f"C_{l}{d}").cpu().numpy().reshape(Tp + 1, B, -1)
            report(f"{label} C L{l}_{stack}", Cs[1:], want, 1e-4)
    report(f"{label} enc_states", e.enc_states().cpu().numpy(), om.enc_states, 1e-4)
    ht = e.debug_fetch("ht").cpu().numpy().reshape(L - 1, B, -1)
    report(f"{label} ht (all steps)", ht, np.stack([c[7] for c in om._dec_cache]), 1e-4)
    report(f"{label} step argmax", e.step_argmax().cpu().numpy(), np.stack(om.step_argmax), 0)
    rl = e.debug_fetch("row_loss").cpu().numpy().reshape(L - 1, B).sum(1)
    report(f"{label} step losses", rl, np.array(om.step_losses), 1e-4)
    report(f"{label} loss", float(lg), float(loss), tol_loss)
    e.backward()
    torch.cuda.synchronize()
    worst = 0
    for k in e.info:
        d, r = rel(e.view(k, grad=True).cpu().numpy(), g[k])
        worst = max(worst, r)
        if r > tol_grad:
            report(f"{label} grad {k}", e.view(k, grad=True).cpu().numpy(), g[k], tol_grad)
    report(f"{label} all grads (worst rel {worst:.2e})", worst, 0.0, tol_grad) if False else None
    ok = worst <= tol_grad
    RES.append((f"{label} grads", ok))
    print(f"[{'ok' if ok else 'FAIL'}] {label} worst grad rel err {worst:.3e} (tol {tol_grad:g})", flush=True)
    # optimizer
    opt = O.OracleAMSGrad(om.p)
    gc = {k: v.copy() for k, v in g.items()}
    opt.update(om.p, gc);
    m = torch.zeros_like(e.params); v = torch.zeros_like(e.params); vh = torch.zeros_like(e.params)
    e.opt_step(m, v, vh, 1, 1e-3, 1e-4, 2.0)
    report(f"{label} grad norm", e.last_grad_norm(), opt.last_norm, 1e-3)
    worst = 0
    for k in e.info:
        d, r = rel(e.view(k).cpu().numpy(), om.p[k])
        worst = max(worst, d)
    ok = worst < 1.1e-3     # first Adam step moves every weight by ~lr*sign(g): only bounded by lr for tiny |g|
    RES.append((f"{label} params after update", ok))
    print(f"[{'ok' if ok else 'FAIL'}] {label} params after AMSGrad step: worst abs diff {worst:.3e}", flush=True)
    return e, om, cfg, P, X, y


@section
def check_model_small():
    model_case(4, 203, 40, 300, 5, 9, 11, False, "small/tf")
    model_case(3, 100, 13, 59, 5, 9, 12, True, "small13/ss")
    model_case(17, 150, 40, 120, 4, 6, 13, True, "B17/ss")


@section
def check_decode():
    cfg = O.default_model_cfg(vocab=200)
    D = 40
    P = O.init_params(cfg, D, seed=21)
    P["out/b"][O.EOS_ID] += 2.5      # make EOS reachable so finished-hyp carry-over is exercised
    X, y, lens = O.synth_batch(3, 160, D, 200, 5, 9, seed=22, Tmin=120)
    om = O.OracleModel(cfg, P, dtype=np.float32)
    e = Engine(cfg, D, 0)
    load_params(e, P)
    want = om.predict(X, O.GO_ID, O.EOS_ID, 20)
    got = e.predict(X, O.GO_ID, O.EOS_ID, 20).cpu().numpy()
    print("  greedy oracle", want.tolist()); print("  greedy gpu   ", got.tolist(), flush=True)
    report("greedy predict tokens", got, want, 0)
    for (N, K) in [(4, 3), (10, 10)]:
        nb = om.decode_beam(X[:1, :lens[0]], 25, N, K)
        r = e.beam_search(X[:1, :lens[0]], 25, N, K)
        from ast_b200.nn import beam_result_to_entries
        ent = beam_result_to_entries(r, O.GO_ID)
        print(f"  beam N{N} K{K}: oracle steps/hyps", [len(x['hyp']) for x in nb], "gpu n_steps", r["n_steps"], flush=True)
        ok = len(ent) == len(nb) and all(a["hyp"] == b["hyp"] for a, b in zip(ent, nb))
        RES.append((f"beam N{N} K{K} hyps identical", ok))
        print(f"[{'ok' if ok else 'FAIL'}] beam N{N} K{K} hyps identical", flush=True)
        if not ok:
            for a, b in zip(ent, nb):
                print("     gpu", a["hyp"], float(a["score"]), "| oracle", b["hyp"], float(b["score"]))
        report(f"beam N{N} K{K} scores", [float(a["score"]) for a in ent], [float(b["score"]) for b in nb], 1e-4)
        if ok:
            report(f"beam N{N} K{K} attn_history[0]", np.stack(ent[0]["attn_history"]), np.stack(nb[0]["attn_history"]), 1e-3)


@section
def check_model_full():
    e, om, cfg, P, X, y = model_case(16, 1000, 40, 1098, 20, 40, 0, False, "C1", tol_loss=1e-4, tol_grad=5e-3)
    # timing (exact mode)
    lens = None
    m = torch.zeros_like(e.params); v = torch.zeros_like(e.params); vh = torch.zeros_like(e.params)
    for mode in (1, 0):
        e.set_option("exact", mode)
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            e.forward_loss(X, y); torch.cuda.synchronize(); t1 = time.time()
            e.backward(); torch.cuda.synchronize(); t2 = time.time()
            e.opt_step(m, v, vh, it + 2, 1e-3, 1e-4, 2.0); torch.cuda.synchronize(); t3 = time.time()
        print(f"  C1 exact={mode}: fwd {1e3*(t1-t0):.2f} ms  bwd {1e3*(t2-t1):.2f} ms  opt {1e3*(t3-t2):.2f} ms", flush=True)


def summary():
    bad = [n for n, ok in RES if not ok]
    print(f"\n==== {len(RES) - len(bad)}/{len(RES)} checks passed ====")
    for n in bad:
        print("   FAILED:", n)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    todo = [check_gemm, check_lstm_seq, check_softmax_ce, check_model_small, check_decode]
    if a.full:
        todo.append(check_model_full)
    for fn in todo:
        if a.only and a.only not in fn.__name__:
            continue
        fn()
    summary()
