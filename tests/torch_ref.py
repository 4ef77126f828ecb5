"""Independent torch-autograd (CPU) re-expression of the seq2seq.py graph.

Test helper only: cross-checks the oracle's hand-written backward (SURVEY 8c pin 3).
It uses library ops (conv2d, batch_norm, cross_entropy) rather than the oracle's im2col /
explicit formulas, so an agreement is evidence for both.
"""
import torch
import torch.nn.functional as F


def torch_forward_loss(cfg, params, X, y, tf_bits=None, dtype=torch.float64):
    P = {k: torch.tensor(v, dtype=dtype, requires_grad=True)
         for k, v in params.items() if v.dtype.kind == "f" and not k.endswith(("avg_mean", "avg_var"))}
    r = cfg["rnn_config"]
    H, E, A, V, nl = r["hidden_units"], r["embedding_units"], r["attn_units"], r["dec_vocab_size"], r["enc_layers"]
    x = torch.tensor(X, dtype=dtype)[:, None]
    for i, l in enumerate(cfg["cnn_config"]["cnn_layers"]):
        x = F.conv2d(x, P[f"CNN_{i}/W"], None, stride=tuple(l["stride"]), padding=tuple(l["pad"]))
        x = F.batch_norm(x, None, None, P[f"CNN_{i}_bn/gamma"], P[f"CNN_{i}_bn/beta"], training=True, eps=2e-5)
        x = F.relu(x)
    B, C, Tp, Fp = x.shape
    xr = x.permute(2, 0, 1, 3).reshape(Tp, B, C * Fp)

    def cell(name, inp, h, c):
        g = inp @ P[f"{name}/upward/W"].T + P[f"{name}/upward/b"]
        if h is not None:
            g = g + h @ P[f"{name}/lateral/W"].T
        g = g.reshape(B, -1, 4)
        a, i_, f, o = torch.tanh(g[..., 0]), torch.sigmoid(g[..., 1]), torch.sigmoid(g[..., 2]), torch.sigmoid(g[..., 3])
        c = a * i_ + (f * c if c is not None else 0)
        return o * torch.tanh(c), c

    outs, fin = {}, {}
    for stack, order in (("enc", list(range(Tp))), ("rev_enc", [(-i) % Tp for i in range(Tp)])):
        hs = [None] * nl; cs = [None] * nl; seq = []
        for t in order:
            inp = xr[t]
            for l in range(nl):
                hs[l], cs[l] = cell(f"L{l}_{stack}", inp, hs[l], cs[l])
                inp = hs[l]
            seq.append(inp)
        outs[stack] = torch.stack(seq)
        fin[stack] = (hs, cs)
    enc = torch.cat((outs["enc"], torch.flip(outs["rev_enc"], (0,))), dim=2).transpose(0, 1)
    dh = [torch.cat((fin["enc"][0][l], fin["rev_enc"][0][l]), 1) for l in range(nl)]
    dc = [torch.cat((fin["enc"][1][l], fin["rev_enc"][1][l]), 1) for l in range(nl)]
    yT = torch.tensor(y, dtype=torch.long).T
    L = yT.shape[0]
    ht = torch.zeros(B, A, dtype=dtype)
    w = torch.ones(V, dtype=dtype); w[0] = 0
    loss = 0
    dec_in = None
    for i in range(L - 1):
        if tf_bits is None or tf_bits[i] or i == 0 or i >= L - 2:
            dec_in = yT[i]
        inp = torch.cat((P["embed_dec/W"][dec_in], ht), 1)
        for l in range(nl):
            dh[l], dc[l] = cell(f"L{l}_dec", inp, dh[l], dc[l])
            inp = dh[l]
        q = inp @ P["attn_Wa/W"].T + P["attn_Wa/b"]
        s = torch.bmm(enc, q[:, :, None])
        al = torch.softmax(s, dim=1)
        cv = torch.bmm(enc.transpose(1, 2), al)[:, :, 0]
        ht = torch.tanh(torch.cat((cv, inp), 1) @ P["context/W"].T + P["context/b"])
        logits = ht @ P["out/W"].T + P["out/b"]
        dec_in = logits.argmax(1)
        # chainer: sum of w[t]*nll / B  (not torch's weighted-mean normalisation)
        loss = loss + F.cross_entropy(logits, yT[i + 1], weight=w, reduction="sum") / B
    loss.backward()
    return float(loss), {k: v.grad.numpy() for k, v in P.items()}, enc.detach().numpy()
