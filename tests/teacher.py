"""Teacher-forced, layer-local parity of a whole engine plan against fp32 PyTorch on the CPU (TEST
INFRASTRUCTURE).

Why not compare end to end only: every bf16 storage point is a discontinuity.  A continuous difference delta
in front of a rounding (two fp32 summation orders differ by ~1e-6) moves a fraction delta/ulp of the stored
values by one ulp, i.e. it comes out as sqrt(delta * ulp) >> delta, so two CORRECT implementations with the same
rounding points decorrelate to the full bf16 noise level within ~10 layers (measured layer by layer in
profiles/r02_parity_layers.txt: 2.6e-6 after the stem, 2e-3 after four layers, 3.7e-2 at the logits), and a
flipped ReLU mask moves a gradient element by O(1).  End-to-end tolerances can therefore never be tight enough
to tell arithmetic from bugs.  This checker is: for EVERY unit of the plan, take the tensors the engine
actually consumed (its stored bf16 inputs, its stored conv output z, its stored gradient operands) and recompute
that unit's outputs with PyTorch fp32 on the CPU.  Each comparison then crosses exactly one rounding:

  bf16 outputs (z, activation, g, dz, per-source data gradients, pooled / upsampled tensors)
        equal up to rare one-ulp flips of the fp32 summation order          -> rel. Frobenius <= 5e-4
  fp32 outputs (batch statistics, logits, weight / bias / BatchNorm gradients)  -> rel. Frobenius <= 1e-4
  integer outputs (max-pool values)                                             -> exact
  analytically zero outputs (a conv bias in front of BatchNorm)                 -> <= 1e-3 of the weight-gradient scale

Together with the chain rule this pins the whole forward / backward pass at any size the CPU finishes in
seconds (BASELINE's 512 x 512 at batch 2 included), without injecting anything of the engine into the oracle.
"""
import torch
import torch.nn.functional as F


def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _nchw(t):
    return t.detach().float().permute(0, 3, 1, 2).cpu().contiguous()


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _conv_input(u, cin):
    xs = []
    for a, up in u["srcs"]:
        t = _nchw(a.buf)
        if up == 2:
            t = F.interpolate(t, scale_factor=2, mode="nearest")
        xs.append(t)
    x = torch.cat(xs, 1) if len(xs) > 1 else xs[0]
    return x[:, :cin].contiguous() if x.shape[1] > cin else x     # image16: channels 3..15 are zero padding


def _gathered(act):
    """fp32 sum of the consumers' contributions to d(loss)/d(act) (2x2 sum-pooled where the consumer read the
    activation through nearest x2), as the engine's gather kernels form it before rounding."""
    total = None
    for buf, pool2 in act.contribs:
        t = _nchw(buf)
        if pool2:
            t = F.avg_pool2d(t, 2) * 4.0
        t = t[:, :act.shape[3]]
        total = t if total is None else total + t
    return total


def check_forward(eng, P, x_in, training=True):
    """P: name -> fp32 CPU tensor (parameters and buffers AS THEY WERE when the forward ran; BatchNorm running
    statistics are only read in eval mode).  Returns [(unit, what, kind, error)], kind in {'bf16', 'f32', 'exact',
    'zero'}."""
    rows = []
    for u in eng.units:
        kind = u["kind"]
        if kind == "maxpool":
            x = _nchw(u["src"].buf)
            want = F.max_pool2d(x, 3, 2, 1) if u.get("k", 3) == 3 else F.max_pool2d(x, 2)
            rows.append((u["out"].name, "maxpool", "exact", float((want != _nchw(u["out"].buf)).sum())))
            continue
        if kind == "upsample":
            want = F.interpolate(_nchw(u["src"].buf), scale_factor=2, mode="bilinear", align_corners=True)
            rows.append((u["out"].name, "bilinear x2", "bf16", _rel(_nchw(u["out"].buf), _r(want))))
            continue
        if kind == "convt":     # nn.ConvTranspose2d(cin, cout, 2, stride=2) with bias
            op = u["op"]
            want = F.conv_transpose2d(_nchw(u["srcs"][0][0].buf), _r(P[op["conv"] + ".weight"]), P[op["conv"] + ".bias"],
                                      stride=2)
            rows.append((op["out"], "transposed conv", "bf16", _rel(_nchw(u["out"].buf), _r(want))))
            continue
        if kind not in ("stem", "conv", "head"):
            continue
        op = u["op"]
        name = op["out"]
        # the pointwise head (csrc/pointwise_head.cu) multiplies in fp32 by the fp32 master weights
        W = P[op["conv"] + ".weight"] if u.get("pw") else _r(P[op["conv"] + ".weight"])
        if kind == "stem":
            x, stride, pad = _r(x_in.detach().float().cpu()), 2, 3
        else:
            x, stride, pad = _conv_input(u, W.shape[1]), u["s"], u["pad"]
        bias = P[op["conv"] + ".bias"] if (kind == "head" or op.get("bias")) else None
        z = F.conv2d(x, W, bias, stride, pad)
        if kind == "head":
            rows.append((name, "logits", "f32", _rel(u["out"].buf.detach().float().cpu(), z)))
            if u.get("up", 1) > 1 and training:
                full = F.interpolate(u["out"].buf.detach().float().cpu(), scale_factor=u["up"], mode="nearest")
                rows.append((name, "nearest x%d logits" % u["up"], "f32", _rel(u["result"].detach().float().cpu(), full)))
            continue
        res = _nchw(u["res"].buf) if u.get("res") is not None else None
        bn = op.get("bn")
        if bn and training:
            ez = _nchw(u["z"])
            rows.append((name, "z", "bf16", _rel(ez, _r(z))))
            mean = ez.double().mean((0, 2, 3))
            var = ez.double().var((0, 2, 3), unbiased=False)
            rows.append((name, "batch mean (error / std)", "f32",
                         ((u["mean"].cpu().double() - mean).norm() / var.sqrt().norm().clamp_min(1e-20)).item()))
            rows.append((name, "batch invstd", "f32", _rel(u["invstd"].cpu().double(), (var + 1e-5).rsqrt())))
            y = F.batch_norm(ez, None, None, P[bn + ".weight"], P[bn + ".bias"], True, 0.0, 1e-5)
        elif bn:
            y = F.batch_norm(z, P[bn + ".running_mean"], P[bn + ".running_var"], P[bn + ".weight"], P[bn + ".bias"],
                             False, 0.0, 1e-5)
        else:
            y = z
        if res is not None:
            y = y + res
        if op["relu"]:
            y = torch.relu(y)
        rows.append((name, "activation", "bf16", _rel(_nchw(u["out"].buf), _r(y))))
    return rows


def check_backward(eng, P, G, x_in):
    """After eng.backward() of an engine built with MMR_NO_ARENA_REUSE=1 (every backward temporary keeps its own
    memory).  P as in check_forward; G: name -> fp32 CPU gradient the engine produced.  Units are visited in
    backward order; each unit's gradient operands are the ENGINE's stored tensors."""
    rows = []
    for u in reversed(eng.units):
        kind = u["kind"]
        if kind in ("maxpool", "upsample"):
            src, out = u["src"], u["out"]
            if ("gin", id(u)) not in eng.arena_keys:
                continue
            x = _nchw(src.buf).requires_grad_(True)
            if kind == "maxpool":
                y = F.max_pool2d(x, 3, 2, 1) if u.get("k", 3) == 3 else F.max_pool2d(x, 2)
            else:
                y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
            y.backward(_gathered(out))
            rows.append((out.name, kind + " input grad", "bf16", _rel(_nchw(eng.arena_view(("gin", id(u)), src.shape)), _r(x.grad))))
            continue
        if kind == "convt":
            op, out, src = u["op"], u["out"], u["srcs"][0][0]
            g_eng = _nchw(eng.arena_view(("g", id(u)), out.shape))
            rows.append((op["out"], "g", "bf16", _rel(g_eng, _r(_gathered(out)))))
            rows.append((op["out"], "bias grad", "f32", _rel(G[op["conv"] + ".bias"], _gathered(out).sum((0, 2, 3)))))
            x = _nchw(src.buf).requires_grad_(True)
            W = _r(P[op["conv"] + ".weight"]).requires_grad_(True)
            F.conv_transpose2d(x, W, None, stride=2).backward(g_eng.contiguous())
            rows.append((op["out"], "weight grad", "f32", _rel(G[op["conv"] + ".weight"], W.grad)))
            rows.append((op["out"], "data grad -> %s" % src.name, "bf16",
                         _rel(_nchw(eng.arena_view(("dx", id(u), 0), src.shape)), _r(x.grad))))
            continue
        if kind not in ("stem", "conv", "head"):
            continue
        op = u["op"]
        name = op["out"]
        conv = op["conv"]
        out = u["out"]
        n, ho, wo = out.shape[0], out.shape[1], out.shape[2]
        cout, cpad = u["cout"], u.get("cpad", u["cout"])
        if kind == "head" and u.get("pw"):
            # one fp32 pass over the fp32 dlogits and the stored bf16 activation: head weight / bias gradients, the
            # source's data gradient (times its ReLU mask when the head is its only reader: that unit's dz) and the
            # column sums of the stored gradient (that unit's bias gradient)
            dl = u["dlogits"].detach().float().cpu()
            src = u["srcs"][0][0]
            L = eng._pw_fuse_target(u)
            x = _nchw(src.buf).requires_grad_(True)
            W = P[conv + ".weight"].clone().requires_grad_(True)
            F.conv2d(x, W, None).backward(dl)
            rows.append((name, "bias grad", "f32", _rel(G[conv + ".bias"], dl.sum((0, 2, 3)))))
            rows.append((name, "weight grad", "f32", _rel(G[conv + ".weight"], W.grad)))
            if L is not None:
                got = _nchw(eng.arena_view(("g", id(L)), src.shape))
                rows.append((name, "data grad x ReLU mask -> dz of %s" % src.name, "bf16",
                             _rel(got, _r(x.grad * (x.detach() > 0).float()))))
                if L["op"].get("bias"):
                    rows.append((L["op"]["out"], "bias grad", "f32", _rel(G[L["op"]["conv"] + ".bias"], got.sum((0, 2, 3)))))
            else:
                rows.append((name, "data grad -> %s" % src.name, "bf16",
                             _rel(_nchw(eng.arena_view(("dx", id(u), 0), src.shape)), _r(x.grad))))
            continue
        g_eng = _nchw(eng.arena_view(("g", id(u)), (n, ho, wo, cpad)))[:, :cout]
        if kind == "head":
            dl = u["dlogits"].detach().float().cpu()
            if u.get("up", 1) > 1:
                pooled = F.avg_pool2d(u["dlogits_full"].detach().float().cpu(), u["up"]) * float(u["up"] ** 2)
                rows.append((name, "sum-pooled dlogits", "f32", _rel(dl, pooled)))
            rows.append((name, "g (bf16 dlogits)", "bf16", _rel(g_eng, _r(dl))))
            rows.append((name, "bias grad", "f32", _rel(G[conv + ".bias"], dl.sum((0, 2, 3)))))
            dz = g_eng
        elif u.get("dz_by_reader") is not None:
            dz = g_eng      # written (masked, bias gradient included) by the pointwise head's backward: checked there
        else:
            total = _gathered(out)
            mask = (_nchw(out.buf) > 0).float() if op["relu"] else 1.0
            g_full = total * mask       # fp32: what the engine's reductions sum (they see g before it is stored)
            g_ref = _r(g_full)
            bn = u.get("bn")
            res_free = u.get("res") is None
            written = not (bn and op["relu"] and res_free and len(out.contribs) == 1 and not out.contribs[0][1])
            if written:
                rows.append((name, "g", "bf16", _rel(g_eng, g_ref)))
                g = g_eng
            else:       # single full-resolution contribution: the engine never materialises g (it IS the masked contribution)
                g = g_ref
            if bn:
                ez = _nchw(u["z"]).double()
                mean = ez.mean((0, 2, 3), keepdim=True)
                invstd = (ez.var((0, 2, 3), unbiased=False, keepdim=True) + 1e-5).rsqrt()
                xhat = (ez - mean) * invstd
                gd = g.double()
                dbeta = g_full.double().sum((0, 2, 3))
                dgamma = (g_full.double() * xhat).sum((0, 2, 3))
                cnt = float(n * ho * wo)
                gamma = P[bn + ".weight"].double().view(1, -1, 1, 1)
                dz_ref = gamma * invstd * (gd - dbeta.view(1, -1, 1, 1) / cnt - xhat * dgamma.view(1, -1, 1, 1) / cnt)
                dz = _nchw(eng.arena_view(("dz", id(u)), (n, ho, wo, cpad)))[:, :cout]
                rows.append((name, "dz", "bf16", _rel(dz, _r(dz_ref.float()))))
                rows.append((name, "BN weight grad", "f32", _rel(G[bn + ".weight"].double(), dgamma)))
                rows.append((name, "BN bias grad", "f32", _rel(G[bn + ".bias"].double(), dbeta)))
                if op.get("bias"):      # a conv bias in front of BatchNorm: its gradient is sum(dz) = 0 analytically
                    scale = G[conv + ".weight"].abs().max().item()
                    rows.append((name, "bias grad in front of BN / weight grad scale", "zero",
                                 G[conv + ".bias"].abs().max().item() / max(scale, 1e-20)))
            else:
                dz = g
                if op.get("bias"):
                    rows.append((name, "bias grad", "f32", _rel(G[conv + ".bias"], g.sum((0, 2, 3)))))
        W = _r(P[conv + ".weight"]).requires_grad_(True)
        if kind == "stem":
            x, stride, pad = _r(x_in.detach().float().cpu()), 2, 3
        else:
            x, stride, pad = _conv_input(u, W.shape[1]), u["s"], u["pad"]
        need_dx = kind != "stem" and any(("dx", id(u), si) in eng.arena_keys for si in range(len(u["srcs"])))
        x.requires_grad_(need_dx)
        F.conv2d(x, W, None, stride, pad).backward(dz.contiguous())
        rows.append((name, "weight grad", "f32", _rel(G[conv + ".weight"], W.grad)))
        if need_dx:
            c0 = 0
            hin, win = u["in_hw"]
            for si, (a, up) in enumerate(u["srcs"]):
                c = min(a.shape[3], W.shape[1] - c0)
                want = x.grad[:, c0:c0 + c]
                pooled = (up == 2 and u.get("halo") and getattr(eng, "pool_dgrad", False) and hin % 2 == 0 and win % 2 == 0
                          and int(u["dcfg"].get("direct", u["dcfg"]["sg"] < 64)) == int(u["dcfg"]["sg"] < 64))
                if pooled:   # the epilogue stores the 2x2 sum (fp32) of the conv-resolution gradient, rounded once
                    want = F.avg_pool2d(want, 2) * 4.0
                    got = _nchw(eng.arena_view(("dx", id(u), si), (a.shape[0], hin // 2, win // 2, a.shape[3])))[:, :c]
                else:
                    got = _nchw(eng.arena_view(("dx", id(u), si), (a.shape[0], hin, win, a.shape[3])))[:, :c]
                rows.append((name, "data grad -> %s" % a.name, "bf16", _rel(got, _r(want))))
                c0 += a.shape[3]
    return rows


def summarize(rows):
    """{kind: (worst error, unit, what)}"""
    out = {}
    for unit, what, kind, err in rows:
        if kind not in out or err > out[kind][0]:
            out[kind] = (err, unit, what)
    return out


def snapshot(model):
    P = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    return P


def grads_of(model):
    return {k: p.grad.detach().float().cpu().clone() for k, p in model.named_parameters() if p.grad is not None}
