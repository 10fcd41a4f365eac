"""Torch (CPU) emulation of the kernel schedule of libbeatthis_sm100.so using the *packed*
parameters: same folds, same [B, F, L, C] layout, same slab-GEMM formulation of the
convolutions and of frontend.linear.  Test infrastructure: it lets the fold/layout logic of
beat_this_b200/weights.py be checked against the oracle on CPU, before any GPU time."""
import math

import torch
import torch.nn.functional as F


def _p(packed, name, *shape):
    return torch.from_numpy(packed[name]).view(*shape)


def _norm(x):
    return x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def _rope(t, cos, sin, pos):
    # t [..., n, 32] with positions pos [n]; interleaved pairs
    c, s = cos[pos], sin[pos]  # [n,16]
    x0, x1 = t[..., 0::2], t[..., 1::2]
    return torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1).reshape(t.shape)


def _attention(x, packed, p, C, cos, sin, freq_mode):
    """x [B, F, L, C]; sequences over F (freq_mode) or L."""
    heads = C // 32
    xn = _norm(x)
    qkv = xn @ _p(packed, p + ".wqkv", 3 * C, C).T
    gates = torch.sigmoid(xn @ _p(packed, p + ".wg", 32, C)[:heads].T + _p(packed, p + ".bg", 32)[:heads])
    B, Fq, L, _ = x.shape
    q, k, v = qkv.split(C, dim=-1)
    sh = lambda t: t.view(B, Fq, L, heads, 32)
    q, k, v = sh(q), sh(k), sh(v)
    if freq_mode:  # sequence axis = F
        q, k, v = [t.permute(0, 2, 3, 1, 4) for t in (q, k, v)]  # B L h F d
        pos = torch.arange(Fq)
    else:
        q, k, v = [t.permute(0, 1, 3, 2, 4) for t in (q, k, v)]  # B F h L d
        pos = torch.arange(L)
    q, k = _rope(q, cos, sin, pos), _rope(k, cos, sin, pos)
    o = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(32), dim=-1) @ v
    if freq_mode:
        o = o.permute(0, 3, 1, 2, 4)  # B F L h d
    else:
        o = o.permute(0, 1, 3, 2, 4)
    o = o * gates.unsqueeze(-1)
    return x + o.reshape(B, Fq, L, C) @ _p(packed, p + ".wout", C, C).T


def _ff(x, packed, p, C, mult=4):
    h = F.gelu(_norm(x) @ _p(packed, p + ".w1", mult * C, C).T + _p(packed, p + ".b1", mult * C))
    return x + h @ _p(packed, p + ".w2", C, mult * C).T + _p(packed, p + ".b2", C)


def forward(packed, hp, x, taps=None):
    """x [B, L, 128] (already chunked / zero padded) -> (beat, downbeat) [B, L]."""
    B, L, _ = x.shape
    cos, sin = _p(packed, "rope.cos", 1500, 16), _p(packed, "rope.sin", 1500, 16)
    # stem: BN1d per tap, zero time padding AFTER BN1d
    xb = x * _p(packed, "stem.bn1_scale", 128) + _p(packed, "stem.bn1_shift", 128)  # [B, L, 128]
    xb = F.pad(xb, (0, 0, 1, 1))  # time padding with true zeros
    w = _p(packed, "stem.w", 32, 4, 3)
    cols = torch.stack([xb[:, dt : dt + L] for dt in range(3)], dim=-1)  # [B, L, 128, 3]
    cols = cols.view(B, L, 32, 4, 3)  # f', df, dt
    h = torch.einsum("blfdt,cdt->bflc", cols, w) + _p(packed, "stem.bias", 32)
    h = F.gelu(h)  # [B, 32, L, 32]
    if taps is not None:
        taps["stem"] = h.clone()
    C, Fq = 32, 32
    for i in range(3):
        if hp.get("partial_transformers", True):
            h = _attention(h, packed, f"b{i}.attnF", C, cos, sin, True)
            if taps is not None: taps[f"b{i}.attnF"] = h.clone()
            h = _ff(h, packed, f"b{i}.ffF", C)
            if taps is not None: taps[f"b{i}.ffF"] = h.clone()
            h = _attention(h, packed, f"b{i}.attnT", C, cos, sin, False)
            if taps is not None: taps[f"b{i}.attnT"] = h.clone()
            h = _ff(h, packed, f"b{i}.ffT", C)
            if taps is not None: taps[f"b{i}.ffT"] = h.clone()
        # conv as 6 shifted slabs
        wc = _p(packed, f"b{i}.conv.w", 2 * C, 6 * C)
        hp_ = F.pad(h, (0, 0, 1, 1))  # pad time
        acc = 0
        for df in range(2):
            for dt in range(3):
                s = df * 3 + dt
                a = hp_[:, df::2, dt : dt + L]  # [B, F/2, L, C]
                acc = acc + a @ wc[:, s * C : (s + 1) * C].T
        h = F.gelu(acc + _p(packed, f"b{i}.conv.bias", 2 * C))
        C, Fq = 2 * C, Fq // 2
        if taps is not None: taps[f"b{i}.conv"] = h.clone()
    D = hp["transformer_dim"]
    wl = _p(packed, "lin.w", D, Fq * C)
    x_ = sum(h[:, f] @ wl[:, f * C : (f + 1) * C].T for f in range(Fq)) + _p(packed, "lin.b", D)  # [B, L, D]
    if taps is not None: taps["frontend"] = x_.clone()
    x_ = x_[:, None]  # [B,1,L,D]
    for l in range(hp["n_layers"]):
        x_ = _attention(x_, packed, f"l{l}.attn", D, cos, sin, False)
        if taps is not None: taps[f"l{l}.attn"] = x_[:, 0].clone()
        x_ = _ff(x_, packed, f"l{l}.ff", D, hp["ff_mult"])
        if taps is not None: taps[f"l{l}.ff"] = x_[:, 0].clone()
    o = _norm(x_[:, 0]) @ _p(packed, "head.w", 2, D).T + _p(packed, "head.b", 2)
    return (o[..., 0] + o[..., 1] if hp.get("sum_head", True) else o[..., 0]), o[..., 1]
