"""CPU emulation of the libcpcsv.so entry points, at the tensor-level interface of
``cpcsv_b200.ops`` (TEST INFRASTRUCTURE).

Purpose: the host-side logic of the product (tap tables, sub-pixel phases, parity views,
weight re-layouts, BN/activation plumbing, the autograd engine) can be verified against the
oracle on a machine without a GPU, by monkeypatching ``cpcsv_b200.ops`` with these functions
inside a test.  The product itself never imports this file and has no CPU path; on the GPU
box the same tests run against the real kernels.

Each function restates the contract documented in include/cpcsv.h, including the TMA
semantics the kernels rely on (out-of-range coordinates read as zero).
"""
import torch

TORCH16 = {0: torch.float16, 1: torch.bfloat16}


def install(monkeypatch):
    import cpcsv_b200.ops as ops
    for name, fn in list(globals().items()):
        if callable(fn) and not name.startswith("_") and hasattr(ops, name) and name not in (
                "View", "GemmJob", "install", "AdamHyper"):
            monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops, "_ptr", lambda t, dtype=None: None)


# ------------------------------------------------------------------------------ conv_gemm
def _dense_view(view):
    """Materialise a View (c, w, p, h, n; byte strides) as float64 [n, h, p, w, c]."""
    t = view.tensor
    flat = t.reshape(-1) if t.is_contiguous() else None
    base = torch.empty(0, dtype=t.dtype).set_(t.untyped_storage(), 0, (t.untyped_storage().nbytes() // 2,), (1,))
    off = t.storage_offset()
    c, w, p, h, n = view.dims
    s = [x // 2 for x in view.strides]
    need = off + (n - 1) * s[4] + (h - 1) * s[3] + (p - 1) * s[2] + (w - 1) * s[1] + (c - 1) + 1
    assert need <= base.numel(), ("view exceeds storage", need, base.numel())
    del flat
    return torch.as_strided(base, (n, h, p, w, c), (s[4], s[3], s[2], s[1], 1), off).double()


def _gather(dense, n_idx, h_idx, p, w_idx, c0, width):
    """dense [n,h,p,w,c] -> [len(n), len(h), len(w), width] with zero fill outside."""
    N, H, P, W, Cc = dense.shape
    out = torch.zeros(len(n_idx), len(h_idx), len(w_idx), width, dtype=torch.float64)
    if p < 0 or p >= P:
        return out
    nv = (n_idx >= 0) & (n_idx < N)
    hv = (h_idx >= 0) & (h_idx < H)
    wv = (w_idx >= 0) & (w_idx < W)
    c_lo, c_hi = max(c0, 0), min(c0 + width, Cc)
    if c_hi <= c_lo or not (nv.any() and hv.any() and wv.any()):
        return out
    sub = dense[n_idx[nv]][:, h_idx[hv]][:, :, p][:, :, w_idx[wv]][..., c_lo:c_hi]
    tmp = torch.zeros(int(nv.sum()), int(hv.sum()), int(wv.sum()), width, dtype=torch.float64)
    tmp[..., c_lo - c0:c_hi - c0] = sub
    out[nv.nonzero().view(-1, 1, 1), hv.nonzero().view(1, -1, 1), wv.nonzero().view(1, 1, -1)] = tmp
    return out


def conv_gemm(job):
    N, H, W = job.grid
    tn, th, tw = job.tile
    assert tn * th * tw == (128 if job.mode == 0 else 64)
    assert 16 <= job.block_n <= 256 and job.block_n % 16 == 0
    assert job.n_valid % 4 == 0 and job.n_valid <= job.n_tiles * job.block_n
    assert job.planes in (1, 2) and len(job.taps) <= 16
    out = job.out
    assert out.dtype == torch.float32
    # element offsets below are relative to the output POINTER (a strided view's first element)
    n_store = out.untyped_storage().nbytes() // 4
    out_flat = torch.empty(0, dtype=torch.float32).set_(out.untyped_storage(), 0, (n_store,), (1,))[out.storage_offset():]
    alpha = float(job.alpha) if job.alpha is not None else 1.0
    A = [_dense_view(v) for v in job.a[:job.planes]]
    B = [_dense_view(v) for v in job.b[:job.planes]]
    accumulate = job.accumulate
    if job.splits > 1 and not accumulate:
        # what ops.conv_gemm does for a split-K job: the WHOLE output tensor is cleared, partial tiles are combined
        # with red.add (so two split jobs that share one output tensor must both be 'accumulate')
        out.zero_()
        accumulate = True
    ncols = job.n_tiles * job.block_n
    if job.mode == 0:
        assert all(s % 4 == 0 for s in job.out_strides)
        K = job.k_blocks * 64
        n_idx, h_idx, w_idx = torch.arange(N), torch.arange(H), torch.arange(W)
        osn, osh, osw = job.out_strides
        pix_off = (n_idx.view(-1, 1, 1) * osn + h_idx.view(1, -1, 1) * osh + w_idx.view(1, 1, -1) * osw)
        for g in range(job.groups):
            acc = torch.zeros(N, H, W, ncols, dtype=torch.float64)
            for t in range(job.taps_per_group):
                a4, b4, _ = job.taps[g * job.taps_per_group + t]
                Ap = [_gather(d, n_idx, h_idx + a4[3], a4[2], w_idx + a4[1], a4[0], K) for d in A]
                # weights: dense [1,1,1,rows,k]
                Bp = []
                for d in B:
                    rows = d.shape[3]
                    r0 = b4[0]
                    blk = torch.zeros(ncols, K, dtype=torch.float64)
                    r_hi = min(rows, r0 + ncols)
                    k_hi = min(d.shape[4], K)
                    if r_hi > r0:
                        blk[:r_hi - r0, :k_hi] = d[0, 0, 0, r0:r_hi, :k_hi]
                    Bp.append(blk)
                acc += Ap[0] @ Bp[0].t()
                if job.planes == 2:
                    acc += Ap[1] @ Bp[0].t() + Ap[0] @ Bp[1].t()
            off = job.taps[g * job.taps_per_group][2]
            idx = (off + pix_off).reshape(-1, 1) + torch.arange(job.n_valid).view(1, -1)
            vals = (acc[..., :job.n_valid] * alpha).reshape(-1, job.n_valid).float()
            if getattr(job, "stats", None) is not None:
                assert job.splits == 1 and not accumulate
                job.stats[0, :job.n_valid] += vals.double().sum(0)
                job.stats[1, :job.n_valid] += (vals.double() ** 2).sum(0)
            if accumulate:
                out_flat[idx.reshape(-1)] += vals.reshape(-1)
            else:
                out_flat[idx.reshape(-1)] = vals.reshape(-1)
    else:
        assert job.block_n % 64 == 0 and job.ldc % 4 == 0
        Nr = -(-N // tn) * tn
        Hr = -(-H // th) * th
        Wr = -(-W // tw) * tw
        n_idx, h_idx, w_idx = torch.arange(Nr), torch.arange(Hr), torch.arange(Wr)
        m_cols = -(-job.m_valid // 128) * 128
        for g in range(job.groups):
            a4, b4, off = job.taps[g]
            acc = torch.zeros(m_cols, ncols, dtype=torch.float64)
            for pa, pb in ([(0, 0)] if job.planes == 1 else [(0, 0), (1, 0), (0, 1)]):
                Ap = _gather(A[pa], n_idx, h_idx + a4[3], a4[2], w_idx + a4[1], a4[0], m_cols)
                Bp = _gather(B[pb], n_idx, h_idx + b4[3], b4[2], w_idx + b4[1], b4[0], ncols)
                acc += Ap.reshape(-1, m_cols).t() @ Bp.reshape(-1, ncols)
            idx = off + torch.arange(job.m_valid).view(-1, 1) * job.ldc + torch.arange(job.n_valid).view(1, -1)
            vals = (acc[:job.m_valid, :job.n_valid] * alpha).float()
            if accumulate:
                out_flat[idx.reshape(-1)] += vals.reshape(-1)
            else:
                out_flat[idx.reshape(-1)] = vals.reshape(-1)


# ------------------------------------------------------------------------------ BN / packing
def _act(v, act):
    if act == 1:
        return torch.relu(v)
    if act == 2:
        return torch.where(v > 0, v, 0.2 * v)
    return v


def _act_grad(pre, act):
    if act == 1:
        return (pre > 0).to(pre.dtype)
    if act == 2:
        return torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, 0.2))
    return torch.ones_like(pre)


def _split16(v, dtype):
    t = TORCH16[dtype]
    hi = v.to(t)
    lo = (v - hi.float()).to(t)
    return hi, lo


def bn_workspace(rows, channels, device):
    return torch.zeros(2 * channels, device=device, dtype=torch.float64)


def bn_stats(x, stats):
    """ADDS into the caller-zeroed accumulator (include/cpcsv.h)"""
    Cc = x.shape[1]
    xd = x.double()
    stats[:Cc] += xd.sum(0)
    stats[Cc:2 * Cc] += (xd * xd).sum(0)


def bn_finalize(stats, rows, gamma, beta, running_mean, running_var, chan_map, c_valid,
                mean, invstd, scale, shift, eps=1e-5, momentum=0.1):
    Cc = mean.numel()
    m = stats[:Cc] / rows
    var = (stats[Cc:2 * Cc] / rows - m * m).clamp_min(0)
    inv = 1.0 / torch.sqrt(var + eps)
    idx = chan_map.long() if chan_map is not None else torch.arange(Cc)
    valid = (torch.arange(Cc) < c_valid) & (idx >= 0)
    pv = idx[valid]
    g = torch.zeros(Cc, dtype=torch.float64)
    b = torch.zeros(Cc, dtype=torch.float64)
    g[valid] = gamma[pv].double()
    b[valid] = beta[pv].double()
    mean.copy_(torch.where(valid, m, torch.zeros_like(m)).float())
    invstd.copy_(torch.where(valid, inv, torch.zeros_like(inv)).float())
    sc = (g.float() * inv.float())
    scale.copy_(torch.where(valid, sc, torch.zeros_like(sc)))
    shift.copy_(torch.where(valid, b.float() - m.float() * sc, torch.zeros_like(sc)))
    if running_mean is not None:
        unb = var * rows / (rows - 1.0) if rows > 1 else var
        running_mean[pv] = (1 - momentum) * running_mean[pv] + momentum * m[valid].float()
        running_var[pv] = (1 - momentum) * running_var[pv] + momentum * unb[valid].float()


def bn_act_pack(x, scale, shift, act, mod=None, y=None, hi=None, lo=None, dtype=1):
    assert x.shape[1] % 4 == 0
    v = x.float()
    if scale is not None:
        v = torch.addcmul(shift.view(1, -1), v, scale.view(1, -1))
    v = _act(v, act)
    if mod is not None:
        v = v * (1.0 + mod)
    if y is not None:
        y.copy_(v)
    if hi is not None:
        h, l = _split16(v, dtype)
        hi.copy_(h)
        if lo is not None:
            lo.copy_(l)


def _bwd_common(x, dy, scale, shift, mean, invstd, act, mod):
    pre = x.float()
    if scale is not None:
        pre = torch.addcmul(shift.view(1, -1), pre, scale.view(1, -1))
    a = _act(pre, act)
    da = dy * (1.0 + mod) if mod is not None else dy
    g = da * _act_grad(pre, act)
    xhat = (x - mean.view(1, -1)) * invstd.view(1, -1) if mean is not None else torch.zeros_like(x)
    return a, g, xhat


def bn_bwd_reduce(x, dy, scale, shift, mean, invstd, act, mod, sums):
    Cc = x.shape[1]
    _a, g, xhat = _bwd_common(x, dy, scale, shift, mean, invstd, act, mod)
    sums[:Cc] += g.double().sum(0)
    sums[Cc:2 * Cc] += (g.double() * xhat.double()).sum(0)


def bn_bwd_apply(x, dy, scale, shift, mean, invstd, chan_map, c_valid, act, mod, sums, has_bn,
                 dx=None, dx16=None, dmod=None, dmod16=None, dgamma=None, dbeta=None):
    rows, Cc = x.shape
    a, g, xhat = _bwd_common(x, dy, scale, shift, mean if has_bn else None, invstd, act, mod)
    if has_bn:
        mg = (sums[:Cc] / rows).float()
        mgx = (sums[Cc:2 * Cc] / rows).float()
        d = scale.view(1, -1) * (g - mg.view(1, -1) - xhat * mgx.view(1, -1))
    else:
        d = g
    if dx is not None:
        dx.copy_(d)
    if dx16 is not None:
        dx16.copy_(d.to(torch.bfloat16))
    if dmod is not None:
        dmod.copy_(dy * a)
    if dmod16 is not None:
        dmod16.copy_((dy * a).to(torch.bfloat16))
    if dgamma is not None and has_bn:
        idx = chan_map.long()[:c_valid] if chan_map is not None else torch.arange(c_valid)
        ok = idx >= 0
        dgamma[idx[ok]] = sums[Cc:Cc + c_valid].float()[ok]
        dbeta[idx[ok]] = sums[:c_valid].float()[ok]


def bn_norm_act_pack(x, stats, gamma, beta, running_mean, running_var, chan_map, c_valid, vec, act, mod=None,
                     y=None, hi=None, lo=None, dtype=1, eps=1e-5, momentum=0.1):
    v4 = vec.view(4, -1)
    bn_finalize(stats, x.shape[0], gamma, beta, running_mean, running_var, chan_map, c_valid,
                v4[0], v4[1], v4[2], v4[3], eps, momentum)
    bn_act_pack(x, v4[2], v4[3], act, mod, y, hi, lo, dtype)


# ------------------------------------------------------------------------------ layout kernels
def images_to_u8(x, out):
    arr = (x.float().clamp(-1.0, 1.0) + 1.0) / 2.0 * 255.0
    out.copy_(arr.permute(1, 2, 0).to(torch.uint8))


def pack_nchw(x, bcast, hi, lo, cpad, dtype=1):
    N, Cc, H, W = x.shape
    full = torch.zeros(N, H, W, cpad)
    full[..., :Cc] = x.permute(0, 2, 3, 1)
    if bcast is not None:
        cb = bcast.shape[1]
        full[..., Cc:Cc + cb] = bcast.view(N, 1, 1, cb)
    h, l = _split16(full, dtype)
    hi.reshape(N, H, W, cpad).copy_(h)
    if lo is not None:
        lo.reshape(N, H, W, cpad).copy_(l)


def im2col_small(x, k, s, p, hi, lo, ldp, dtype=1):
    N, Cc, H, W = x.shape
    cols = torch.nn.functional.unfold(x.float(), k, padding=p, stride=s)      # [N, C*k*k, L]
    L = cols.shape[-1]
    cols = cols.view(N, Cc, k * k, L).permute(0, 3, 2, 1).reshape(N * L, k * k * Cc)  # [(n,oh,ow), tap*C+c]
    full = torch.zeros(N * L, ldp)
    full[:, :k * k * Cc] = cols
    h, l = _split16(full, dtype)
    hi.reshape(N * L, ldp).copy_(h)
    if lo is not None:
        lo.reshape(N * L, ldp).copy_(l)


def enc0_lrelu_fwd(x, w, alpha, slope, hi, lo, ldp, dtype=1):
    N, Cc, H, W = x.shape
    Co = w.shape[0]
    assert Cc in (1, 3) and H % 4 == 0 and W % 64 == 0 and Co <= ldp <= 128 and ldp % 8 == 0
    z = torch.nn.functional.conv2d(x.float(), w.float(), stride=2, padding=1)
    if alpha is not None:
        z = z * alpha.float()
    a = torch.where(z > 0, z, slope * z).permute(0, 2, 3, 1)
    full = torch.zeros(N, H // 2, W // 2, ldp)
    full[..., :Co] = a
    h, l = _split16(full, dtype)
    hi.copy_(h.view_as(hi))
    if lo is not None:
        lo.copy_(l.view_as(lo))


def lrelu_bwd16(dy, a_hi, slope, dz):
    assert dy.numel() % 8 == 0
    dz.copy_((dy * torch.where(a_hi.float() > 0, 1.0, slope)).to(torch.bfloat16))


def col2im_small(dcol, N, Cc, H, W, k, s, p, dx):
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d = dcol[:, :k * k * Cc].reshape(N, OH * OW, k * k, Cc).permute(0, 3, 2, 1).reshape(N, Cc * k * k, OH * OW)
    dx.copy_(torch.nn.functional.fold(d, (H, W), k, padding=p, stride=s))


def head_gather_tanh(z, N, H, W, Co, y):
    zz = z[:, :9 * Co].reshape(N, H, W, 9, Co).double()
    pad = torch.nn.functional.pad(zz, (0, 0, 0, 0, 1, 1, 1, 1))          # pad W and H by 1
    acc = torch.zeros(N, H, W, Co, dtype=torch.float64)
    for ky in range(3):
        for kx in range(3):
            acc += pad[:, ky:ky + H, kx:kx + W, ky * 3 + kx]
    y.copy_(torch.tanh(acc).permute(0, 3, 1, 2))


def tanh_bwd_im2col(dy, y, col, dtype=1):
    N, Cc, H, W = y.shape
    dz = dy * (1.0 - y * y)
    # col[p, tap*C + c] = dz[p - delta_tap, c] with delta = (ky-1, kx-1): unfold of the
    # 180-degree-flipped neighbourhood
    pad = torch.nn.functional.pad(dz, (1, 1, 1, 1))
    out = torch.zeros(N, H, W, col.shape[1])
    for ky in range(3):
        for kx in range(3):
            sh = pad[:, :, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W]       # dz[h-(ky-1), w-(kx-1)]
            out[..., (ky * 3 + kx) * Cc:(ky * 3 + kx + 1) * Cc] = sh.permute(0, 2, 3, 1)
    col.copy_(out.reshape(N * H * W, -1).to(TORCH16[dtype]))


def pack_matrix(w, rows_out, cols_out, cols_valid, ld_r, ld_c, row_map, hi, lo, dtype=1, col_map=None):
    flat = w.reshape(-1)
    r = row_map.long() if row_map is not None else torch.arange(rows_out)
    full = torch.zeros(rows_out, cols_out)
    cidx = torch.arange(cols_valid)
    csrc = col_map.long()[:cols_valid] if col_map is not None else cidx
    ok = r >= 0
    cok = csrc >= 0
    idx = r[ok].view(-1, 1) * ld_r + csrc[cok].view(1, -1) * ld_c
    full[ok.nonzero().view(-1, 1), cidx[cok].view(1, -1)] = flat[idx]
    h, l = _split16(full, dtype)
    hi[:rows_out, :cols_out].copy_(h)
    if lo is not None:
        lo[:rows_out, :cols_out].copy_(l)


def scatter_rows_f32(src, row_map, dst, rows, cols):
    r = row_map.long()[:rows] if row_map is not None else torch.arange(rows)
    ok = r >= 0
    dst[r[ok], :cols] = src[:rows, :cols][ok]


_MERGE = {(0, 0): (0, 0), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2, 2)}


def _merged_taps(w):
    """w [Cout,Cin,3,3] -> [16, Cout, Cin] in (a,b,i,j) order."""
    out = []
    for a in range(2):
        for b in range(2):
            for i in range(2):
                for j in range(2):
                    y0, y1 = _MERGE[(a, i)]
                    x0, x1 = _MERGE[(b, j)]
                    out.append(w[:, :, y0:y1 + 1, x0:x1 + 1].sum((2, 3)))
    return torch.stack(out, 0)


def pack_conv_weight(w, kind, rows_pad, cols_pad, hi, lo, dtype=1):
    Cout, Cin, kh, kw = w.shape
    taps = _merged_taps(w) if kind >= 2 else w.reshape(Cout, Cin, kh * kw).permute(2, 0, 1)
    if kind in (1, 3):
        taps = taps.transpose(1, 2)
    T, R, Cc = taps.shape
    full = torch.zeros(T, rows_pad, cols_pad)
    full[:, :R, :Cc] = taps
    h, l = _split16(full, dtype)
    hi.reshape(T, rows_pad, cols_pad).copy_(h)
    if lo is not None:
        lo.reshape(T, rows_pad, cols_pad).copy_(l)


def unpack_conv_wgrad(dwt, mat_stride, ldc, kind, alpha, dw):
    Cout, Cin, kh, kw = dw.shape
    T = 16 if kind >= 2 else kh * kw
    R, Cc = (Cin, Cout) if kind in (1, 3) else (Cout, Cin)
    flat = dwt.reshape(-1)
    idx = (torch.arange(T).view(-1, 1, 1) * mat_stride + torch.arange(R).view(1, -1, 1) * ldc
           + torch.arange(Cc).view(1, 1, -1))
    mats = flat[idx]
    if kind in (1, 3):
        mats = mats.transpose(1, 2)                   # [T, Cout, Cin]
    if kind < 2:
        res = mats.permute(1, 2, 0).reshape(Cout, Cin, kh, kw)
    else:
        res = torch.zeros(Cout, Cin, 3, 3)
        t = 0
        for a in range(2):
            for b in range(2):
                for i in range(2):
                    for j in range(2):
                        y0, y1 = _MERGE[(a, i)]
                        x0, x1 = _MERGE[(b, j)]
                        res[:, :, y0:y1 + 1, x0:x1 + 1] += mats[t].view(Cout, Cin, 1, 1)
                        t += 1
    dw.copy_(res * (float(alpha) if alpha is not None else 1.0))


# ------------------------------------------------------------------------------ fp32 small ops
def linear_f32(x, w, bias, y, accumulate=False):
    r = x @ w.t()
    if bias is not None:
        r = r + bias
    y.copy_(y + r if accumulate else r)


def linear_tn_f32(a, b, y, accumulate=False):
    r = a.t() @ b
    y.copy_(y + r if accumulate else r)


def linear_nn_f32(x, w, y, accumulate=False):
    r = x @ w
    y.copy_(y + r if accumulate else r)


def gru_gates_fwd(gi, gh, h, hnew, save):
    H = h.shape[1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    hn = gh[:, 2 * H:]
    n = torch.tanh(gi[:, 2 * H:] + r * hn)
    hnew.copy_((1 - z) * n + z * h)
    save.copy_(torch.cat((r, z, n, hn), 1))


def gru_gates_bwd(dhnew, h, save, dgi, dgh, dh):
    H = h.shape[1]
    r, z, n, hn = save[:, :H], save[:, H:2 * H], save[:, 2 * H:3 * H], save[:, 3 * H:]
    dn = dhnew * (1 - z)
    dz = dhnew * (h - n)
    dpn = dn * (1 - n * n)
    dpr = dpn * hn * r * (1 - r)
    dpz = dz * z * (1 - z)
    dgi.copy_(torch.cat((dpr, dpz, dpn), 1))
    dgh.copy_(torch.cat((dpr, dpz, dpn * r), 1))
    dh.copy_(dhnew * z)


def ca_fwd(pre, eps, mu, logvar, code):
    Cc = mu.shape[1]
    x = torch.relu(pre)
    mu.copy_(x[:, :Cc])
    logvar.copy_(x[:, Cc:])
    code.copy_(eps * torch.exp(0.5 * x[:, Cc:]) + x[:, :Cc])


def ca_bwd(pre, eps, dmu, dlogvar, dcode, dpre):
    Cc = eps.shape[1]
    lv = torch.relu(pre[:, Cc:])
    z = torch.zeros_like(eps)
    dc = dcode if dcode is not None else z
    gm = (dmu if dmu is not None else z) + dc
    gl = (dlogvar if dlogvar is not None else z) + dc * eps * 0.5 * torch.exp(0.5 * lv)
    dpre[:, :Cc] = gm * (pre[:, :Cc] > 0)
    dpre[:, Cc:] = gl * (pre[:, Cc:] > 0)


def dfn1d_fwd(img, filt, out):
    N, Cc, L = img.shape
    K = filt.shape[-1]
    r = torch.nn.functional.conv1d(img.reshape(1, N * Cc, L), filt.reshape(N, Cc, K), padding=K // 2, groups=N)
    out.copy_(r.reshape(out.shape))


def dfn1d_bwd(img, filt, dout, dimg, dfilt):
    N, Cc, L = img.shape
    K = filt.shape[-1]
    with torch.enable_grad():
        i2 = img.detach().clone().requires_grad_(True)
        f2 = filt.detach().clone().reshape(N, Cc, K).requires_grad_(True)
        r = torch.nn.functional.conv1d(i2.reshape(1, N * Cc, L), f2, padding=K // 2, groups=N)
        gi, gf = torch.autograd.grad(r, (i2, f2), dout.reshape(r.shape))
    dimg.copy_(gi)
    dfilt.copy_(gf.reshape(dfilt.shape))


def tanh_fwd(x, y):
    y.copy_(torch.tanh(x))


def tanh_bwd(y, dy, dx):
    dx.copy_(dy * (1 - y * y))


def affine_sigmoid_fwd(t, alpha, bias, out):
    a = float(alpha) if alpha is not None else 1.0
    b = float(bias) if bias is not None else 0.0
    out.copy_(torch.sigmoid(t * a + b))


def affine_sigmoid_bwd(dout, out, alpha, dt, dz):
    a = float(alpha) if alpha is not None else 1.0
    g = dout * out * (1 - out)
    dz.copy_(g)
    dt.copy_(g * a)


def spectral_sigma(w2d, u, v, power_iteration, sigma, inv_sigma, scratch, eps=1e-12):
    if power_iteration:
        v.copy_(torch.nn.functional.normalize(torch.mv(w2d.t(), u), dim=0, eps=eps))
        u.copy_(torch.nn.functional.normalize(torch.mv(w2d, v), dim=0, eps=eps))
    s = torch.dot(u, torch.mv(w2d, v))
    sigma.copy_(s.reshape(sigma.shape))
    inv_sigma.copy_((1.0 / s).reshape(inv_sigma.shape))


def spectral_bwd(g2d, w2d, u, v, sigma, dw2d, scratch):
    s = float(sigma)
    gw = float((g2d * w2d).sum())
    dw2d.copy_((g2d - (gw / s) * torch.outer(u, v)) / s)


# ------------------------------------------------------------------------------ optimiser
def adam_tick(step, beta1, beta2, bc):
    step += 1
    t = float(step)
    bc[0] = 1.0 / (1.0 - beta1 ** t)
    bc[1] = 1.0 / (1.0 - beta2 ** t) ** 0.5


def _adam(p, g, m, v, hyper):
    m.add_((g - m) * (1.0 - hyper.beta1))
    v.mul_(hyper.beta2).add_(g * g * (1.0 - hyper.beta2))
    step_size = float(hyper.lr) * float(hyper.bc[0])
    denom = v.sqrt() * float(hyper.bc[1]) + hyper.eps
    p.sub_(step_size * (m / denom))


def adam_multi(tensors, hyper):
    with torch.no_grad():
        for p, g, m, v in tensors:
            _adam(p, g, m, v, hyper)


def adam_pack_conv(w, g, m, v, planes, hyper=None):
    with torch.no_grad():
        if g is not None:
            _adam(w, g, m, v, hyper)
        for kind, dtype, rows_pad, cols_pad, hi, lo in planes:
            pack_conv_weight(w.detach(), kind, rows_pad, cols_pad, hi, lo, dtype)


def adam_pack_fc(w, g, m, v, C_, P, Cp, Kp, fwd16=None, fwd_hi=None, fwd_lo=None, bwd=None, hyper=None):
    with torch.no_grad():
        if g is not None:
            _adam(w, g, m, v, hyper)
        K = w.shape[1]
        full = torch.zeros(P, Cp, Kp)
        full[:, :C_, :K] = w.detach().reshape(C_, P, K).permute(1, 0, 2)        # [pos, c, k]
        full = full.reshape(P * Cp, Kp)
        if fwd16 is not None:
            fwd16.copy_(full.to(torch.float16))
        if fwd_hi is not None:
            h, l = _split16(full, 1)
            fwd_hi.copy_(h)
            if fwd_lo is not None:
                fwd_lo.copy_(l)
        if bwd is not None:
            bwd.copy_(full.t().to(torch.bfloat16))


def unpack_conv_wgrad_dot(dwt, mat_stride, ldc, kind, alpha, dw, w, dot):
    unpack_conv_wgrad(dwt, mat_stride, ldc, kind, alpha, dw)
    dot[0] += float((dw.double() * w.double()).sum())


def spectral_bwd_apply(g2d, u, v, sigma, dot, dw2d):
    s = float(sigma)
    dw2d.copy_((g2d - (float(dot[0]) / s) * torch.outer(u, v)) / s)
