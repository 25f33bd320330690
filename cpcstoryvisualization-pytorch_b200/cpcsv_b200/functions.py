"""torch.autograd.Function wrappers of the fp32 conditioning-path kernels (CA_NET, GRU cells,
Linear + BatchNorm1d, dynamic 1-D filter; reference model.py:37-65, 223-224, 250-257,
302-346, layers.py:62-80).  These layers stay fp32.  Every
forward/backward below is one or a few libcpcsv.so launches; torch only allocates.
"""
import torch

from . import ops
from .engine import SpectralNorm, StateOrder


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _rowmat(t):
    """2-D tensor with unit column stride (row pitch free)."""
    if t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1) and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


class LinearFn(torch.autograd.Function):
    """y = x W^T + b  (fp32)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _rowmat(x), _rowmat(w)
        y = torch.empty(x.shape[0], w.shape[0], device=x.device)
        ops.linear_f32(x, w, b, y)
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _rowmat(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x, memory_format=torch.contiguous_format)
            ops.linear_nn_f32(dy, w, dx)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(w.shape, device=w.device)
            ops.linear_tn_f32(dy, x, dw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            ones = torch.ones(dy.shape[0], 1, device=dy.device)
            db = torch.empty(dy.shape[1], 1, device=dy.device)
            ops.linear_tn_f32(dy, ones, db)
            db = db.view(-1)
        return dx, dw, db


class BatchNorm1dFn(torch.autograd.Function):
    """BatchNorm1d with batch statistics (+ optional tanh), running stats updated in place."""

    @staticmethod
    def forward(ctx, x, gamma, beta, rmean, rvar, nbt, training, act_tanh):
        M, C = x.shape
        Cp = (C + 3) // 4 * 4
        dev = x.device
        xp = torch.zeros(M, Cp, device=dev)
        xp[:, :C].copy_(x)
        stats = ops.bn_workspace(M, Cp, dev)
        vec = torch.empty(4, Cp, device=dev)
        yp = torch.empty(M, Cp, device=dev)
        if not training:
            # eval mode: running statistics, no state update (reference inference.py:88 netG.eval())
            sc = gamma.detach() / torch.sqrt(rvar + 1e-5)
            scale = torch.zeros(Cp, device=dev)
            shift = torch.zeros(Cp, device=dev)
            scale[:C].copy_(sc)
            shift[:C].copy_(beta.detach() - rmean * sc)
            ops.bn_act_pack(xp, scale, shift, ops.ACT_NONE, y=yp)
            out = yp
            if act_tanh:
                out = torch.empty_like(yp)
                ops.tanh_fwd(yp, out)
            ctx.eval_mode = True
            return out[:, :C]
        ctx.eval_mode = False
        rm, rv = (rmean, rvar) if training else (None, None)
        StateOrder.before(rm)
        ops.bn_stats(xp, stats)
        ops.bn_norm_act_pack(xp, stats, gamma.detach(), beta.detach(), rm, rv, None, C, vec, ops.ACT_NONE, y=yp)
        if training and nbt is not None:
            nbt.add_(1)
        StateOrder.after(rm)
        out = yp
        if act_tanh:
            out = torch.empty_like(yp)
            ops.tanh_fwd(yp, out)
        ctx.save_for_backward(xp, vec, out if act_tanh else None)
        ctx.C, ctx.act_tanh = C, act_tanh
        ctx.gshape = gamma.shape
        return out[:, :C]

    @staticmethod
    def backward(ctx, dy):
        if ctx.eval_mode:
            raise NotImplementedError("cpcsv_b200: backward through eval-mode BatchNorm1d is not implemented "
                                      "(the reference only evaluates under torch.no_grad())")
        xp, vec, tout = ctx.saved_tensors
        M, Cp = xp.shape
        C = ctx.C
        dev = xp.device
        dyp = torch.zeros(M, Cp, device=dev)
        dyp[:, :C].copy_(dy)
        if ctx.act_tanh:
            d2 = torch.empty_like(dyp)
            ops.tanh_bwd(tout, dyp, d2)
            dyp = d2
        sums = ops.bn_workspace(M, Cp, dev)
        ops.bn_bwd_reduce(xp, dyp, vec[2], vec[3], vec[0], vec[1], ops.ACT_NONE, None, sums)
        dx = torch.empty(M, Cp, device=dev)
        dgamma = torch.zeros(ctx.gshape, device=dev)
        dbeta = torch.zeros(ctx.gshape, device=dev)
        ops.bn_bwd_apply(xp, dyp, vec[2], vec[3], vec[0], vec[1], None, C, ops.ACT_NONE, None, sums, True,
                         dx=dx, dgamma=dgamma, dbeta=dbeta)
        return dx[:, :C], dgamma, dbeta, None, None, None, None, None


class GRUSeqFn(torch.autograd.Function):
    """T steps of torch.nn.GRUCell (gate order r, z, n) over a known input sequence.

    x_all [T, B, I], h0 [B, H] -> h_all [T, B, H].  The input projections of all steps are ONE
    GEMM (they do not depend on the recurrence); each step is then one recurrent GEMM + one gate
    kernel.  Backward: per step one gate kernel + one GEMM for the carried dh; the weight / bias /
    input gradients are single GEMMs over all T*B rows after the loop."""

    @staticmethod
    def forward(ctx, x_all, h0, w_ih, w_hh, b_ih, b_hh):
        T, B, I = x_all.shape
        H = h0.shape[1]
        dev = x_all.device
        x2 = _c(x_all).view(T * B, I)
        gi = torch.empty(T * B, 3 * H, device=dev)
        ops.linear_f32(x2, w_ih, b_ih, gi)
        # rows [0, B) of h_in hold h0, rows of step t+1 hold h_t: h_in[:T*B] are the step inputs
        h_in = torch.empty((T + 1) * B, H, device=dev)
        h_in[:B].copy_(h0)
        save = torch.empty(T * B, 4 * H, device=dev)
        gh = torch.empty(B, 3 * H, device=dev)
        for t in range(T):
            hp = h_in[t * B:(t + 1) * B]
            ops.linear_f32(hp, w_hh, b_hh, gh)
            ops.gru_gates_fwd(gi[t * B:(t + 1) * B], gh, hp, h_in[(t + 1) * B:(t + 2) * B],
                              save[t * B:(t + 1) * B])
        ctx.save_for_backward(x2, h_in, w_ih, w_hh, save)
        ctx.dims = (T, B, I, H)
        return h_in[B:].view(T, B, H)

    @staticmethod
    def backward(ctx, dh_all):
        x2, h_in, w_ih, w_hh, save = ctx.saved_tensors
        T, B, I, H = ctx.dims
        dev = x2.device
        dh_all = _c(dh_all)
        dgi = torch.empty(T * B, 3 * H, device=dev)
        dgh = torch.empty(T * B, 3 * H, device=dev)
        dh = None          # gradient carried into h_{t-1}
        for t in range(T - 1, -1, -1):
            dhn = dh_all[t] if dh is None else dh_all[t] + dh
            dh = torch.empty(B, H, device=dev)
            ops.gru_gates_bwd(dhn, h_in[t * B:(t + 1) * B], save[t * B:(t + 1) * B],
                              dgi[t * B:(t + 1) * B], dgh[t * B:(t + 1) * B], dh)
            if t > 0 or ctx.needs_input_grad[1]:
                ops.linear_nn_f32(dgh[t * B:(t + 1) * B], w_hh, dh, accumulate=True)
        dx = dwi = dwh = dbi = dbh = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(T * B, I, device=dev)
            ops.linear_nn_f32(dgi, w_ih, dx)
            dx = dx.view(T, B, I)
        if ctx.needs_input_grad[2]:
            dwi = torch.empty(w_ih.shape, device=dev)
            ops.linear_tn_f32(dgi, x2, dwi)
        if ctx.needs_input_grad[3]:
            dwh = torch.empty(w_hh.shape, device=dev)
            ops.linear_tn_f32(dgh, h_in[:T * B], dwh)
        if ctx.needs_input_grad[4] or ctx.needs_input_grad[5]:
            ones = torch.ones(T * B, 1, device=dev)
            if ctx.needs_input_grad[4]:
                dbi = torch.empty(3 * H, 1, device=dev)
                ops.linear_tn_f32(dgi, ones, dbi)
                dbi = dbi.view(-1)
            if ctx.needs_input_grad[5]:
                dbh = torch.empty(3 * H, 1, device=dev)
                ops.linear_tn_f32(dgh, ones, dbh)
                dbh = dbh.view(-1)
        return dx, (dh if ctx.needs_input_grad[1] else None), dwi, dwh, dbi, dbh


class CondAugFn(torch.autograd.Function):
    """CA_NET tail: relu, split into (mu, logvar), reparameterise with the given eps."""

    @staticmethod
    def forward(ctx, pre, eps):
        pre, eps = _c(pre), _c(eps)
        B, C = eps.shape
        mu, logvar, code = (torch.empty(B, C, device=pre.device) for _ in range(3))
        ops.ca_fwd(pre, eps, mu, logvar, code)
        ctx.save_for_backward(pre, eps)
        return mu, logvar, code

    @staticmethod
    def backward(ctx, dmu, dlogvar, dcode):
        pre, eps = ctx.saved_tensors
        dpre = torch.empty_like(pre)
        ops.ca_bwd(pre, eps, _c(dmu) if dmu is not None else None,
                   _c(dlogvar) if dlogvar is not None else None,
                   _c(dcode) if dcode is not None else None, dpre)
        return dpre, None


class DynamicFilter1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, filt):
        img, filt = _c(img), _c(filt)
        N, C, L = img.shape
        out = torch.empty(N, 1, L, device=img.device)
        ops.dfn1d_fwd(img, filt, out)
        ctx.save_for_backward(img, filt)
        return out

    @staticmethod
    def backward(ctx, dout):
        img, filt = ctx.saved_tensors
        dimg, dfilt = torch.empty_like(img), torch.empty_like(filt)
        ops.dfn1d_bwd(img, filt, _c(dout), dimg, dfilt)
        return dimg, dfilt


class SpectralWeightFn(torch.autograd.Function):
    """weight = weight_orig / sigma of the legacy torch spectral_norm hook (one power iteration per training-mode
    call, u / v updated in place) for a Linear that runs as an fp32 kernel (the detector of the VideoEncoder,
    reference model.py:192-197)"""

    @staticmethod
    def forward(ctx, w_orig, u, v, training):
        sn = SpectralNorm(u, v)
        inv_sigma = sn.forward(w_orig, training, True)
        ctx.sn = sn
        ctx.save_for_backward(w_orig)
        return w_orig.detach() * inv_sigma

    @staticmethod
    def backward(ctx, g):
        (w_orig,) = ctx.saved_tensors
        sn, ctx.sn = ctx.sn, None
        return sn.backward(_c(g), w_orig), None, None, None


def spectral_weight(mod):
    """mod: holder module carrying weight_orig / weight_u / weight_v"""
    return SpectralWeightFn.apply(mod.weight_orig, mod.weight_u, mod.weight_v, mod.training)


def linear(x, w, b=None):
    return LinearFn.apply(x, w, b)


def batch_norm_1d(x, bn, act_tanh=False):
    """bn: nn.BatchNorm1d holder module."""
    return BatchNorm1dFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                               bn.num_batches_tracked, bn.training, act_tanh)


def gru_sequence(x_all, h0, cell):
    """cell: nn.GRUCell holder module; x_all [T, B, I], h0 [B, H] -> h_all [T, B, H]."""
    return GRUSeqFn.apply(x_all, h0, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)


def gru_cell(x, h, cell):
    """one step (torch.nn.GRUCell semantics)"""
    return gru_sequence(x.unsqueeze(0), h, cell)[0]
