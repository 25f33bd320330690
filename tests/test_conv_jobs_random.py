"""Randomised shapes (hypothesis) for the GEMM job builders of cpcsv_b200/conv.py on the CPU emulator
of the kernel contract: ragged batch sizes (pixel boxes that run past N), every legal map size of
the step (4..32), 64 / 128 / 192 channels, 1 or 2 operand planes, forward + data gradient + weight
gradient against torch.nn.functional in fp64.  The fixed-shape cases of tests/test_conv_jobs.py are
the ones that also run on the GPU."""
import pytest
import torch
import torch.nn.functional as F

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

import emulator  # noqa: E402
from cpcsv_b200 import conv, ops  # noqa: E402
from test_conv_jobs import bf16r, eff_weight, nhwc, pack_w, rel, rnd, split  # noqa: E402


class _Patch:
    def __init__(self):
        self.saved = []

    def setattr(self, obj, name, value):
        self.saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def undo(self):
        for obj, name, old in reversed(self.saved):
            setattr(obj, name, old)


@pytest.fixture
def emu():
    p = _Patch()
    emulator.install(p)
    yield torch.device("cpu")
    p.undo()


SHAPES = st.tuples(st.integers(1, 7), st.sampled_from([4, 8, 16]), st.sampled_from([64, 128, 192]),
                   st.sampled_from([64, 128, 192]), st.sampled_from([1, 2]), st.sampled_from(["s1", "up", "s2"]))


@settings(max_examples=18, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture])     # the fixture only installs the emulator
@given(SHAPES)
def test_random_forward_and_backward(emu, shape):
    N, H, Ci, Co, planes, kind = shape
    dev = emu
    k = 4 if kind == "s2" else 3
    x = rnd(N, Ci, H, H, seed=21, dev=dev)
    w = rnd(Co, Ci, k, k, seed=22, dev=dev, scale=0.05)
    xp, x_eff = split(nhwc(x), planes)
    x_eff = x_eff.permute(0, 3, 1, 2).double()
    if kind == "s1":
        Ho, fwd = H, lambda a, b: F.conv2d(a, b, padding=1)
        out = torch.zeros(N, Ho, Ho, Co, device=dev)
        ops.conv_gemm(conv.conv_s1_fwd(xp, pack_w(w, 0, Co, Ci, planes, dev), out))
        tol_f = 3e-5 if planes == 2 else 1e-5
        ref = fwd(x_eff, eff_weight(w, planes).double())
    elif kind == "up":
        Ho, fwd = 2 * H, lambda a, b: F.conv2d(F.interpolate(a, scale_factor=2, mode="nearest"), b, padding=1)
        out = torch.zeros(N, Ho, Ho, Co, device=dev)
        ops.conv_gemm(conv.upconv_fwd(xp, pack_w(w, 2, Co, Ci, planes, dev), out))
        tol_f = 3e-5 if planes == 2 else 4e-3          # merged taps are rounded after merging
        ref = fwd(x_eff, w.double())
    else:
        Ho, fwd = H // 2, lambda a, b: F.conv2d(a, b, stride=2, padding=1)
        out = torch.zeros(N, Ho, Ho, Co, device=dev)
        ops.conv_gemm(conv.conv_s2_fwd(xp, pack_w(w, 0, Co, Ci, planes, dev), out))
        tol_f = 3e-5 if planes == 2 else 1e-5
        ref = fwd(x_eff, eff_weight(w, planes).double())
    assert rel(out.permute(0, 3, 1, 2), ref) < tol_f, shape
    # backward (single-pass bf16 operands)
    xb = bf16r(x)
    wb = w if kind == "up" else bf16r(w)
    dy = bf16r(rnd(N, Co, Ho, Ho, seed=23, dev=dev))
    xd, wd = xb.double().requires_grad_(True), wb.double().requires_grad_(True)
    gx, gw = torch.autograd.grad(fwd(xd, wd), (xd, wd), dy.double())
    dy16 = nhwc(dy).to(torch.bfloat16)
    x16 = nhwc(xb).to(torch.bfloat16)
    dx = torch.zeros(N, H, H, Ci, device=dev)
    ntap = {"s1": 9, "up": 16, "s2": 16}[kind]
    dwt = torch.zeros(ntap, Co, Ci, device=dev)
    dw = torch.empty(Co, Ci, k, k, device=dev)
    if kind == "s1":
        ops.conv_gemm(conv.conv_s1_dgrad(dy16, pack_w(wb, 1, Ci, Co, 1, dev)[0], dx))
        ops.conv_gemm(conv.conv_s1_wgrad(dy16, x16, dwt))
        ukind, tol_x = 0, 1e-5
    elif kind == "up":
        ops.conv_gemm(conv.upconv_dgrad(dy16, pack_w(wb, 3, Ci, Co, 1, dev)[0], dx))
        ops.conv_gemm(conv.upconv_wgrad(dy16, x16, dwt))
        ukind, tol_x = 2, 4e-3
    else:
        ops.conv_gemm(conv.conv_s2_dgrad(dy16, pack_w(wb, 1, Ci, Co, 1, dev)[0], dx))
        ops.conv_gemm(conv.conv_s2_wgrad(dy16, x16, dwt))
        ukind, tol_x = 0, 1e-5
    ops.unpack_conv_wgrad(dwt, Co * Ci, Ci, ukind, None, dw)
    assert rel(dx.permute(0, 3, 1, 2), gx) < tol_x, shape
    assert rel(dw, gw) < 1e-5, shape
