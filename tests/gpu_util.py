"""Helpers for the -m gpu parity tests: thin ctypes callers of the single-operator entry points."""
import ctypes as C

import torch

from vidtok_b200 import _native as N


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_cl(x, dtype):
    """[B,C,T,H,W] cpu -> channels-last [B,T,H,W,C] cuda"""
    return x.permute(0, 2, 3, 4, 1).contiguous().to("cuda", dtype)


def from_cl(y):
    return y.float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def op_conv(x, w, b, *, stride=(1, 1, 1), pt=None, pads=None, up=(1, 1, 1), res=None, res_mode=0, alpha=0.0,
            precision=N.PREC_EXACT, force_simt=False):
    """x [B,Ci,T,H,W], w [Co,Ci,kt,kh,kw] (cpu fp32) -> [B,Co,To,Ho,Wo] cpu fp32 via vt_op_conv."""
    dt = torch.float32 if precision == N.PREC_EXACT else torch.bfloat16
    B, Ci, T, H, W = x.shape
    Co, _, kt, kh, kw = w.shape
    d = N.ConvDesc()
    d.B, d.Ti, d.Hi, d.Wi, d.Ci, d.Co = B, T, H, W, Ci, Co
    d.kt, d.kh, d.kw = kt, kh, kw
    d.st, d.sh, d.sw = stride
    d.pt = (kt - 1) + (1 - stride[0]) if pt is None else pt
    if pads is None:
        hp, wp = (kh - 1) + (1 - stride[1]), (kw - 1) + (1 - stride[2])
        pads = (hp // 2, hp - hp // 2, wp // 2, wp - wp // 2)
    d.ph0, d.ph1, d.pw0, d.pw1 = pads
    d.ut, d.uh, d.uw = up
    d.res_mode, d.alpha = res_mode, alpha
    To = (up[0] * T + d.pt - kt) // stride[0] + 1
    Ho = (up[1] * H + pads[0] + pads[1] - kh) // stride[1] + 1
    Wo = (up[2] * W + pads[2] + pads[3] - kw) // stride[2] + 1
    xc = to_cl(x, dt)
    wd, bd = w.contiguous().cuda(), b.contiguous().cuda()
    rc_ = to_cl(res, dt) if res is not None else None
    out = torch.empty((B, To, Ho, Wo, Co), dtype=dt, device="cuda")
    N.check(N.lib().vt_op_conv(precision, int(force_simt), C.byref(d), _p(xc), _p(wd), _p(bd), _p(rc_), _p(out), stream()))
    torch.cuda.synchronize()
    return from_cl(out)
