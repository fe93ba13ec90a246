"""CPU emulation of tensor-core operand-rounding schemes on the oracle UNet (SURVEY.md section 7 probe,
repeated for cheaper-than-bf16x3 candidates).  Usage: python tools/experiments/precision_emul.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as RF
import oracle.unet_oracle as uo

torch.set_num_threads(os.cpu_count())
E4 = torch.float8_e4m3fn

def bf16(x): return x.to(torch.bfloat16).float()
def fp16(x): return x.to(torch.float16).float()
def e4m3(x): return x.clamp(-448, 448).to(E4).float()

STATS = {}

def make(scheme):
    def prod(op, a, w):
        """op(a, w) -> linear map without bias"""
        if scheme == "fp32":
            return op(a, w)
        if scheme == "bf16x3":
            ah, wh = bf16(a), bf16(w)
            al, wl = bf16(a - ah), bf16(w - wh)
            return op(ah, wh) + op(al, wh) + op(ah, wl)
        if scheme == "fp16x1":
            return op(fp16(a), fp16(w))
        if scheme == "fp16_2a":  # (a_hi + a_lo) * w_hi
            ah = fp16(a); al = fp16(a - ah)
            return op(ah, fp16(w)) + op(al, fp16(w))
        if scheme.startswith("fp16_f8"):
            # hi*hi in fp16; cross terms in e4m3 with power-of-two scales
            ka_hi, kw_hi = 0, 6
            ah, wh = fp16(a), fp16(w)
            al, wl = a - ah, w - wh
            sa_hi, sw_hi = 2.0 ** ka_hi, 2.0 ** kw_hi
            sa_lo, sw_lo = sa_hi * 2048, sw_hi * 2048
            STATS["amax"] = max(STATS.get("amax", 0.0), float(a.abs().max()))
            STATS["wmax"] = max(STATS.get("wmax", 0.0), float(w.abs().max()))
            al8, ah8 = e4m3(al * sa_lo), e4m3(ah * sa_hi)
            wl8, wh8 = e4m3(wl * sw_lo), e4m3(wh * sw_hi)
            cross = op(al8, wh8) + op(ah8, wl8)
            if scheme == "fp16_f8_1":  # only the activation-lo cross term
                cross = op(al8, wh8)
            return op(ah, wh) + cross / (sa_lo * sw_hi)
        raise ValueError(scheme)
    return prod

class FProxy:
    """torch.nn.functional with conv2d / linear replaced by operand-rounded versions"""
    def __init__(self, conv_scheme, lin_scheme):
        self.cp, self.lp = make(conv_scheme), make(lin_scheme)
    def __getattr__(self, k): return getattr(RF, k)
    def conv2d(self, x, w, b=None, stride=1, padding=0):
        if w.shape[1] < 16 or w.shape[0] < 16:  # first / last conv: fp32 kernels
            return RF.conv2d(x, w, b, stride=stride, padding=padding)
        y = self.cp(lambda a, ww: RF.conv2d(a, ww, None, stride=stride, padding=padding), x, w)
        return y if b is None else y + b[None, :, None, None]
    def linear(self, x, w, b=None):
        if x.dim() == 2:  # time-embedding MLP / emb projections: fp32 kernels
            return RF.linear(x, w, b)
        y = self.lp(lambda a, ww: RF.linear(a, ww), x, w)
        return y if b is None else y + b

def run(conv_scheme, lin_scheme, sd, cfg, x, t, c):
    uo.F = FProxy(conv_scheme, lin_scheme)
    try:
        return uo.unet_forward(sd, cfg, x, t, c)
    finally:
        uo.F = RF

if __name__ == "__main__" and len(sys.argv) == 1:
    from polyffusion_b200.stable_diffusion.model.unet import UNetModel
    torch.manual_seed(0)
    kw = dict(in_channels=2, out_channels=2, channels=64, n_res_blocks=2, attention_levels=[2, 3],
              channel_multipliers=[1, 2, 4, 4], n_heads=4, tf_layers=1, d_cond=512)
    sd = UNetModel(**kw).state_dict()
    cfg = uo.UNetCfg(d_cond=512)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 128, 128, generator=g)
    c = torch.randn(2, 1, 512, generator=g)
    t = torch.tensor([999, 3])
    ref = run("fp32", "fp32", sd, cfg, x, t, c).double()
    print("ref abs mean", float(ref.abs().mean()), "max", float(ref.abs().max()))
    for cs, ls in [("bf16x3", "bf16x3"), ("fp16_f8", "bf16x3"), ("fp16_f8", "fp16_f8"), ("fp16_f8_1", "bf16x3"),
                   ("fp16_2a", "bf16x3"), ("fp16x1", "bf16x3")]:
        STATS.clear()
        y = run(cs, ls, sd, cfg, x, t, c).double()
        d = (y - ref).abs()
        ok = (d <= 1e-4 + 1e-3 * ref.abs()).double().mean()
        print(f"conv={cs:10s} lin={ls:8s} max {d.max():.3e} rms {d.pow(2).mean().sqrt():.3e} within-tol {100*ok:.3f}%  {STATS}")


def sensitivity():
    """which convolutions contribute most of the f16f8 error at the output?"""
    from polyffusion_b200.stable_diffusion.model.unet import UNetModel
    torch.manual_seed(0)
    kw = dict(in_channels=2, out_channels=2, channels=64, n_res_blocks=2, attention_levels=[2, 3],
              channel_multipliers=[1, 2, 4, 4], n_heads=4, tf_layers=1, d_cond=512)
    sd = UNetModel(**kw).state_dict()
    cfg = uo.UNetCfg(d_cond=512)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 128, 128, generator=g)
    c = torch.randn(2, 1, 512, generator=g)
    t = torch.tensor([999, 3])
    ref = run("fp32", "fp32", sd, cfg, x, t, c).double()

    class Sel(FProxy):
        def __init__(self, pred):
            super().__init__("fp16_f8", "bf16x3")
            self.alt = make("bf16x3")
            self.pred = pred
            self.n = 0
        def conv2d(self, x, w, b=None, stride=1, padding=0):
            if w.shape[1] < 16 or w.shape[0] < 16:
                return RF.conv2d(x, w, b, stride=stride, padding=padding)
            i = self.n
            self.n += 1
            p = self.cp if self.pred(i, x, w) else self.alt
            y = p(lambda a, ww: RF.conv2d(a, ww, None, stride=stride, padding=padding), x, w)
            return y if b is None else y + b[None, :, None, None]

    def go(name, pred):
        px = Sel(pred)
        uo.F = px
        try:
            y = uo.unet_forward(sd, cfg, x, t, c).double()
        finally:
            uo.F = RF
        d = (y - ref).abs()
        print(f"{name:40s} convs={px.n} max {d.max():.3e} rms {d.pow(2).mean().sqrt():.3e}")
        return px.n

    n = go("all f8", lambda i, x, w: True)
    go("none f8 (bf16x3)", lambda i, x, w: False)
    for k in (2, 4, 8, 16):
        go(f"f8 except last {k} convs", lambda i, x, w, k=k: i < n - k)
    go("f8 only at 128x128", lambda i, x, w: x.shape[-1] == 128)
    go("f8 except 128x128", lambda i, x, w: x.shape[-1] != 128)
    go("f8 only 3x3", lambda i, x, w: w.shape[-1] == 3)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sens":
    sensitivity()
