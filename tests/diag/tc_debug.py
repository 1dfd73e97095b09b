"""Debug aid for the tcgen05 kernel: tiny shapes, one-hot activations -> shows which (k, n) each output picked up."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import eetq_b200  # noqa: E402
from eetq_b200 import _cabi  # noqa: E402
from eetq_b200.ops import w8_a16_gemm_bias  # noqa: E402
from oracle import w8a16_oracle as o  # noqa: E402


def run(M, K, N, dtype=torch.float16, onehot=False):
    dev = torch.device("cuda", 0)
    w = o.synth_weight(K, N, seed=1)
    q, s, _ = o.quantize(w)
    s = s.to(dtype)
    wq = o.b200_layout(q).to(dev)
    if onehot:
        x = torch.zeros(M, K, dtype=dtype)
        for t in range(M):
            x[t, t % K] = 1
    else:
        x = o.synth_act(M, K, dtype=dtype)
    y = w8_a16_gemm_bias(x.to(dev), wq, s.to(dev), None, flags=_cabi.FLAG_FORCE_TC)
    torch.cuda.synchronize()
    if dtype == torch.float16:
        yr = o.gemm(x, q, s)
    else:
        yr = ((x.float() @ q.float()) * s.float()).to(dtype)
    err = o.norm_rel_err(y.cpu(), yr)
    print(f"M={M} K={K} N={N} {dtype} onehot={onehot}: err={err:.3e} nan={torch.isnan(y).any().item()}", flush=True)
    if err > 1e-2 and onehot:
        wd = o.dequantize(q, s).float()  # [K, N]
        yc = y.cpu().float()
        for t in range(min(M, 16)):
            d = (wd - yc[t][None, :]).abs().sum(dim=1)
            kbest = int(d.argmin())
            print(f"  token {t} (expect k={t % K}): best match k={kbest} resid={d[kbest]:.3g} ; y[:4]={yc[t][:4].tolist()} exp={wd[t % K][:4].tolist()}")
    return err


if __name__ == "__main__":
    torch.cuda.set_device(0)
    ok = True
    if "--tiny" in sys.argv:   # compute-sanitizer target: a few small launches only
        for (M, K, N) in [(16, 512, 256), (64, 1024, 384), (256, 1024, 256)]:
            ok &= run(M, K, N) <= 1e-3
        print("TC_DEBUG", "PASS" if ok else "FAIL")
        sys.exit(0)
    for (M, K, N, oh) in [(16, 64, 128, True), (64, 64, 128, True), (16, 64, 128, False), (16, 128, 128, False), (16, 512, 256, False),
                          (32, 512, 256, False), (64, 1024, 256, False), (128, 1024, 256, False), (256, 1024, 256, False),
                          (300, 1024, 320, False), (16, 4096, 4096, False), (1024, 4096, 4096, False)]:
        ok &= run(M, K, N, onehot=oh) <= 1e-3
    ok &= run(64, 1024, 256, dtype=torch.bfloat16) <= 4e-3
    print("TC_DEBUG", "PASS" if ok else "FAIL")
