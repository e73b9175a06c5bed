"""Runs each hot kernel of a DeiT-Base layer back to back for ~1.5 s and reports the sustained time per launch with the
SM clock and board power seen meanwhile (nvidia-smi) -- what the power cap does to each kernel (development tool)."""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgevisiontransformer_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
D, H, I, S = 768, 12, 3072, 197
M = B * S


class Smi:
    def __init__(self):
        self.rows, self.stop = [], threading.Event()
    def run(self):
        while not self.stop.is_set():
            out = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout.strip()
            try:
                c, p = out.split(",")
                self.rows.append((float(c), float(p)))
            except Exception:
                pass
            self.stop.wait(0.1)


def sustained(name, fn, flops=0.0, nbytes=0.0, secs=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    smi = Smi()
    t = threading.Thread(target=smi.run, daemon=True)
    t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    smi.stop.set()
    t.join()
    ms = e0.elapsed_time(e1) / n
    rows = smi.rows[len(smi.rows) // 3:] or smi.rows or [(0, 0)]
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sorted(r[1] for r in rows)[len(rows) // 2]
    extra = f"{flops / ms / 1e9:7.0f} TFLOP/s" if flops else f"{nbytes / ms / 1e6:7.0f} GB/s"
    print(f"{name:10s} {ms:7.3f} ms  {extra}  sm {clk:5.0f} MHz  {pw:5.0f} W", flush=True)
    return ms


def main():
    x = torch.randn(M, D, device="cuda")
    g, b0 = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    wqkv = (torch.randn(3 * D, D, device="cuda") * 0.02).bfloat16()
    wo = (torch.randn(D, D, device="cuda") * 0.02).bfloat16()
    w1 = (torch.randn(I, D, device="cuda") * 0.02).bfloat16()
    w2 = (torch.randn(D, I, device="cuda") * 0.02).bfloat16()
    bq, bo, b1, b2 = (torch.zeros(n, device="cuda") for n in (3 * D, D, I, D))
    xn = ops.layernorm(x, g, b0, 1e-12)
    qkv = ops.linear(xn, wqkv, bq)
    ctx = ops.attention(qkv, B, S, H)
    h = ops.linear(xn, w1, b1, act="gelu_erf")
    t = {}
    t["ln"] = sustained("layernorm", lambda: ops.layernorm(x, g, b0, 1e-12), nbytes=M * D * 6)
    t["qkv"] = sustained("qkv", lambda: ops.linear(xn, wqkv, bq, out=qkv), flops=2.0 * M * 3 * D * D)
    t["attn"] = sustained("attention", lambda: ops.attention(qkv, B, S, H), flops=4.0 * S * S * 64 * H * B)
    t["proj"] = sustained("out-proj", lambda: ops.linear(ctx, wo, bo, residual=x, out=x, out_dtype=torch.float32), flops=2.0 * M * D * D)
    t["fc1"] = sustained("fc1+gelu", lambda: ops.linear(xn, w1, b1, act="gelu_erf", out=h), flops=2.0 * M * I * D)
    t["fc2"] = sustained("fc2", lambda: ops.linear(h, w2, b2, residual=x, out=x, out_dtype=torch.float32), flops=2.0 * M * D * I)
    layer = t["qkv"] + t["attn"] + t["proj"] + t["fc1"] + t["fc2"] + 2 * t["ln"]
    print(f"layer {layer:.3f} ms -> {B / (12 * layer) * 1e3:.0f} img/s encoder-only (sustained, kernel by kernel)")


    def whole():
        ops.layernorm(x, g, b0, 1e-12)
        ops.linear(xn, wqkv, bq, out=qkv)
        ops.attention(qkv, B, S, H)
        ops.linear(ctx, wo, bo, residual=x, out=x, out_dtype=torch.float32)
        ops.layernorm(x, g, b0, 1e-12)
        ops.linear(xn, w1, b1, act="gelu_erf", out=h)
        ops.linear(h, w2, b2, residual=x, out=x, out_dtype=torch.float32)


    ms = sustained("layer", whole, flops=2.0 * M * (4 * D * D + 2 * D * I) + 4.0 * S * S * 64 * H * B, secs=3.0)
    print(f"whole layer {ms:.3f} ms -> {B / (12 * ms) * 1e3:.0f} img/s encoder-only")


if __name__ == "__main__":
    main()
