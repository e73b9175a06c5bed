"""Markdown summary of the `ncu --set full` capture of one encoder layer (tools/profile_target.py) -> profiles/*.md.

    python tools/ncu_layer_md.py gpurun_out/r02_layer.ncu-rep 1024 [round] [profiles/traffic.json] > profiles/r02_layer_ncu_full.md

With a fourth argument the DRAM bytes of every GEMM launch are also written to that JSON index (kernel|shape -> bytes),
the table bench.py's `roofline.traffic` is looked up in.
"""
import json
import csv
import subprocess
import sys

KEYS = [("duration", "gpu__time_duration.sum"), ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
        ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("L2 % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("XU (MUFU) pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("warp instructions", "smsp__inst_executed.sum")]


def main(path, batch, rnd=2, traffic_json=None):
    M = batch * 197
    titles = [f"LayerNorm (ln_rows4_kernel<6,bf16>) rows={M} D=768 f32 -> bf16",
              f"QKV GEMM (gemm_pair_kernel<256,bf16 out,no act>) M={M} N=2304 K=768",
              f"fused attention (attention2_kernel) B={batch} S=197 heads=12",
              f"out-proj GEMM (gemm_pair_kernel<256,f32 TMA reduce-add>) M={M} N=768 K=768",
              "LayerNorm (second of the layer)",
              f"FC1 GEMM + erf-GELU (gemm_pair_kernel<256,bf16 out,gelu_erf>) M={M} N=3072 K=768",
              f"FC2 GEMM (gemm_pair_kernel<256,f32 TMA reduce-add>) M={M} N=768 K=3072"]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# Round {rnd} -- ncu `--set full --clock-control none` of one DeiT-Base encoder layer, per-GPU batch {batch} (final kernels of the round)\n")
    print("Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on "
          f"-k regex:\"gemm_|attention2_kernel|ln_rows\" -s 14 -c 7 python tools/profile_target.py {batch}`\n")
    print("Per-launch values (cold-cache, serialised: compare shares, not absolutes).  GEMMs are the CTA-pair kernel "
          "(`tcgen05.mma.cta_group::2`, 256 x 256 tiles).\n")
    tot = wt = 0.0
    gemm_keys = {1: ("gemm_pair_kernel<256,bf16,none>", 2304, 768), 3: ("gemm_pair_kernel<256,f32,reduce_add>", 768, 768),
                 5: ("gemm_pair_kernel<256,bf16,gelu_erf>", 3072, 768), 6: ("gemm_pair_kernel<256,f32,reduce_add>", 768, 3072)}
    traffic = {}

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    for i, (r, title) in enumerate(zip(rows[2:], titles)):
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        if i in gemm_keys:
            k, N, K = gemm_keys[i]
            traffic[f"{k}|M={M},N={N},K={K}"] = {
                "dram_read": to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
                "dram_write": to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
                "source": f"profiles/r{rnd:02d}_layer_ncu_full.md (ncu --set full, tools/profile_target.py {batch})"}
        print(f"## {title}\n\n| metric | value |\n|---|---|")
        for name, k in KEYS:
            print(f"| {name} (`{k}`) | {d.get(k, '?')} {u.get(k, '')} |")
        print()
        dur = float(d["gpu__time_duration.sum"].replace(",", ""))
        dur = dur / 1000 if u["gpu__time_duration.sum"] == "ns" else dur * 1000 if u["gpu__time_duration.sum"] == "ms" else dur
        tot += dur
        wt += dur * float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"].replace(",", ""))
    print("## Layer summary\n")
    print(f"Sum of the 7 launches: {tot:.1f} us for {batch} images.  Time-weighted tensor-pipe activity over the whole layer "
          f"(LayerNorm and attention included): **{wt / tot:.1f} %** (target in BASELINE.json: >= 60 %).")
    if traffic_json:
        try:
            old = json.load(open(traffic_json))
        except Exception:
            old = {}
        old.update(traffic)
        json.dump(old, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1024, int(sys.argv[3]) if len(sys.argv) > 3 else 2,
         sys.argv[4] if len(sys.argv) > 4 else None)
