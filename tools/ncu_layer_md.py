"""Markdown summary of the `ncu --set full` capture of one encoder layer (tools/profile_target.py) -> profiles/*.md.

    python tools/ncu_layer_md.py gpurun_out/r01_layer_v3.ncu-rep 1024 > profiles/r01_layer_ncu_full.md
"""
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


def main(path, batch):
    M = batch * 197
    titles = [f"LayerNorm (ln_rows4_kernel<6,bf16>) rows={M} D=768 f32 -> bf16",
              f"QKV GEMM (gemm_pair_kernel<256,bf16 out,no act>) M={M} N=2304 K=768",
              f"fused attention (attention_kernel) B={batch} S=197 heads=12",
              f"out-proj GEMM (gemm_pair_kernel<256,f32 TMA reduce-add>) M={M} N=768 K=768",
              "LayerNorm (second of the layer)",
              f"FC1 GEMM + erf-GELU (gemm_pair_kernel<256,bf16 out,gelu_erf>) M={M} N=3072 K=768",
              f"FC2 GEMM (gemm_pair_kernel<256,f32 TMA reduce-add>) M={M} N=768 K=3072"]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# Round 1 -- ncu `--set full --clock-control none` of one DeiT-Base encoder layer, per-GPU batch {batch} (final kernels of the round)\n")
    print("Command (after the same command exited 0 without ncu): `ncu --set full --clock-control none --import-source on "
          f"-k regex:\"gemm_|attention_kernel|ln_rows\" -s 14 -c 7 python tools/profile_target.py {batch}`\n")
    print("Per-launch values (cold-cache, serialised: compare shares, not absolutes).  GEMMs are the CTA-pair kernel "
          "(`tcgen05.mma.cta_group::2`, 256 x 256 tiles).\n")
    tot = wt = 0.0
    for r, title in zip(rows[2:], titles):
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
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


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
