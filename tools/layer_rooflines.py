"""Per-layer floors of the C4 network (8 frames of 1024x1024) next to the measured per-op times.

  python tools/layer_rooflines.py profiles/r01_per_op_fast_epilogue.txt profiles/r01_layer_rooflines.md

Floors: tensor = algorithmic FLOPs / measured sustained bf16 peak (MEASURED_PEAKS.json, 1377 TF/s);
HBM = (activations in + out (+ fused pool output) at their storage width) / 6.0 TB/s (profiles/r01_store_probe.txt,
r01_bw_probe.txt: best streaming rate observed on this part).  The bound of a layer is the larger floor.
"""
import re
import sys

B = 8
PEAK_TF, PEAK_TBS = 1376.9, 6.0
# (op index in the per-op table, name, Cin, Cout, input grid H(=W), output grid, taps, pooled copy, out bytes/elem)
LAYERS = [
    (1, "conv 1->16 (Toeplitz view)", 1, 16, 1024, 1024, 9, False, 2),
    (2, "conv 16->16 + pool", 16, 16, 1024, 1024, 9, True, 2),
    (4, "conv 16->32", 16, 32, 512, 512, 9, False, 2),
    (5, "conv 32->32 + pool", 32, 32, 512, 512, 9, True, 2),
    (7, "conv 32->64", 32, 64, 256, 256, 9, False, 2),
    (8, "conv 64->64 + pool", 64, 64, 256, 256, 9, True, 2),
    (10, "conv 64->128", 64, 128, 128, 128, 9, False, 2),
    (11, "conv 128->128 + pool", 128, 128, 128, 128, 9, True, 2),
    (13, "conv 128->256", 128, 256, 64, 64, 9, False, 2),
    (14, "conv 256->256 + pool", 256, 256, 64, 64, 9, True, 2),
    (16, "conv 256->512 (middle)", 256, 512, 32, 32, 9, False, 2),
    (17, "conv 512->512 (middle)", 512, 512, 32, 32, 9, False, 2),
    (18, "tconv 512->256 (32->64)", 512, 256, 32, 64, 2.25, False, 2),
    (19, "conv 512->256", 512, 256, 64, 64, 9, False, 2),
    (20, "conv 256->256", 256, 256, 64, 64, 9, False, 2),
    (21, "tconv 256->128 (64->128)", 256, 128, 64, 128, 2.25, False, 2),
    (22, "conv 256->128", 256, 128, 128, 128, 9, False, 2),
    (23, "conv 128->128", 128, 128, 128, 128, 9, False, 2),
    (24, "tconv 128->64 (128->256)", 128, 64, 128, 256, 2.25, False, 2),
    (25, "conv 128->64", 128, 64, 256, 256, 9, False, 2),
    (26, "conv 64->64", 64, 64, 256, 256, 9, False, 2),
    (27, "head 1x1 64->13 (fp32 out)", 64, 13, 256, 256, 1, False, 4),
    (28, "head 1x1 128->24 (fp32 out)", 128, 24, 128, 128, 1, False, 4),
]


def main(src, out):
    ms = {}
    for line in open(src):
        m = re.match(r"\[op\s*(\d+)\] kind=\d+\s+([\d.]+) us", line)
        if m:
            ms[int(m.group(1))] = float(m.group(2))
    rows = ["| op | layer | measured us | GFLOP | tensor floor us | bytes MB | HBM floor us | bound | floor / measured |",
            "|---|---|---|---|---|---|---|---|---|"]
    tot_m = tot_f = 0.0
    layers = list(LAYERS)
    if ms.get(2, 1e9) < 10.0:        # fused first block (k_conv01): op 1 holds both convs + the pool, op 2 is an empty slot
        flops = sum(2.0 * 9 * ci * co * 1024 * 1024 * B for ci, co in ((1, 16), (16, 16)))
        t_tensor = flops / (PEAK_TF * 1e12) * 1e6
        byts = B * 1024 * 1024 * 1 + B * 512 * 512 * 16 * 2          # uint8 frames in, pooled fp16 out: nothing else leaves the SM
        t_hbm = byts / (PEAK_TBS * 1e12) * 1e6
        meas = ms[1] + ms[2]
        floor = max(t_tensor, t_hbm)
        tot_m += meas
        tot_f += floor
        rows.append(f"| 1+2 | fused first block 1->16->16 + pool (k_conv01) @1024² | {meas:.1f} | {flops / 1e9:.2f} | {t_tensor:.1f} | {byts / 1e6:.1f} | "
                    f"{t_hbm:.1f} | {'tensor' if t_tensor >= t_hbm else 'HBM'} | {floor / meas:.2f} |")
        layers = [l for l in layers if l[0] not in (1, 2)]
    for op, name, cin, cout, hin, hout, taps, pool, ob in layers:
        flops = 2.0 * taps * cin * cout * hout * hout * B
        t_tensor = flops / (PEAK_TF * 1e12) * 1e6
        in_b = B * hin * hin * cin * (1 if cin == 1 else 2)
        out_b = B * hout * hout * cout * ob * (1.25 if pool else 1.0)
        t_hbm = (in_b + out_b) / (PEAK_TBS * 1e12) * 1e6
        floor = max(t_tensor, t_hbm)
        meas = ms.get(op, float("nan"))
        tot_m += meas
        tot_f += floor
        rows.append(f"| {op} | {name} @{hin}² | {meas:.1f} | {flops / 1e9:.2f} | {t_tensor:.1f} | {(in_b + out_b) / 1e6:.1f} | {t_hbm:.1f} | "
                    f"{'tensor' if t_tensor >= t_hbm else 'HBM'} | {floor / meas:.2f} |")
    rows.append(f"| | **all conv layers** | **{tot_m:.1f}** | | | | | | **{tot_f / tot_m:.2f}** (sum of floors {tot_f:.1f} us) |")
    text = ("# C4 per-layer floors vs measured (8 frames, one B200)\n\nMeasured: CUDA-event per-op times of `bench.py` (`sb_model_profile_ops`), file `" + src +
            "`.  Floors: see tools/layer_rooflines.py.\n\n" + "\n".join(rows) + "\n")
    open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
