"""ncu --csv launch list -> markdown table (per kernel family: launches, total ms, share; optional time-weighted tensor-pipe utilisation).

    python tools/launch_table.py gpurun_out/r2_launches.csv > profiles/r02_launches_cfg2.md
"""
import csv
import re
import sys
from collections import defaultdict

ENCODER = ("gemm2_kernel", "gemm_kernel", "attn_fwd_kernel", "attn_bwd_kernel", "ln_fwd_", "ln_bwd_", "colsum_kernel",
           "attn_dq_convert", "attn_delta")


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("at::native::", "").replace("(anonymous namespace)::", "")
    return name[:100]


def main(path):
    rows = defaultdict(dict)
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        rows[int(r["ID"])]["name"] = r["Kernel Name"]
        rows[int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    fam = defaultdict(lambda: [0, 0.0, 0.0])
    tot = 0.0
    enc_t = enc_tw = 0.0
    tp_key = next((k for k in next(iter(rows.values())) if "pipe_tensor" in k), None)
    for r in rows.values():
        ns = r.get("gpu__time_duration.sum", 0.0)
        k = short(r["name"])
        fam[k][0] += 1
        fam[k][1] += ns
        tot += ns
        if tp_key:
            fam[k][2] += ns * r.get(tp_key, 0.0)
            if any(s in r["name"] for s in ENCODER):
                enc_t += ns
                enc_tw += ns * r.get(tp_key, 0.0)
    mine = sum(v[1] for k, v in fam.items() if "simvgb::" in k)
    print("%d launches, %.1f ms summed; %.1f %% of the time is in libsimvg_b200 kernels (`simvgb::*`), %d of the launches.\n"
          % (len(rows), tot / 1e6, 100 * mine / tot, sum(v[0] for k, v in fam.items() if "simvgb::" in k)))
    if tp_key:
        print("Time-weighted `%s` over the encoder-block kernels (GEMMs, attention, LayerNorm / GELU / bias-gradient passes): "
              "**%.1f %%** of %.1f ms.\n" % (tp_key, enc_tw / max(enc_t, 1), enc_t / 1e6))
    hdr = "| kernel | launches | total ms | share |" + (" tensor pipe % |" if tp_key else "")
    print(hdr)
    print("|---|---:|---:|---:|" + ("---:|" if tp_key else ""))
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:45]:
        line = "| `%s` | %d | %.3f | %.1f %% |" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot)
        if tp_key:
            line += " %.1f |" % (v[2] / max(v[1], 1))
        print(line)


if __name__ == "__main__":
    main(sys.argv[1])
