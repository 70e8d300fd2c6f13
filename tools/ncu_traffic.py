#!/usr/bin/env python
"""Turn an ``ncu --set full`` capture of the microbenchmark into the table bench.py reads for ``roofline.traffic``.

    ncu -i gpurun_out/full_kernels.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv "profiles/r2_full_kernels_raw.csv" > profiles/r2_ncu_traffic.json

Per kernel family (the names bench.py's roofline table uses) the launch with the most DRAM traffic is kept:
``dram_bytes`` = dram__bytes_read.sum + dram__bytes_write.sum of that ONE launch, plus its duration and grid, so the
figure can be set against the algorithmic bytes of the same tensor (the microbenchmark runs one tensor per launch).
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def family(name):
    """Kernel function name -> the family key of ops.py's _Timed labels (first word of the label)."""
    base = re.sub(r"^void\s+", "", name)
    base = re.sub(r"^oodfq::", "", base)
    fn = base.split("(")[0]
    plain, _, targs = fn.partition("<")
    targs = [t.strip() for t in targs.rstrip(">").split(",")] if targs else []
    if plain.startswith("res_tail_bwd"):
        # the step only launches the variants that take the forward's ReLU mask (4th template argument); the variant
        # that re-derives the mask from x1 and r is kept under its own key
        masked = len(targs) >= 4 and targs[3].strip("()bool ") in ("1", "true")
        return "res_tail_bwd_kernel" if masked else "res_tail_bwd_kernel<no mask>"
    if plain.startswith("res_tail_fwd"):
        return "res_tail_fwd_kernel"
    if plain.startswith("bn_pool_fwd_tma"):
        return "bn_pool_fwd_kernel"                  # the default stem forward (TMA-staged ring)
    if plain.startswith("bn_pool_fwd"):
        return "bn_pool_fwd_kernel<register fallback>"
    if plain.startswith("bn_pool_bwd"):
        return "bn_pool_bwd_kernel"
    if plain.startswith("s2d_stem"):
        return "s2d_stem_kernel"
    if plain.startswith("fq_flat"):
        return "fq_flat_kernel"
    if plain.startswith("act_calib_onchip"):
        return "act_calib_stats" if targs and targs[0].strip("()int ") not in ("0",) else "act_calib_onchip_kernel"
    m = re.match(r"bn_(nhwc|plane|group)_(fwd|bwdx|stats|bwd)", plain)
    if m:
        kind = m.group(2)
        if kind == "fwd":
            quant = len(targs) >= 2 and targs[1].strip("()bool ") in ("1", "true")
            return "bn_*_fwd_kernel<relu,quant>" if quant else "bn_*_fwd_kernel"
        return f"bn_*_{kind}_kernel"
    if plain.startswith("energy") or plain.startswith("plane_energy"):
        return "energy_*_kernel"
    return plain


def main(path, source):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(d, name, table):
        v = float(d[col[name]].replace(",", ""))
        return v * table.get(units[col[name]], 1.0)

    best = {}
    for d in data:
        if len(d) < len(hdr):
            continue
        name = d[col["Kernel Name"]]
        fam = family(name)
        bytes_ = val(d, "dram__bytes_read.sum", UNIT) + val(d, "dram__bytes_write.sum", UNIT)
        row = {"dram_bytes": int(bytes_), "dram_read": int(val(d, "dram__bytes_read.sum", UNIT)),
               "dram_write": int(val(d, "dram__bytes_write.sum", UNIT)),
               "duration_us": round(val(d, "gpu__time_duration.sum", TIME), 2),
               "launch": f"{name.split('(')[0]} grid {d[col['Grid Size']]} block {d[col['Block Size']]}"}
        if fam not in best or row["dram_bytes"] > best[fam]["dram_bytes"]:
            best[fam] = row
    json.dump({"source": source, "kernels": best}, sys.stdout, indent=1, sort_keys=True)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
