"""Turn the .ncu-rep captures under gpurun_out/ into the committed artefacts under profiles/:
  profiles/r01_ncu_<name>.txt      headline metrics, stall mix and hottest source lines per kernel
  profiles/traffic_per_px.json     dram bytes per label-resolution pixel of the streaming kernels (bench.py reads it)
Usage: python scripts/make_profiles.py <rep> <pixels in the captured launch> <workload> [<rep> <pixels> <workload> ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = {"k3f_pass1": "k3f_pass1", "k3f_pass2": "k3f_pass2", "k3t_pass2": "k3t_pass2", "k_upsample4": "k_upsample", "k_upsample4_adjoint": "k_upsample_adjoint", "k3_pass1": "k3_pass1", "k3_pass2": "k3_pass2",
         "k_bce2_fast": "k_bce2_fast", "k_bce2_fused": "k_bce2_fused", "k_decode_vec": "k_decode", "k_decode": "k_decode", "k3f_prep": "k3f_prep"}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "raw"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    args = sys.argv[1:]
    tab_path = os.path.join(ROOT, "profiles", "traffic_per_px.json")
    tab = json.load(open(tab_path)) if os.path.exists(tab_path) else {}
    seen = {}
    for rep, px, workload in zip(args[0::3], args[1::3], args[2::3]):
        px = float(px)
        hdr, units, rows = raw(rep)
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            short = next((k for k in STAGE if k + "<" in name or name.split("(")[0].endswith(k)), None)
            if short is None:
                continue
            rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            key = f"{workload}:{STAGE[short]}"
            if key in seen and seen[key] >= rd + wr:
                continue      # e.g. the instantiation of k3f_pass2 that serves no image of the batch and leaves at once
            seen[key] = rd + wr
            tab[key] = {
                "bytes_per_px": (rd + wr) / px, "read_per_px": rd / px, "write_per_px": wr / px,
                "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of {short}, "
                          f"{os.path.basename(rep)}, {int(px)} pixels in the captured launch"}
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, "25"],
                              capture_output=True, text=True).stdout
        dst = os.path.join(ROOT, "profiles", os.environ.get("PROFILE_ROUND", "r02") + "_ncu_" + os.path.basename(rep).replace(".ncu-rep", "") + ".txt")
        with open(dst, "w") as fh:
            fh.write(f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on, "
                     f"{int(px)} label-resolution pixels per launch ({workload})\n" + summ)
        print("wrote", dst)
    with open(tab_path, "w") as fh:
        json.dump(tab, fh, indent=1, sort_keys=True)
    print("wrote", tab_path)


if __name__ == "__main__":
    main()
