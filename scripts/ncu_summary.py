"""Summarise an .ncu-rep (needs `ncu` on PATH): headline metrics, opcode mix, stall reasons and the hottest
source lines per kernel.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines]"""
import csv
import io
import subprocess
import sys


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def headline(rep):
    rows = ncu_csv(rep, "--page", "raw")
    hdr = rows[0]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
    for r in rows[2:]:
        print("----")
        for k in keys:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]}  [{rows[1][hdr.index(k)]}]")


def source(rep, top):
    rows = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    fn = cur_file = hdr = None
    lines, ops, stalls, tot = {}, {}, {}, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1].split("(")[0][-40:]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or fn is None:
            continue
        try:
            line = int(r[0])
            n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        sass = r[3].split()
        op = (sass[1] if sass and sass[0].startswith("@") and len(sass) > 1 else (sass[0] if sass else "")).split(".")[0]
        lines[(fn, cur_file, line, r[1].strip()[:100])] = lines.get((fn, cur_file, line, r[1].strip()[:100]), 0) + n
        ops[(fn, op)] = ops.get((fn, op), 0) + n
        tot[fn] = tot.get(fn, 0) + n
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stalls[(fn, h)] = stalls.get((fn, h), 0) + int(r[i])
                except ValueError:
                    pass
    for f in tot:
        print(f"== {f}: {tot[f]} warp-instructions")
        st = sorted(((v, k[1]) for k, v in stalls.items() if k[0] == f), reverse=True)
        ssum = sum(v for v, _ in st) or 1
        print("   stalls: " + ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for v, k in st[:9]))
        op = sorted(((v, k[1]) for k, v in ops.items() if k[0] == f), reverse=True)
        print("   opcodes: " + ", ".join(f"{k} {100 * v / tot[f]:.1f}%" for v, k in op[:24]))
        for v, k in sorted(((v, k) for k, v in lines.items() if k[0] == f), reverse=True)[:top]:
            print(f"   {100 * v / tot[f]:5.1f}% {k[1]}:{k[2]}  {k[3]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    headline(rep)
    source(rep, top)
