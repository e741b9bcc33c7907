"""Compact table out of `ncu -i X.ncu-rep --page raw --csv`: one row per captured launch with the metrics the
roofline discussion needs (duration, DRAM bytes, tensor / FMA pipe activity, occupancy limits)."""
import csv, re, sys

COLS = [
    ("dur_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fma_%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fma_inst_%act", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("issue_%", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("xu_%act", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("warps_act_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]

def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("clica::", "")
    m = re.match(r"(?:void )?([\w:]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name

def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for r in data:
        out = []
        for c, m in COLS:
            i = idx.get(m)
            if i is None or r[i] == "":
                out.append("-"); continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if c == "dur_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if c.endswith("_MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            out.append(f"{v:.2f}" if abs(v) < 1000 else f"{v:.0f}")
        print(f"| {r[idx['ID']]} | `{short(r[idx['Kernel Name']])}` | " + " | ".join(out) + " |")

if __name__ == "__main__":
    main(sys.argv[1])
