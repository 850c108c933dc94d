"""Developer tool: summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`).

  python tools/ncu_summary.py raw.csv                 # one line per captured launch
  python tools/ncu_summary.py raw.csv --json out.json # per-launch records for profiles/ (bench.py reads the DRAM
                                                      # traffic of the conv kernels from it: roofline.traffic)
"""
import csv
import json
import sys

WANT = {"us": "gpu__time_duration.sum",
        "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "ipc": "sm__inst_executed.avg.per_cycle_elapsed",
        "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l2_to_sm_bytes": "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm_cycles": "sm__cycles_elapsed.avg", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size", "block": "launch__block_size"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        rec = {"id": int(r[0]), "kernel": name[name.find("abl::") + 5:] if "abl::" in name else name}
        for k, m in WANT.items():
            if m not in hdr:
                continue
            i = hdr.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            rec[k] = v * SCALE.get(units[i], 1.0)
        out.append(rec)
    return out


def main():
    recs = load(sys.argv[1])
    if "--json" in sys.argv:
        json.dump({"source": "ncu --set full --clock-control none (cold-cache, serialised replays: compare shares)",
                   "launches": recs}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    for r in recs:
        print(f"{r['id']:3d} {r['kernel'][:58]:58s} {r.get('us', 0):8.1f} us  tensor {r.get('tensor_pipe_pct', 0):5.1f}%  "
              f"dram rd {r.get('dram_read_bytes', 0) / 1e6:8.1f} MB wr {r.get('dram_write_bytes', 0) / 1e6:8.1f} MB "
              f"({r.get('dram_pct', 0):4.1f}%)  L2->SM {r.get('l2_to_sm_bytes', 0) / 1e9:6.2f} GB  regs {int(r.get('regs', 0))}")


if __name__ == "__main__":
    main()
