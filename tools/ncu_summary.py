"""Developer tool: one line per captured launch from an `ncu --page raw --csv` export."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "launch__registers_per_thread"]
idx = [hdr.index(w) if w in hdr else None for w in want]
print("units:", [rows[1][i] if i is not None else "-" for i in idx])
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(r[0], name[name.find("::") + 2:][:48].ljust(48), [r[i][:9] if i is not None else "-" for i in idx])
