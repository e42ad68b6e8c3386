"""Print selected metrics of an ncu raw CSV export (ncu -i X.ncu-rep --page raw --csv)."""
import csv
import re
import sys

pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|sm__cycles_elapsed.avg.per_second|pipe_alu|pipe_fma|pipe_tensor|pipe_xu|pipe_uniform|"
                 r"inst_executed.avg.per_cycle_elapsed|issue_active.avg.pct|lts__t_bytes.sum|lts__t_sectors.sum$|"
                 r"dram__bytes_(read|write).sum$|sm__throughput|lts__throughput|l1tex__throughput|registers_per_thread|sm__warps_active.avg.pct|"
                 r"lts__t_sector_hit_rate|gpu__dram_throughput|smsp__average_warps_issue_stalled.*_per_issue_active|l1tex__m_xbar2l1tex_read_bytes.sum$")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d["Kernel Name"], d["Grid Size"], d["Block Size"])
    for k, u in zip(hdr, units):
        if pat.search(k) and d[k] not in ("", "0"):
            print(f"   {k} = {d[k]} {u}")
