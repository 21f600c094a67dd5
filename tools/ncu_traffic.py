"""Writes profiles/r02_ncu_dram_bytes.csv — kernel, duration, dram bytes (read + write) per captured launch — from one
`ncu --set full` report, so that bench.py's `roofline.traffic` comes from a committed measurement, not a hand-typed constant.
Usage: python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep [more reports] > profiles/r02_ncu_dram_bytes.csv"""
import csv
import subprocess
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
w = csv.writer(sys.stdout)
w.writerow(["report", "kernel", "duration_us", "dram_bytes", "dram_read_bytes", "dram_write_bytes", "grid", "regs"])
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def col(r, name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * UNITS.get(units[i], 1.0), units[i]
    for r in rows[2:]:
        rd, _ = col(r, "dram__bytes_read.sum")
        wr, _ = col(r, "dram__bytes_write.sum")
        dur, du = col(r, "gpu__time_duration.sum")
        dur_us = dur / 1e3 if du in ("ns", "nsecond") else dur * 1e3 if du in ("ms", "msecond") else dur
        w.writerow([rep.split("/")[-1], r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("uf::", ""), "%.2f" % dur_us,
                    "%.0f" % (rd + wr), "%.0f" % rd, "%.0f" % wr, r[hdr.index("launch__grid_size")], r[hdr.index("launch__registers_per_thread")]])
