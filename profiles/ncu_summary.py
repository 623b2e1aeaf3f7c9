"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline uses.
usage: python profiles/ncu_summary.py report.ncu-rep [> profiles/rNN_<name>.summary.txt]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"== {d['Kernel Name'][:90]}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
    for k in KEYS:
        if k in d:
            print(f"   {k:70s} {d[k]:>18s} {units[hdr.index(k)]}")
    try:
        tu = units[hdr.index("gpu__time_duration.sum")]
        t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[tu.replace("second", "s").replace("usecond", "us").replace("nsecond", "ns").replace("msecond", "ms")]
        u = units[hdr.index("dram__bytes_read.sum")]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * sc
        u = units[hdr.index("dram__bytes_write.sum")]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * sc
        print(f"   -> dram traffic {rd + wr:.4e} B per launch ({(rd + wr) / t / 1e9:.1f} GB/s under the profiler)")
    except Exception as e:
        print("   (no traffic summary)", e)
