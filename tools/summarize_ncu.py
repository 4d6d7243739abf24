"""Turn ncu output into the markdown summaries kept under profiles/.

  launches <launches.csv> [title]      launch list of `ncu --metrics gpu__time_duration.sum --csv`
  full <report.ncu-rep> [kernel-regex] key metrics of one `ncu --set full` capture (needs ncu here)
"""
import collections
import csv
import io
import subprocess
import sys


def _rows(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def launches(path, title="ncu launch list"):
    rows = [r for r in _rows(open(path).read()) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg.setdefault(k, dict(n=0, t=0.0, grid=r["Grid Size"], block=r["Block Size"]))
        a["n"] += 1
        a["t"] += us
    tot = sum(a["t"] for a in agg.values())
    print("# %s\n" % title)
    print("| kernel | launches | total µs | avg µs | share | grid | block |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% | %s | %s |" % (k[:60], a["n"], a["t"], a["t"] / a["n"], 100 * a["t"] / tot,
                                                                  a["grid"], a["block"]))
    print("\ntotal %.1f µs over %d launches" % (tot, sum(a["n"] for a in agg.values())))


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
]


def full(path, kernel=None):
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"]
    if kernel:
        cmd += ["-k", "regex:" + kernel]
    text = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if not r:
            continue
        d = dict(zip(names, r))
        print("## `%s`  grid %s block %s\n" % (d.get("Kernel Name", "?")[:70], d.get("Grid Size"), d.get("Block Size")))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print("| %s | %s | %s |" % (k, d[k], units[names.index(k)]))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:])
    else:
        full(*sys.argv[2:])
