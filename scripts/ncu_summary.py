"""Write a markdown summary of an ncu report (raw metrics + hottest
instructions) for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
]
lines = [f"# ncu summary: `{rep}`", "", note, "", "| metric | value | unit |", "|---|---|---|"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.endswith("_per_issue_active.ratio"):
        try:
            if h.endswith("_per_issue_active.ratio") and float(v) < 0.01:
                continue
        except ValueError:
            pass
        lines.append(f"| {h} | {v} | {u} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
# the page lists every function of the module (the kernel and the device
# functions it calls): "Kernel Name" line, column header, instructions
body, func, ix, shdr = [], "", None, None
for r in srows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        func = r[1]
    elif r[0] == "Address":
        shdr = r
        ix = {h: i for i, h in enumerate(shdr)}
    elif ix is not None and len(r) >= len(shdr) - 1:
        body.append((func, r))
tot = sum(int(r[ix["# Samples"]]) for _, r in body)
per_func = {}
for f, r in body:
    per_func[f] = per_func.get(f, 0) + int(r[ix["# Samples"]])
lines += ["", f"## samples per function ({len(body)} SASS instructions, {tot} samples)", "",
          "| function | % samples |", "|---|---|"]
top = sorted(per_func.items(), key=lambda kv: -kv[1])
for f, c in top[:12]:
    lines.append(f"| {f} | {100 * c / max(tot, 1):.2f} |")
if len(top) > 12:
    rest = sum(c for _, c in top[12:])
    lines.append(f"| ({len(top) - 12} more) | {100 * rest / max(tot, 1):.2f} |")
lines += ["", "## hottest instructions", "",
          "| % samples | function | executed | smem wavefronts (actual/ideal) | SASS | top stall reasons |",
          "|---|---|---|---|---|---|"]
stall_cols = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
for f, r in sorted(body, key=lambda fr: -int(fr[1][ix["# Samples"]]))[:25]:
    stalls = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]].isdigit()), reverse=True)[:2]
    lines.append(f'| {100*int(r[ix["# Samples"]])/max(tot, 1):.2f} | {f} | {r[ix["Instructions Executed"]]} | '
                 f'{r[ix["L1 Wavefronts Shared"]]}/{r[ix["L1 Wavefronts Shared Ideal"]]} | '
                 f'`{r[ix["Source"]].strip()}` | ' + ", ".join(f"{n} {c}" for c, n in stalls) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
