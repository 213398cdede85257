"""Development helper: summarise the source page of an ncu report:
top instructions by stall samples and shared-memory wavefront excess."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("kernel:", rows[0][1][:100])
print("instructions:", len(body), "total samples:", tot)
srt = sorted(body, key=lambda r: -int(r[ix["# Samples"]]))
stall_cols = [h for h in hdr if h.startswith("stall_")]
for r in srt[:top]:
    stalls = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols if r[ix[c]].isdigit()), reverse=True)[:3]
    print(f'{r[ix["Address"]][-5:]} {int(r[ix["# Samples"]]):7d} {100*int(r[ix["# Samples"]])/tot:5.1f}% '
          f'exe={r[ix["Instructions Executed"]]:>10} wf={r[ix["L1 Wavefronts Shared"]]:>10}/'
          f'{r[ix["L1 Wavefronts Shared Ideal"]]:>10} {r[ix["Source"]].strip()[:60]:60s} '
          + " ".join(f"{n}:{c}" for c, n in stalls))
