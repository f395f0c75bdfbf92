"""Shared-memory wavefronts per SASS instruction of an .ncu-rep (source page): excessive vs ideal, top offenders.
usage: python tools/ncu_conflicts.py <report.ncu-rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; out = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            exc, wf, ideal = float(d["L1 Wavefronts Shared Excessive"] or 0), float(d["L1 Wavefronts Shared"] or 0), float(d["L1 Wavefronts Shared Ideal"] or 0)
        except ValueError:
            continue
        if wf > 0:
            out.append((exc, wf, ideal, d["Address"], d["Source"][:90]))
tot_e, tot_w = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"shared wavefronts {tot_w:.3g}, excessive {tot_e:.3g} ({100 * tot_e / max(tot_w, 1):.1f} %)")
for exc, wf, ideal, addr, src in sorted(out, reverse=True)[:top]:
    print(f"{exc:12.0f} excessive of {wf:12.0f} (ideal {ideal:12.0f})  {addr}  {src}")
