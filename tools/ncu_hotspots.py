"""Per-region summary of an `ncu --set full --import-source on` capture (development tool).

    python tools/ncu_hotspots.py <report.ncu-rep> [nregions]

Reads the SASS source page, splits the kernel at its backward-branch targets / barriers into
regions, and prints for every region: share of stall samples, instructions executed, shared /
global wavefronts, and the dominating stall reasons."""
import csv, io, subprocess, sys, collections

def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))

def main():
    rep = sys.argv[1]
    rows = load(rep)
    tot_s = sum(int(r["# Samples"]) for r in rows)
    tot_i = sum(int(r["Instructions Executed"]) for r in rows)
    print(f"{len(rows)} SASS instructions, {tot_s} samples, {tot_i} warp instructions executed")
    stall_keys = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    # regions: cut at BAR.SYNC and at instruction-count discontinuities (factor 1.5)
    regions, cur = [], []
    prev = None
    for r in rows:
        n = int(r["Instructions Executed"])
        if cur and (("BAR.SYNC" in r["Source"]) or (prev is not None and (n > 1.6 * prev + 50 or prev > 1.6 * n + 50))):
            regions.append(cur); cur = []
        cur.append(r); prev = n
    regions.append(cur)
    print(f"{'addr':>8} {'ninstr':>6} {'exec/instr':>11} {'%inst':>6} {'%samples':>8} {'shwf%':>6} {'glsect%':>7}  top stalls / first instruction")
    tot_sh = sum(int(r["L1 Wavefronts Shared"]) for r in rows) or 1
    tot_gl = sum(int(r["L2 Theoretical Sectors Global"]) for r in rows) or 1
    for reg in regions:
        s = sum(int(r["# Samples"]) for r in reg)
        i = sum(int(r["Instructions Executed"]) for r in reg)
        if s < 0.004 * tot_s and i < 0.004 * tot_i:
            continue
        sh = sum(int(r["L1 Wavefronts Shared"]) for r in reg)
        gl = sum(int(r["L2 Theoretical Sectors Global"]) for r in reg)
        st = collections.Counter()
        for r in reg:
            for k in stall_keys:
                st[k[6:]] += int(r[k])
        top = " ".join(f"{k}={v * 100 // max(1, s)}%" for k, v in st.most_common(3))
        addr = reg[0]["Address"][-5:]
        print(f"{addr:>8} {len(reg):6d} {i // len(reg):11d} {100 * i / tot_i:6.1f} {100 * s / tot_s:8.1f} {100 * sh / tot_sh:6.1f} {100 * gl / tot_gl:7.1f}  {top} | {reg[0]['Source'].strip()[:40]}")

if __name__ == "__main__":
    main()
