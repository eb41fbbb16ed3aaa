"""Per-step counters of a workload from an ncu launch list -> profiles/counters.json.

    python tools/counters_from_ncu.py <launches.csv> <key> <builds> [--source NAME] [--skip-builds S]

<launches.csv> is an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,...` log of `tools/prof_jk.py` (a known
number of identical J/K builds, <builds>); <key> is "<workload>/<boys>/<world>" as bench.py
looks it up.  The first S builds (default 1: it holds the one-off Schwarz launches) are skipped
by dropping the first S/<builds> of the ERI launches; the per-step figures are sums over the
launches of ONE build.  bench.py reads the file for `roofline.traffic` (DRAM bytes per step) and
`roofline.fp64_pipe_active` (FP64 warp instructions per step over the LIVE kernel time).
"""
import argparse
import collections
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ik, im, iv, iid = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        try:
            per.setdefault(r[iid], {"name": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
        except ValueError:
            pass
    return list(per.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("key")
    ap.add_argument("builds", type=int)
    ap.add_argument("--source", default=None)
    ap.add_argument("--skip-builds", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "counters.json"))
    a = ap.parse_args()
    # (the host-buffer entry point probes D for asymmetry before every build: not part of the
    # build's own launch pattern)
    launches = [l for l in load(a.csv) if not l["name"].startswith("asym_probe")]
    # launches of the J/K builds proper: everything from the first pack_d of a build on; the
    # set-up (Schwarz) launches come before the first pack_d_kernel
    first = next(i for i, l in enumerate(launches) if l["name"].startswith("pack_d_kernel"))
    body = launches[first:]
    per_build = len(body) // a.builds
    assert per_build * a.builds == len(body), (len(body), a.builds)
    one = body[a.skip_builds * per_build:(a.skip_builds + 1) * per_build] if a.builds > a.skip_builds else body[:per_build]
    tot = collections.defaultdict(float)
    for l in one:
        for k, v in l.items():
            if k != "name":
                tot[k] += v
    eri = [l for l in one if "eri_" in l["name"]]
    top = max(eri, key=lambda l: l.get("gpu__time_duration.sum", 0.0))
    entry = {
        "source": a.source or os.path.basename(a.csv),
        "launches_per_step": per_build,
        "serialised_ms_per_step": tot["gpu__time_duration.sum"] / 1e6,
        "dram_bytes_per_step": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
        "fp64_warp_inst_per_step": tot.get("smsp__inst_executed_pipe_fp64.sum", 0.0),
        "warp_inst_per_step": tot.get("smsp__inst_executed.sum", 0.0),
        "local_ld_st_warp_inst_per_step": tot.get("smsp__inst_executed_op_local_ld.sum", 0.0)
        + tot.get("smsp__inst_executed_op_local_st.sum", 0.0),
        "top_launch": {"kernel": top["name"][:80], "ms": top.get("gpu__time_duration.sum", 0.0) / 1e6,
                       "share_of_step": top.get("gpu__time_duration.sum", 0.0) / tot["gpu__time_duration.sum"]},
    }
    try:
        with open(a.out) as fh:
            db = json.load(fh)
    except (OSError, ValueError):
        db = {}
    db[a.key] = entry
    with open(a.out, "w") as fh:
        json.dump(db, fh, indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
