"""Offline reading of an `ncu --set full --import-source on` report: where do the warps of a kernel wait?

    python tools/ncu_stalls.py gpurun_out/ncu_msda_step_final.ncu-rep [kernel-substring] [--top 25] [--trace]

No GPU needed (`ncu -i` only parses the report).  For every kernel whose name contains the substring (largest launch of
each distinct name) it prints
  * the stall reasons per issued instruction (`smsp__average_warps_issue_stalled_*`), resident / eligible warps per
    scheduler, issue rate, instruction-cache hit rate                                         (raw page)
  * stall samples by opcode and the top stalled SASS instructions                              (source page)
  * with --trace: the order of LDG / REDG / SHFL / branch instructions in the hot loop with their samples, which shows
    how many loads are in flight before the first dependent instruction waits
This is how profiles/ncu_msda_stalls_r1.txt and the load-pipelining / x8 variants of the MSDA backward were derived.
"""
import collections
import csv
import io
import re
import subprocess
import sys

REASONS = ["long_scoreboard", "short_scoreboard", "wait", "not_selected", "no_instruction", "branch_resolving",
           "barrier", "math_pipe_throttle", "mio_throttle", "lg_throttle", "tex_throttle", "dispatch_stall", "drain",
           "membar", "sleeping", "imc_miss"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def raw_summary(rep, sub):
    rows = page(rep, "raw")
    hdr = rows[0]
    ni, di = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    best = {}
    for r in rows[2:]:
        if len(r) > di and sub in r[ni]:
            d = num(r[di]) or 0.0
            if r[ni] not in best or d > best[r[ni]][0]:
                best[r[ni]] = (d, r)
    for name, (d, r) in best.items():
        print(f"== {name[:110]}\n   duration {d:.1f} {rows[1][di]}")
        get = lambda k: num(r[hdr.index(k)]) if k in hdr else None
        for k, label in (("smsp__warps_active.avg.per_cycle_active", "warps resident / scheduler"),
                         ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible / scheduler"),
                         ("smsp__issue_active.avg.per_cycle_active", "instructions issued / cycle / scheduler"),
                         ("smsp__average_warp_latency_per_inst_issued.ratio", "warp latency per issued instruction"),
                         ("sm__icc_request_hit_rate.pct", "instruction cache hit rate %")):
            v = get(k)
            if v is not None:
                print(f"   {label:42s} {v:8.3f}")
        vals = [(get(f"smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio"), x) for x in REASONS]
        for v, x in sorted((p for p in vals if p[0]), reverse=True):
            print(f"   stalled on {x:31s} {v:8.3f}")


def source_sections(rep):
    secs, cur = [], None
    for row in page(rep, "source", ("--print-source", "sass")):
        if row and row[0] == "Kernel Name":
            cur = dict(name=row[1], rows=[], hdr=None)
            secs.append(cur)
        elif cur is not None and row:
            if row[0] == "Address":
                cur["hdr"] = row
            else:
                cur["rows"].append(row)
    return secs


def source_summary(rep, sub, top, trace):
    best = {}
    for s in source_sections(rep):
        if sub not in s["name"] or not s["hdr"]:
            continue
        ie = s["hdr"].index("Instructions Executed")
        tot = sum(int(r[ie]) for r in s["rows"])
        if s["name"] not in best or tot > best[s["name"]][0]:
            best[s["name"]] = (tot, s)
    for name, (tot, s) in best.items():
        h = s["hdr"]
        ie, sa, src = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
        samples = sum(int(r[sa]) for r in s["rows"]) or 1
        print(f"== {name[:110]}\n   {len(s['rows'])} SASS lines, {tot} warp instructions executed, {samples} stall samples")
        by = collections.Counter()
        ex = collections.Counter()
        for r in s["rows"]:
            m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[src])
            op = m.group(2) if m else "?"
            by[op] += int(r[sa])
            ex[op] += int(r[ie])
        print("   samples by opcode (the instruction a warp is stalled AT)      executed")
        for op, c in by.most_common(12):
            print(f"     {op:10s} {c:7d} {100 * c / samples:5.1f} %   {100 * ex[op] / max(tot, 1):5.1f} %")
        print("   top stalled instructions")
        for r in sorted(s["rows"], key=lambda r: -int(r[sa]))[:top]:
            print(f"     {int(r[sa]):6d} {100 * int(r[sa]) / samples:4.1f} %  {r[src].strip()[:100]}")
        if trace:
            print("   order of memory / shuffle / branch instructions (index, samples, instruction)")
            shown = 0
            for i, r in enumerate(s["rows"]):
                t = r[src].strip()
                m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", t)
                if m and (m.group(2) in ("LDG", "REDG", "STG", "SHFL", "BSSY", "BSYNC", "LDS", "STS", "ATOMS")
                          or int(r[sa]) * 200 > samples):
                    print(f"     {i:5d} {int(r[sa]):6d}  {t[:90]}")
                    shown += 1
                    if shown > 400:
                        break


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rep, sub = args[0], (args[1] if len(args) > 1 else "")
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw_summary(rep, sub)
    source_summary(rep, sub, top, "--trace" in sys.argv)


if __name__ == "__main__":
    main()
