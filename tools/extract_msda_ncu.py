"""profiles/ncu_msda_step_rN.txt + profiles/ncu_msda_traffic_rN.json from an `ncu --set full` capture of the MSDA
kernels inside one train step:

    ncu --set full --clock-control none --profile-from-start off -k regex:msda_ -o gpurun_out/ncu_msda python bench.py --ncu
    ncu -i gpurun_out/ncu_msda.ncu-rep --page raw --csv > /tmp/ncu_msda_raw.csv
    python tools/extract_msda_ncu.py /tmp/ncu_msda_raw.csv profiles/ncu_msda_step_r1.txt profiles/ncu_msda_traffic_r1.json

bench.py reads the JSON for `roofline.traffic` (DRAM bytes per launch of the dominant kernel)."""
import collections
import csv
import json
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
SCALE = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1, 'us': 1, 'ms': 1e3, 'ns': 1e-3}


def main(raw_csv, out_txt, out_json):
    rows = list(csv.reader(open(raw_csv)))
    h, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}

    def val(r, n):
        return float(r[idx[n]].replace(',', '')) * SCALE.get(units[idx[n]], 1)
    groups = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        kind = 'fwd' if 'msda_fwd' in name else 'bwd'
        shape = 'enc' if val(r, 'gpu__time_duration.sum') > 120 else 'dec'
        groups.setdefault(f'msda_{kind}_{shape}', []).append(r)
    out = ["# ncu --set full --clock-control none --profile-from-start off -k regex:msda_ python bench.py --ncu  "
           "(one eager train step, configs[1]: N=2, S=22223; B200)",
           "# 24 launches: 6 encoder + 6 decoder forward, 6 + 6 backward (fused-prologue kernels).  Averages per group; "
           "per-launch DRAM bytes feed roofline.traffic in bench.py.", ""]
    traffic = {}
    for g, rs in groups.items():
        out.append(f"== {g}: {len(rs)} launches, kernel {rs[0][idx['Kernel Name']][:90]}")
        for n in WANT:
            if n in idx:
                vs = [float(r[idx[n]].replace(',', '')) for r in rs]
                out.append(f"{n:80s} avg {sum(vs) / len(vs):16.3f} {units[idx[n]]:14s} min {min(vs):.3f} max {max(vs):.3f}")
        tb = [val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum') for r in rs]
        traffic[g] = dict(dram_bytes_per_launch=sum(tb) / len(tb), launches=len(rs),
                          avg_us_under_ncu=sum(val(r, 'gpu__time_duration.sum') for r in rs) / len(rs))
        out.append(f"{'dram read+write per launch':80s} avg {sum(tb) / len(tb) / 1e6:16.3f} MB")
        out.append("")
    open(out_txt, 'w').write("\n".join(out))
    json.dump(dict(source=f"{out_txt} (ncu --set full, one capture per launch, B200)", kernels=traffic),
              open(out_json, 'w'), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:4])
