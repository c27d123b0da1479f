"""Turn ncu output into the committed summaries under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md> [title]
        launches.csv = `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
                        --clock-control none -c N --csv --log-file launches.csv python bench.py ...`
    python tools/ncu_summary.py full <capture.ncu-rep> <out.md> <traffic.json> <rows_per_launch> [title]
        capture.ncu-rep = `ncu --set full --clock-control none --import-source on -k regex:... -c N -o capture python bench.py ...`
        (read back with `ncu -i capture.ncu-rep --page raw --csv`)

Names the launches of one Stage-1 forward by the packed op program (tools only; nothing here runs on a GPU).
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(name):
    for k in ("conv_res_tcgen05_kernel", "fc_tcgen05_kernel", "stem_tma_kernel", "stem_tc_kernel", "se_kernel", "sam_finish_kernel", "sam_gate_kernel",
              "fgvc_tail_kernel", "route_count_kernel", "route_scatter_kernel", "finalize_labels_kernel", "extract_blocks_kernel"):
        if k in name:
            if k in ("conv_res_tcgen05_kernel", "se_kernel", "stem_tc_kernel", "stem_tma_kernel", "fc_tcgen05_kernel") and "<" in name:
                return k + name[name.index("<"):name.index(">") + 1]
            return k
    return name.split("(")[0]


def stage1_op_names():
    from cnn_av1_research_b200 import packer, synth
    sd = synth.calibrated_state_dict("stage1", 0)
    return [op.name.replace("backbone.", "") for op in packer.backbone_ops(sd) + packer.head_ops("stage1", sd)]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    by_id = OrderedDict()
    for r in rows:
        e = by_id.setdefault(r[0], {"name": r[4]})
        e["time" if r[12].startswith("gpu__time") else "tensor"] = float(r[14].replace(",", ""))
        e["unit"] = r[13] if r[12].startswith("gpu__time") else e.get("unit", "ns")
    agg = OrderedDict()
    for e in by_id.values():
        t = e.get("time", 0.0) * ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(e.get("unit", "ns"), 1e-3))
        a = agg.setdefault(short(e["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += t * e.get("tensor", 0.0)
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n{len(by_id)} launches, cold-cache and serialised per-launch times: compare SHARES with bench.py's CUDA-event classes.\n\n")
        f.write("| kernel | launches | total ms | share | avg us | tensor-pipe active % (time-weighted) |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / total:.1f}% | {a[1] / a[0]:.1f} | {a[2] / a[1] if a[1] else 0:.1f} |\n")
    print(open(out).read())


def full(rep, out, traffic_json, rows_per_launch, title):
    # `rep` may also be the raw page already exported on the GPU box (`ncu -i capture.ncu-rep --page raw --csv > capture.csv`):
    # a --set full capture of 25 launches is larger than what gpurun copies back
    raw = open(rep).read() if rep.endswith(".csv") else \
        subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, key, default=0.0):
        try:
            return float(r[col[key]].replace(",", ""))
        except Exception:
            return default

    def scaled(r, key, want):          # normalise ncu's auto-scaled units
        v, u = g(r, key), units[col[key]]
        f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
        return v * f / want
    names = stage1_op_names()
    lines, traffic = [], {}
    for i, r in enumerate(data):
        k = short(r[col["Kernel Name"]])
        op = names[i] if i < len(names) else "(stage 2)"
        t_us = scaled(r, "gpu__time_duration.sum", 1e-6)
        rd, wr = scaled(r, "dram__bytes_read.sum", 1e9), scaled(r, "dram__bytes_write.sum", 1e9)
        lines.append(f"| {op} ({k}) | {t_us:.0f} | {g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {rd:.2f} | {wr:.2f} | "
                     f"{(rd + wr) * 1e3 / t_us:.2f} | {g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                     f"{g(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | {g(r, 'sm__cycles_elapsed.avg.per_second'):.2f} |")
        if i < len(names):
            cls = "fc_tcgen05" if "fc_tcgen05" in k else "conv_res_tcgen05" if "conv_res" in k else "stem" if "stem" in k else None
            if cls:
                e = traffic.setdefault(cls, {"launches_captured": 0, "bytes": 0.0})
                e["launches_captured"] += 1
                e["bytes"] += (rd + wr) * 1e9
    with open(out, "w") as f:
        f.write(f"# {title}\n\nCold-cache, serialised launches (compare shares).  `tensor %` = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed; "
                "DRAM TB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.\n\n")
        f.write("| launch | time us | tensor % | DRAM rd GB | DRAM wr GB | DRAM TB/s | LTS % | LSU wavefronts % | SM GHz |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        f.write("\n".join(lines) + "\n")
    tj = {"source": f"{os.path.basename(out)} (ncu --set full, stage-1 launches on {rows_per_launch} block rows)"}
    for cls, e in traffic.items():
        tj[cls] = {"launches_captured": e["launches_captured"], "dram_bytes_per_row_per_stage_forward": e["bytes"] / rows_per_launch,
                   "dram_bytes_per_launch": e["bytes"] / e["launches_captured"]}
    json.dump(tj, open(traffic_json, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), sys.argv[6] if len(sys.argv) > 6 else "ncu --set full")
