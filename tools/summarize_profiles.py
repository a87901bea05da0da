"""Turns the raw captures of tools/profile_round.sh (gpurun_out/<tag>_*) into the committed summaries under
profiles/:  python tools/summarize_profiles.py <tag>   (runs here, no GPU: `ncu -i` only reads the reports)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__cluster_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val) * mul


def ncu_summary(tag, name, header_lines):
    rep = os.path.join(GO, "%s_%s.ncu-rep" % (tag, name))
    hdr, units, rows = raw_rows(rep)
    lines = list(header_lines)
    res = []
    for r in rows:
        lines.append("")
        lines.append("%-80s %s" % ("Kernel Name", r[hdr.index("Kernel Name")]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("%-80s %s %s" % (k, r[i], units[i]))
        res.append({k: (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr})
    return lines, res


def launches(tag):
    path = os.path.join(GO, "%s_launches.csv" % tag)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        k = r[4]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6
    tot = sum(v[1] for v in agg.values())
    lines = ["# %s: ncu launch list of `python bench.py --steps 1 --warmup 0 --batch 4 --chunk 4 --no-cpu-baseline`" % tag,
             "# (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)",
             "# the run executes the step 3x (timed step, e2e warm-up, e2e step); total kernel time %.1f ms" % tot, "",
             "%-100s %6s %11s %7s" % ("kernel", "n", "ms", "share")]
    conv = 0.0
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append("%-100s %6d %11.3f %6.1f%%" % (k[:100], n, ms, 100 * ms / tot))
    for k, (n, ms) in agg.items():
        if "k_conv3x3_" in k:
            conv += ms
    lines += ["", "all tcgen05 conv launches: %.1f%% of kernel time" % (100 * conv / tot)]
    shutil.copy(path, os.path.join(PR, "%s_launches_batch4.csv" % tag))
    return lines


def bench_workloads():
    sys.path.insert(0, ROOT)
    import bench
    return bench.WORKLOADS


def main():
    tag = sys.argv[1]
    for f in ("bench_n1.json", "reference_arm.json", "clocks.csv"):
        src = os.path.join(GO, "%s_%s" % (tag, f))
        if os.path.exists(src):
            dst = {"bench_n1.json": "%s_bench_n1_batch32.json", "reference_arm.json": "%s_bench_reference_arm.json",
                   "clocks.csv": "%s_clocks_during_bench.csv"}[f] % tag
            shutil.copy(src, os.path.join(PR, dst))
    open(os.path.join(PR, "%s_launch_shares.txt" % tag), "w").write("\n".join(launches(tag)) + "\n")
    lines, res = ncu_summary(tag, "tower", [
        "# %s: ncu --set full --clock-control none -k regex:^k_conv3x3_tc2$ -s 1 -c 1  (bench.py --batch 2 --chunk 2)" % tag,
        "# kernel: tc::k_conv3x3_tc2<64,HIDDEN,...> = 256->256 tower conv on CTA pairs (tcgen05 cta_group::2), fp16x3 split",
        "# launch = hidden tower conv over the live maps of 2 images x 30 samples x 2 passes of 96x160x256"])
    r = res[0]
    rd, wr = to_bytes(*r["dram__bytes_read.sum"]), to_bytes(*r["dram__bytes_write.sum"])
    bj = json.loads(open(os.path.join(GO, "%s_bench_n1.json" % tag)).read().strip().splitlines()[-1])
    live_maps = 2 * (29 * 2)                      # class tower, skip_unread: (N-1)*passes live maps per image
    alg = 2.0 * live_maps * 96 * 160 * 256 * 4
    lines += ["", "algorithmic bytes of this launch: read %d x 15360 x 256 x 4 B + the same written = %.3f GB" % (live_maps, alg / 1e9),
              "measured DRAM traffic: %.3f GB read + %.3f GB written = %.3f GB  -> traffic / algorithmic = %.3f"
              % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, (rd + wr) / alg)]
    open(os.path.join(PR, "%s_conv_pair_ncu_full.txt" % tag), "w").write("\n".join(lines) + "\n")
    json.dump({"source": "profiles/%s_conv_pair_ncu_full.txt (ncu --set full: %d maps of 96x160x256, hidden tower conv on CTA pairs)" % (tag, live_maps),
               "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": alg, "maps": live_maps, "H": 96, "W": 160},
              open(os.path.join(PR, "conv_traffic.json"), "w"), indent=1)
    lines, _ = ncu_summary(tag, "out", [
        "# %s: ncu --set full --clock-control none -k regex:^k_conv3x3_wt$ -c 3  (bench.py --batch 2 --chunk 2)" % tag,
        "# kernels: the output convolutions at P3 (cls_score 63, cls_var 63, bbox_pred 36 channels, all padded to 64 rows), single CTA,",
        "# weights-as-A: stacked [w_hi; w_lo] as the M=128 operand, 16x16 pixels as N=256, 2 MMAs per K-step"])
    open(os.path.join(PR, "%s_out_conv_ncu_full.txt" % tag), "w").write("\n".join(lines) + "\n")
    side = os.path.join(GO, "%s_side_workloads.jsonl" % tag)
    if os.path.exists(side):
        out = ["# %s side workloads on ONE B200, batch 32 per step (python bench.py --workload <w> --no-cpu-baseline); not the headline metric" % tag,
               "%-14s %9s %10s %9s %11s %9s %7s" % ("workload", "images/s", "e2e img/s", "mma_frac", "conv share", "launches", "SM MHz")]
        for ln in open(side):
            ln = ln.strip()
            if not ln.startswith("{"):
                continue
            j = json.loads(ln)
            r = j.get("roofline") or {}
            wl = j["config"]["workload"]
            name = [k for k, v in bench_workloads().items() if wl.startswith(v[0])]
            out.append("%-14s %9.1f %10.1f %9.3f %11.3f %9d %7s   # %s" % (name[0] if name else "?", j["value"], j["e2e"]["value"],
                       r.get("mma_frac", float("nan")), r.get("conv_share_of_step", float("nan")), j["gpu_launches"],
                       j["clocks"]["sm_mhz"], wl))
        open(os.path.join(PR, "%s_side_workloads.txt" % tag), "w").write("\n".join(out) + "\n")
    print("value", bj["value"], "e2e", bj["e2e"]["value"], "roofline", bj["roofline"]["frac"], bj["roofline"]["mma_frac"])


if __name__ == "__main__":
    main()
