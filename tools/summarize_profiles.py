"""Turns the raw captures of tools/profile_round.sh (gpurun_out/<tag>_*) into the committed summaries under
profiles/:  python tools/summarize_profiles.py <tag>   (runs here, no GPU: `ncu -i` only reads the reports)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
# on the GPU box the summaries go to gpurun_out/<tag>_profiles (only gpurun_out/ travels back, <= 64 MiB: the .ncu-rep
# files are summarised there and deleted); here they are then copied into the tracked profiles/
PR = os.environ.get("POD_PROFILE_OUT", os.path.join(ROOT, "profiles"))
os.makedirs(PR, exist_ok=True)
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


def ncu_table(tag, name, title_lines, keep=None):
    """Generic per-kernel summary of an ncu --set full report (one block per captured launch)."""
    rep = os.path.join(GO, "%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        return None
    hdr, units, rows = raw_rows(rep)
    lines = list(title_lines)
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        if keep and not any(k in kn for k in keep):
            continue
        lines.append("")
        lines.append("%-80s %s" % ("Kernel Name", kn))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append("%-80s %s %s" % (k, r[i], units[i]))
        rd, wr, t = r[hdr.index("dram__bytes_read.sum")], r[hdr.index("dram__bytes_write.sum")], r[hdr.index("gpu__time_duration.sum")]
        try:
            b = to_bytes(rd, units[hdr.index("dram__bytes_read.sum")]) + to_bytes(wr, units[hdr.index("dram__bytes_write.sum")])
            tu = units[hdr.index("gpu__time_duration.sum")]
            sec = float(t) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
            lines.append("%-80s %.1f GB/s" % ("=> DRAM read+write / duration", b / sec / 1e9))
        except Exception:  # noqa: BLE001
            pass
    return lines


def main():
    tag = sys.argv[1]
    for f in ("bench_n1.json", "reference_arm.json", "clocks.csv"):
        src = os.path.join(GO, "%s_%s" % (tag, f))
        if os.path.exists(src):
            dst = {"bench_n1.json": "%s_bench_n1_batch32.json", "reference_arm.json": "%s_bench_reference_arm.json",
                   "clocks.csv": "%s_clocks_during_bench.csv"}[f] % tag
            shutil.copy(src, os.path.join(PR, dst))
    for f in ("sanitizer_memcheck.txt", "sanitizer_racecheck.txt", "sanitizer_synccheck.txt", "bayesod_envelope.txt", "mma_shape.txt",
              "q1_experiments.txt", "tower_clock.txt", "pytest_gpu.txt"):
        src = os.path.join(GO, "%s_%s" % (tag, f))
        if os.path.exists(src):
            shutil.copy(src, os.path.join(PR, "%s_%s" % (tag, f)))
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
        "# kernels: the output convolutions at P3 (weights-as-A: stacked [w_hi; w_lo] as the M=128 operand, 16x16 pixels as N=256,",
        "# 2 MMAs per K-step); with the fused sample mean cls_score / cls_var run on ONE mean map per image, bbox_pred per sample"])
    open(os.path.join(PR, "%s_out_conv_ncu_full.txt" % tag), "w").write("\n".join(lines) + "\n")
    for name, fn, title in (
            ("hbm", "hbm_kernels_ncu.txt", ["# %s: ncu --set full -k regex:^(k_q1_mean_act|k_mask_expand|k_sample_mean_q1_v4)$ -c 3 (bench.py --batch 4 --chunk 4)" % tag,
                                            "# the HBM-bound streaming kernels of the step (first launches = P3, class tower)"]),
            ("backbone", "backbone_ncu.txt", ["# %s: ncu --set full -k regex:^(k_conv3x3_tc|k_stem_conv7|k_maxpool3s2)$ -s 2 -c 8" % tag,
                                              "# (bench.py --workload loss_att --from-images --batch 4): stem + first bottleneck convolutions of the ResNet-50-FPN backbone"]),
            ("post", "post_kernels_ncu.txt", ["# %s: ncu --set full -k regex:^(k_decode|k_nms_fuse|k_cluster_merge|k_scores|k_topk)$ -c 10" % tag,
                                              "# (bench.py --workload mc_post --n-mc 30 --batch 4): the post-processing kernels of the post-NMS merge mode (120 runs per launch)"])):
        t = ncu_table(tag, name, title)
        if t:
            open(os.path.join(PR, "%s_%s" % (tag, fn)), "w").write("\n".join(t) + "\n")
    side = os.path.join(GO, "%s_side_workloads.jsonl" % tag)
    if os.path.exists(side):
        out = ["# %s side workloads on ONE B200 (python bench.py --workload <w> ... --no-cpu-baseline); not the headline metric" % tag,
               "%9s %10s %9s %9s %7s  %s" % ("images/s", "e2e img/s", "ms/step", "launches", "SM MHz", "workload")]
        for ln in open(side):
            ln = ln.strip()
            if not ln.startswith("{"):
                continue
            j = json.loads(ln)
            out.append("%9.1f %10.1f %9.2f %9d %7s  %s%s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["gpu_launches"],
                       j["clocks"]["sm_mhz"], j["config"]["workload"], " [CUDA graph]" if j["config"].get("cuda_graph") else ""))
            by = (j.get("roofline") or {}).get("ms_per_step_by_kernel")
            if by:
                out.append("%49s per-step kernel ms: %s" % ("", by))
        open(os.path.join(PR, "%s_side_workloads.txt" % tag), "w").write("\n".join(out) + "\n")
    print("value", bj["value"], "e2e", bj["e2e"]["value"], "roofline", bj["roofline"]["frac"], bj["roofline"]["mma_frac"])


if __name__ == "__main__":
    main()
