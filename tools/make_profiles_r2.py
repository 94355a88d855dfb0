"""Turn the scratch outputs of tools/gpu_profiles.sh (gpurun_out/r2p_*) into the tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, io, json, os, shutil, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, prof = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
P = os.path.join(out, "r2p")
TAG = "r02"
CONV = ("conv_igemm_kernel", "conv_win_kernel", "conv_ring_kernel", "conv_wgrad_kernel", "conv_wgrad_ring_kernel", "fold_conv_kernel",
        "fold_wgrad_kernel", "wgrad_finish", "slab_sum_kernel")


def short(name):
    n = name.replace("void ", "").replace("fv::", "")
    return n.split("(")[0][:70]


def launches():
    src = P + "_launches.csv"
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            per.setdefault(r[idc], {"name": r[kn]})[r[mn]] = float(r[mv].replace(",", ""))
    # keep the LAST complete step: the launches between the last two adam_multi launches
    ids = list(per.keys())
    adam = [i for i, k in enumerate(ids) if "adam_multi" in per[k]["name"]]
    sel = ids[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else ids
    agg = collections.OrderedDict()
    for k in sel:
        d = per[k]
        a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(prof, f"{TAG}_launches_summary.md"), "w") as fh:
        fh.write("# ncu launch list of one eager bench step (round-2 kernels)\n\n"
                 "`FACEVAE_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "-c 4000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-glue-roofline --profile-steps 1`; the table is the last "
                 f"complete train step of the list (between two `adam_multi_kernel` launches; {len(sel)} launches). ncu times are cold-cache "
                 f"and serialised: compare SHARES with `in_graph_kernels` of the bench line, not absolutes. Raw list: {TAG}_launches.csv\n\n"
                 "| kernel | launches | total us | share | DRAM read MB / launch | DRAM write MB / launch |\n|---|---:|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {n} | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / a[0] / 1e6:.1f} | {a[3] / a[0] / 1e6:.1f} |\n")
        conv = {n: a for n, a in agg.items() if any(c in n for c in CONV)}
        cl = sum(a[0] for a in conv.values())
        ct = sum(a[2] + a[3] for a in conv.values())
        fh.write(f"\nConvolution kernels (incl. weight-gradient finish passes): {cl} launches, {sum(a[1] for a in conv.values()) / 1e3:.1f} us "
                 f"= {100 * sum(a[1] for a in conv.values()) / tot:.1f}% of the step's kernel time; DRAM traffic {ct / 1e6:.1f} MB per step, "
                 f"{ct / cl / 1e6:.2f} MB per launch.\n")
    json.dump({"source": f"profiles/{TAG}_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum, last complete eager step)",
               "what": "average DRAM bytes (read + write) per convolution launch of one eager bench step (forward, data gradient, weight "
                       "gradient and the weight-gradient finish passes)",
               "launches": cl, "bytes_per_launch": ct / max(cl, 1), "conv_dram_bytes_per_step": ct,
               "per_kernel": {n: {"launches": a[0], "dram_bytes_per_launch": (a[2] + a[3]) / a[0], "time_us_per_launch": a[1] / a[0] / 1e3}
                              for n, a in conv.items()}},
              open(os.path.join(prof, f"{TAG}_conv_traffic.json"), "w"), indent=1)
    # the tracked raw list: the selected step only (the full list is ~4000 launches)
    with open(os.path.join(prof, f"{TAG}_launches.csv"), "w") as fh:
        w = csv.writer(fh)
        w.writerow(["ID", "Kernel Name", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum [B]", "dram__bytes_write.sum [B]"])
        for k in sel:
            d = per[k]
            w.writerow([k, d["name"][:160], d.get("gpu__time_duration.sum", ""), d.get("dram__bytes_read.sum", ""), d.get("dram__bytes_write.sum", "")])


KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg"]


def full(src=None, dst=None, note=""):
    src = src or P + "_full_raw.csv"
    dst = dst or os.path.join(prof, f"{TAG}_ncu_full_step.md")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src, errors="replace")))
    if len(rows) < 3 or "Kernel Name" not in rows[0]:
        return
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    cols = [(k, hdr.index(k)) for k in KEEP if k in hdr]
    fam = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) > kn:
            fam.setdefault(short(r[kn]), []).append(r)
    with open(dst, "w") as fh:
        fh.write("# `ncu --set full --clock-control none --import-source on` over ONE eager train step (batch 32, 256x256)\n\n" + note +
                 "Command: `STEPS=1 ncu ... python tools/ncu_one.py` (tools/gpu_profiles.sh). One row per launch, launch order within a kernel "
                 "family; times are ncu's (cold cache, serialised). `xbar2l1tex` = bytes delivered L2 -> SM; `tensor%` = "
                 "sm__pipe_tensor_cycles_active of active cycles.\n")
        for n, rs in fam.items():
            fh.write(f"\n## {n} ({len(rs)} launches)\n\n| # | grid | regs | time us | DRAM rd MB | DRAM wr MB | DRAM % | L2->SM MB | tensor % | L2 % | inst M |\n"
                     "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
            for i, r in enumerate(rs):
                g = lambda k: (r[hdr.index(k)].replace(",", "") if k in hdr else "")
                f = lambda k, s=1.0: (f"{float(g(k)) * s:.1f}" if g(k) not in ("", "n/a") else "")
                tu = units[hdr.index("gpu__time_duration.sum")]
                ts = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1e-3)
                def by(k):
                    u = units[hdr.index(k)] if k in hdr else "byte"
                    return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                def cnt(k):
                    u = units[hdr.index(k)] if k in hdr else "inst"
                    return 1e-6
                fh.write(f"| {i} | {g('launch__grid_size')} | {g('launch__registers_per_thread')} | {f('gpu__time_duration.sum', ts)} | "
                         f"{f('dram__bytes_read.sum', by('dram__bytes_read.sum'))} | {f('dram__bytes_write.sum', by('dram__bytes_write.sum'))} | "
                         f"{f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {f('l1tex__m_xbar2l1tex_read_bytes.sum', by('l1tex__m_xbar2l1tex_read_bytes.sum'))} | "
                         f"{f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | {f('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                         f"{f('smsp__inst_executed.sum', cnt('smsp__inst_executed.sum'))} |\n")


def source_pages(top=12):
    with open(os.path.join(prof, f"{TAG}_ncu_hot_lines.md"), "w") as fh:
        fh.write("# Hottest SASS lines (warp-state samples) of one launch per kernel family, from the `--set full --import-source on` capture\n")
        for fn in sorted(os.listdir(out)):
            if not (fn.startswith("r2p_src_") and fn.endswith(".csv")):
                continue
            rows = list(csv.reader(open(os.path.join(out, fn), errors="replace")))
            try:
                hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
            except StopIteration:
                continue
            hdr = rows[hi]
            ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            data = []
            for idx, r in enumerate(rows[hi + 1:]):
                if len(r) > isamp and r[isamp].isdigit():
                    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
                    data.append((int(r[isamp]), idx, r[ia].strip()[:80], int(r[iex] or 0), st))
            tot = sum(d[0] for d in data) or 1
            fh.write(f"\n## {fn[len('r2p_src_'):-4]} (total samples {tot})\n\n| share | SASS line | executed | top stall reasons |\n|---:|---|---:|---|\n")
            for s, idx, src, ex, st in sorted(data, reverse=True)[:top]:
                fh.write(f"| {100 * s / tot:.1f}% | `{src}` | {ex} | {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} |\n")


def copies():
    for a, b in (("_bench_1gpu.json", "_bench_1gpu.json"), ("_bench_reference_cpu.json", "_bench_reference_cpu.json"), ("_timeline.log", "_step_timeline.txt"),
                 ("_bench_512deep_1gpu.json", "_bench_512deep_1gpu.json"), ("_infer.jsonl", "_inference_sweep.jsonl")):
        if os.path.exists(P + a) and os.path.getsize(P + a) > 0:
            shutil.copy(P + a, os.path.join(prof, TAG + b))


if __name__ == "__main__":
    launches()
    full(note="Captured at commit 16c89dc (before the last round-2 optimisations: see r02_ncu_changed_kernels.md for the kernels that changed "
              "afterwards).\n\n")
    full(P + "_changed_raw.csv", os.path.join(prof, f"{TAG}_ncu_changed_kernels.md"),
         "Kernels changed after the full-step capture (filter-resident implicit GEMM, 4-slot ring for N = 128, pipelined pw_bwd_reduce, flat "
         "weight prep, wide slab sum), captured at the final commit of the round with `-k regex:...`.\n\n")
    source_pages()
    copies()
    print("profiles written")
