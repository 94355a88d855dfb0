"""Turn the scratch outputs of a gpurun validation run (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, io, json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, prof = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01_final"

def launches():
    src = os.path.join(out, "launches_final.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > mv:
            per.setdefault(r[idc], {"name": r[kn]})[r[mn]] = float(r[mv].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"][:110], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(prof, f"{tag}_launches_summary.md"), "w") as fh:
        fh.write("# ncu launch list of the eager bench step (final kernels of round 1)\n\n"
                 "`FACEVAE_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "-s 900 -c 520 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 1` (about two steps; cold-cache, "
                 f"serialised: compare shares, not absolutes). Raw list: {tag}_launches.csv\n\n"
                 "| kernel | launches | total us | share | DRAM read MB / launch | DRAM write MB / launch |\n|---|---:|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| {n} | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / a[0] / 1e6:.1f} | {a[3] / a[0] / 1e6:.1f} |\n")
    shutil.copy(src, os.path.join(prof, f"{tag}_launches.csv"))

def ncu_report(rep, dst, title, n_kernels=3, top=14):
    path = os.path.join(out, rep)
    if not os.path.exists(path):
        return
    keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(prof, dst), "w") as fh:
        fh.write(title + "\n")
        for r in rows[2:2 + n_kernels * 3]:
            fh.write("\nkernel: " + r[hdr.index("Kernel Name")][:100] + "\n")
            for k in keep:
                if k in hdr:
                    fh.write(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}\n")
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        try:
            hi = next(i for i, r in enumerate(srows) if "Source" in r and "# Samples" in r)
        except StopIteration:
            return
        h = srows[hi]
        ia, isamp = h.index("Source"), h.index("# Samples")
        stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        data = []
        for idx, r in enumerate(srows[hi + 1:]):
            if len(r) > isamp and r[isamp].isdigit() and int(r[isamp]):
                st = sorted(((int(r[i] or 0), h[i]) for i in stall), reverse=True)[:2]
                data.append((int(r[isamp]), idx, r[ia].strip()[:80], st))
        tot = sum(d[0] for d in data) or 1
        fh.write(f"\nfirst launch, top SASS lines by samples ({tot} samples):\n")
        for s_, idx, text, st in sorted(data, reverse=True)[:top]:
            fh.write(f"  {100 * s_ / tot:5.1f}% #{idx:4d} {text:80s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}\n")

def copy(src, dst):
    p = os.path.join(out, src)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(prof, dst))

launches()
ncu_report("prof_igemm.ncu-rep", f"{tag}_ncu_conv_igemm.txt",
           "ncu --set full of conv_igemm_kernel<64> inside the eager bench step (launches: enc.2 fwd 64->128 @128^2, enc.3 fwd 128->256 @64^2, enc.4 fwd 256->32 @32^2)")
ncu_report("prof_fold.ncu-rep", f"{tag}_ncu_fold_outconv.txt",
           "ncu --set full of the tap-folded out_conv kernels (tools/conv_bench.py --only out; capture taken BEFORE the two-issuer / cp.async changes: 91 / 104 / 102 us)",
           n_kernels=3)
copy("bench.json", f"{tag}_bench_1gpu.json")
copy("bench_ref.json", f"{tag}_bench_reference_cpu.json")
copy("bench_2gpu.json", f"{tag}_bench_2gpu.json")
copy("bench_8gpu.json", f"{tag}_bench_8gpu.json")
copy("timeline3.txt", f"{tag}_step_timeline.txt")
copy("torch_gpu_baseline.json", f"{tag}_torch_cudnn_baseline_b200.json")
copy("bench_512deep.json", f"{tag}_bench_512deep_1gpu.json")
copy("infer_sweep2.txt", f"{tag}_inference_sweep.txt")
copy("trace_outconv2.txt", f"{tag}_trace_outconv.txt")
for i, name in enumerate(sorted(f for f in os.listdir(out) if f.startswith("fold_experiments"))):
    with open(os.path.join(prof, f"{tag}_fold_experiments.txt"), "a" if i else "w") as fh:
        fh.write(f"== {name}\n" + open(os.path.join(out, name)).read())
print(sorted(f for f in os.listdir(prof) if f.startswith(tag)))
