"""Turn the raw captures in gpurun_out/ into the tracked summaries under profiles/.

Expects (from the gpurun command in profiles/README.md):
  gpurun_out/prof_bench_r1.ncu-rep   ncu --set full on `bench.py --steps 2 --warmup 3`
  gpurun_out/launches_r1.csv         ncu --metrics gpu__time_duration.sum launch list of the same command
  gpurun_out/bench_r1_final.json     the bench line, gpurun_out/bench_r1_reference.json the reference arm
"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

raw = "/tmp/raw_%s.csv" % tag
subprocess.run("ncu -i %s/prof_bench_%s.ncu-rep --page raw --csv > %s 2>/dev/null" % (G, tag, raw), shell=True, check=True)
with open(os.path.join(P, "%s_bench_ncu_full_summary.txt" % tag), "w") as f:
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), raw], stdout=f, check=True)

rows = list(csv.reader(open(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
def val(r, k): return float(r[idx[k]]) * scale[units[idx[k]]]
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
def val_t(r): return float(r[idx["gpu__time_duration.sum"]]) * tscale.get(units[idx["gpu__time_duration.sum"]], 1.0)
bench = json.load(open(os.path.join(G, "bench_%s_final.json" % tag)))
B = bench["config"]["batch_per_gpu"]
out = {"source": "ncu --set full --clock-control none on `bench.py --steps 2 --warmup 3` (batch %d x 3840x2160 RGB); "
                 "summary in profiles/%s_bench_ncu_full_summary.txt" % (B, tag), "kernels": []}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    e = {"kernel": name, "grid": int(r[idx["launch__grid_size"]]), "dram_read_bytes": val(r, "dram__bytes_read.sum"),
         "dram_write_bytes": val(r, "dram__bytes_write.sum"),
         "alu_pipe_pct": float(r[idx["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]),
         "issue_active_pct": float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
         "dram_pct": float(r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])}
    e["dram_bytes_per_launch"] = e["dram_read_bytes"] + e["dram_write_bytes"]
    out["kernels"].append(e)
    if "smsp__inst_executed.sum" in idx:
        e["warp_inst"] = float(r[idx["smsp__inst_executed.sum"]])
    e["duration_us"] = val_t(r)
    if "fwd_level_kernel<1, 0, 1>" in name:
        out["fwd_level0_bytes_per_launch"] = e["dram_bytes_per_launch"]
        out["fwd_level0_samples_per_launch"] = B * 3840 * 2160 * 3
        if "warp_inst" in e:  # thread-instructions per sample = 32 x warp instructions / samples (VERDICT r1 item 1's figure)
            out["fwd_level0_thread_inst_per_sample"] = round(32 * e["warp_inst"] / out["fwd_level0_samples_per_launch"], 2)
        out["fwd_level0_issue_active_pct"], out["fwd_level0_alu_pipe_pct"] = e["issue_active_pct"], e["alu_pipe_pct"]
        out["limiter"] = ("integer issue: ncu ALU pipe %.0f %%, issue slots %.0f %%, DRAM %.0f %% of peak (level-0 forward kernel)"
                          % (e["alu_pipe_pct"], e["issue_active_pct"], e["dram_pct"]))
    if "inv_level_kernel<1, 0, 2>" in name:
        out["inv_level0_bytes_per_launch"] = e["dram_bytes_per_launch"]
        out["inv_level0_issue_active_pct"], out["inv_level0_alu_pipe_pct"] = e["issue_active_pct"], e["alu_pipe_pct"]
        if "warp_inst" in e:
            out["inv_level0_thread_inst_per_sample"] = round(32 * e["warp_inst"] / (B * 3840 * 2160 * 3), 2)
json.dump(out, open(os.path.join(P, "traffic.json"), "w"), indent=1)

lrows = list(csv.reader(l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if l.startswith('"')))
lh = {h: i for i, h in enumerate(lrows[0])}
with open(os.path.join(P, "%s_bench_launches.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:level_kernel -c 60 on "
            "`bench.py --steps 2 --warmup 3` (batch %d x 4K RGB); cold-cache, serialised: compare shares, not absolutes\n" % B)
    f.write('"id","kernel","grid","block","duration_us"\n')
    for r in lrows[1:]:
        f.write('"%s","%s","%s","%s","%.1f"\n' % (r[lh["ID"]], r[lh["Kernel Name"]], r[lh["Grid Size"]], r[lh["Block Size"]],
                                               float(r[lh["Metric Value"]]) / 1e3))
for n in ("bench_%s_final.json" % tag, "bench_%s_reference.json" % tag):
    if os.path.exists(os.path.join(G, n)):
        shutil.copy(os.path.join(G, n), os.path.join(P, n))
for n in (1, 2, 4, 8):
    s = os.path.join(G, "scale_%s_n%d.json" % (tag, n))
    if os.path.exists(s):
        shutil.copy(s, os.path.join(P, "scale_%s_n%d.json" % (tag, n)))
print("fwd L0 traffic B/sample:", out["fwd_level0_bytes_per_launch"] / out["fwd_level0_samples_per_launch"], "|", out.get("limiter"))
