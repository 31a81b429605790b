#!/usr/bin/env python
"""Turns the files a tools/profile_round.sh session left in gpurun_out/ into the tracked summaries under profiles/:
   profiles/<tag>_summary.md     bench line digest, ncu launch list (share of the step per kernel), per-kernel ncu --set full digest
   profiles/<tag>_launches.csv   the raw ncu launch list
   profiles/<tag>_full_<kernel>.raw.csv   the raw metric pages
   profiles/traffic.json         DRAM bytes per DOF of one CG iteration (read by bench.py for roofline.traffic)
usage: tools/summarize_profiles.py TAG [--full-dofs N]   (N = DOFs per group of the mesh the --set full captures ran on)
Also extracts the SASS evidence (UBLKCP / SYNCS / LDGSTS counts of the x-row kernel) from the built library with cuobjdump."""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("launch__registers_per_thread", "registers/thread"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps/issue)"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier (warps/issue)"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %")]


def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * f.get(u, 1)


def to_ms(v, u):
    f = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}
    return float(v.replace(",", "")) * f.get(u, 1)


out = [f"# {tag}: profile of the bench workload", ""]
bench = None
bl = os.path.join(G, f"{tag}_bench.log")
if os.path.exists(bl):
    for line in open(bl):
        if line.startswith("{"):
            bench = json.loads(line)
if bench:
    r = bench["roofline"]
    out += ["## bench line (python bench.py, not under a profiler)", "",
            f"* workload: {bench['config']['workload']}; n_phi per group = {bench['n_phi_per_group']}",
            f"* value = {bench['value']:.2f} {bench['unit']} (device time, inputs resident), e2e = {bench['e2e']['value']:.2f} {bench['unit']} "
            f"({bench['e2e']['seconds']:.1f} s wall incl. {bench['e2e']['h2d_bytes_per_step'] * bench['steps'] / 1e9:.1f} GB H2D, "
            f"{bench['e2e']['d2h_bytes_per_step'] * bench['steps'] / 1e9:.1f} GB D2H)",
            f"* one CG iteration = {r['ms_per_launch']:.3f} ms -> {r['achieved']:.0f} GB/s algorithmic ({r['algorithmic_bytes_per_dof']:.0f} B/DOF) = "
            f"**{100 * r['frac']:.1f} %** of {r['peak']:.0f} GB/s ({r['peak_source']})",
            f"* kernels (CUDA events, ms): " + ", ".join(f"{k} {v:.3f}" for k, v in r["kernels_ms"].items() if v and k not in ("path",)),
            f"* clocks: {bench['clocks']}", f"* cpu_baseline: {bench.get('cpu_baseline')}", ""]
ll = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            v = to_ms(d["Metric Value"], d["Metric Unit"])
        except ValueError:
            continue
        a = agg.setdefault(d["Kernel Name"].split("(")[0][:60], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## ncu launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none -s 200 -c 400`)", "",
            "| kernel | launches | total ms | avg ms | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.3f} | {t / n:.4f} | {100 * t / tot:.1f} % |")
    out += ["", "Per-launch times under ncu are cold-cache and serialised: compare the shares with the CUDA-event timings above.", ""]
    shutil.copy(ll, os.path.join(P, f"{tag}_launches.csv"))
full_dofs = None
if "--full-dofs" in sys.argv:
    full_dofs = float(sys.argv[sys.argv.index("--full-dofs") + 1])
elif bench:
    full_dofs = float(bench["n_phi_per_group"])
tot_bytes = 0.0
tot_ms = 0.0
table = []
for k in ("k_xrow", "k_ycol", "k_zfwd2", "k_zback2"):
    f = os.path.join(G, f"{tag}_full_{k}.raw.csv")
    if not os.path.exists(f):
        continue
    rows = list(csv.reader(open(f)))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))
    rd, wr = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]), to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    ms = to_ms(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
    tot_bytes += rd + wr
    tot_ms += ms
    table.append((k, d.get("Kernel Name", k)[:48], ms, rd, wr, d, u))
    shutil.copy(f, os.path.join(P, f"{tag}_full_{k}.raw.csv"))
if table:
    out += [f"## ncu --set full, one launch per kernel ({full_dofs:.0f} DOFs per launch)", "",
            "| kernel | ms | DRAM read GB | DRAM written GB | B/DOF | DRAM GB/s | warp instr / DOF | issue busy % | regs | long-scoreboard stall |",
            "|---|---|---|---|---|---|---|---|---|---|"]
    for k, name, ms, rd, wr, d, u in table:
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        out.append(f"| `{name}` | {ms:.3f} | {rd / 1e9:.2f} | {wr / 1e9:.2f} | {(rd + wr) / full_dofs:.1f} | {(rd + wr) / ms / 1e6:.0f} | "
                   f"{inst * 32 / full_dofs:.0f} thr-instr | {float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']):.1f} | "
                   f"{d['launch__registers_per_thread']} | {float(d['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']):.1f} |")
    out += ["", f"One CG iteration: {tot_bytes / 1e9:.1f} GB of DRAM traffic = **{tot_bytes / full_dofs:.1f} B/DOF** (algorithmic: 90 B/DOF), "
                f"{tot_ms:.3f} ms under ncu.", ""]
    with open(os.path.join(P, "traffic.json"), "w") as fh:
        json.dump({"tag": tag, "path": 3, "dofs_per_launch": full_dofs, "dram_bytes_per_dof_per_cg_iteration": tot_bytes / full_dofs,
                   "kernels": {k: {"ms": ms, "dram_read": rd, "dram_write": wr} for k, _, ms, rd, wr, _, _ in table}}, fh, indent=1)
try:
    import re
    import subprocess
    lib = os.path.join(ROOT, "neutfem_b200", "lib", "libneutfem_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "k_xrowILi1ELi2ELi1ELi17ELi17" in cur:
            for op in ("UBLKCP", "SYNCS", "LDGSTS", "DFMA", "SHFL"):
                if re.search(r"\b" + op, line):
                    counts[op] = counts.get(op, 0) + 1
    if counts:
        out += ["## SASS of the bench variant of k_xrow (cuobjdump -sass, sm_100a cubin)", "",
                "`cp.async.bulk` -> `UBLKCP`, `mbarrier.*` -> `SYNCS`, per-lane `cp.async` (line factors) -> `LDGSTS`:", "",
                ", ".join(f"{k}: {v}" for k, v in counts.items()), ""]
except Exception as e:      # cuobjdump missing: skip the section
    out += [f"(SASS section skipped: {e})", ""]
with open(os.path.join(P, f"{tag}_summary.md"), "w") as fh:
    fh.write("\n".join(out) + "\n")
print("\n".join(out))
