"""profiles/traffic.json + a per-kernel markdown table from an `ncu --set full` report of ONE chain pass.
usage: python scratch/ncu_traffic.py <report.ncu-rep> <n_frames> <W> <H> <out_json> <out_md>"""
import csv, json, subprocess, sys

rep, n, W, H, out_json, out_md = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
TAGS = [("hist_frame_kernel", "hist_frame"), ("tilehist_kernel", "clahe_tilehist"), ("clahe_apply_kernel", "clahe_apply"),
        ("window15_kernel", "dz_window"), ("PipGF1a", "dz_gf1a"), ("PipGF1b", "dz_gf1b"), ("exposure_minmax_kernel", "dz_exposure_minmax"),
        ("splane_kernel", "dz_splane"), ("PipGF2a", "dz_gf2a"), ("PipGF2b", "dz_gf2b"), ("final_kernel", "dz_final")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return v * scale


px = float(n) * W * H
out, md = {}, ["| tag | kernel | time/frame us | DRAM read B/px | DRAM write B/px | DRAM % of peak | issue active % | warp-instr x32 /px | regs | warps active % |", "|---|---|---|---|---|---|---|---|---|---|"]
tot = 0.0
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    for key, tag in TAGS:
        if key in name:
            t = val(r, "gpu__time_duration.sum")
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            out[tag] = {"dram_bytes_per_px": (rd + wr) / px, "dram_read_bytes_per_px": rd / px, "dram_write_bytes_per_px": wr / px,
                        "time_s_per_frame_under_ncu": t / n}
            tot += t / n
            md.append("| %s | %s | %.1f | %.1f | %.1f | %s | %s | %.0f | %s | %s |" % (
                tag, name.split("(")[0][:40], 1e6 * t / n, rd / px, wr / px,
                r[col["FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed"]][:5], r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]][:5],
                32.0 * float(r[col["smsp__inst_executed.sum"]].replace(",", "")) / px, r[col["launch__registers_per_thread"]],
                r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]][:5]))
json.dump(out, open(out_json, "w"), indent=1, sort_keys=True)
md.append("")
md.append("Sum of the listed kernels under ncu: %.3f ms per frame (cold cache, serialised: compare shares)." % (1e3 * tot))
open(out_md, "w").write("\n".join(md) + "\n")
print("\n".join(md))
