"""Summarises an .ncu-rep (ncu --set full capture brought back under gpurun_out/) into the small text/JSON
files that are committed under profiles/.   python profiles/summarize_ncu.py <rep> <out-prefix> <ray_samples>"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_branch.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def main():
    rep, prefix, rays = sys.argv[1], sys.argv[2], float(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[h] = {"value": vals[i], "unit": units[i]}
            if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                d.setdefault("stalls_pct_of_warp_active", {})[h.split("issue_stalled_")[1].replace("_per_warp_active.pct", "")] = float(vals[i] or 0)
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                # warps of this scheduler waiting for that reason, per issued instruction
                d.setdefault("stalled_warps_per_issue", {})[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(vals[i] or 0)
        def num(k):
            v, u = float(d[k]["value"].replace(",", "")), d[k]["unit"]
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)
            return v * mult
        dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        t = num("gpu__time_duration.sum")
        d["derived"] = {
            "ray_samples": rays, "dram_bytes": dram, "dram_bytes_per_ray_sample": dram / rays, "dram_GBps": dram / t / 1e9,
            "ray_samples_per_s_under_ncu": rays / t,
            "l2_to_l1_bytes_per_ray_sample": float(d["lts__t_sectors_srcunit_tex_op_read.sum"]["value"]) * 32 / rays,
            "l1_global_load_bytes_per_ray_sample": float(d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]["value"]) * 32 / rays,
            "warp_inst_per_ray_sample": float(d["smsp__inst_executed.sum"]["value"]) / rays,
            "thread_inst_per_ray_sample": float(d["smsp__inst_executed.sum"]["value"]) * float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"]) / rays,
        }
        if "stalled_warps_per_issue" in d:
            d["stalled_warps_per_issue"] = dict(sorted(d["stalled_warps_per_issue"].items(), key=lambda kv: -kv[1])[:8])
        if "stalls_pct_of_warp_active" in d:
            d["stalls_pct_of_warp_active"] = dict(sorted(d["stalls_pct_of_warp_active"].items(), key=lambda kv: -kv[1])[:8])
        out.append(d)
    json.dump(out, open(prefix + ".json", "w"), indent=1)
    for d in out:
        print(d["kernel"][:80])
        for k, v in d.items():
            if isinstance(v, dict) and "value" in v:
                print("  %-70s %s %s" % (k, v["value"], v["unit"]))
        print("  derived:", json.dumps(d["derived"]))
        print("  stalled warps per issued instruction:", json.dumps(d.get("stalled_warps_per_issue")))


if __name__ == "__main__":
    main()
