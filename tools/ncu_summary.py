"""Turn an `ncu --set full --import-source on` report into the summaries committed under profiles/:
   python tools/ncu_summary.py gpurun_out/r2_full.ncu-rep profiles/r2   ->  profiles/r2_ncu_full_summary.json,
   profiles/r2_ncu_traffic.json, profiles/r2_ncu_source_hotspots.txt   (runs here: ncu only READS the report, no GPU needed)"""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(rep, prefix):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    summ, traffic, seen = [], {}, set()
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        d = {"kernel": name}
        for i, h in enumerate(hdr):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                d[h] = (r[i] + " " + units[i]).strip()
        summ.append(d)
        short = "basis_kernel" if "basis" in name else "transform_kernel" if "transform" in name else "invert_branches_kernel" if "invert" in name else name
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        traffic[short] = rd + wr; traffic[short + "_read"] = rd; traffic[short + "_write"] = wr
    traffic["what"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, tools/ncu_fill.py "
                       "(config 4, N = n = 8192, chopped, sparse store, steady state): report " + rep.split("/")[-1])
    json.dump(summ, open(prefix + "_ncu_full_summary.json", "w"), indent=1)
    json.dump(traffic, open(prefix + "_ncu_traffic.json", "w"), indent=1)
    # source page: instruction mix, stall samples, hottest instructions per kernel
    rows = ncu(rep, "source")
    with open(prefix + "_ncu_source_hotspots.txt", "w") as f:
        f.write("ncu --set full --import-source on, report %s (tools/ncu_fill.py, config 4 steady state), source page, per kernel\n" % rep.split("/")[-1])
        i, done = 0, set()
        while i < len(rows):
            if rows[i] and rows[i][0] == "Kernel Name":
                kname = rows[i][1]
                hdr = rows[i + 1]
                j = i + 2
                data = []
                while j < len(rows) and rows[j] and rows[j][0].startswith("0x"):
                    data.append(rows[j]); j += 1
                i = j
                if kname in done:
                    continue
                done.add(kname)
                si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
                sc = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
                tot = sum(int(r[si]) for r in data) or 1
                mix, stalls = {}, {}
                for r in data:
                    op = r[1].split()
                    op = [t for t in op if not t.startswith("@")][0].split(".")[0] if op else "?"
                    mix[op] = mix.get(op, 0) + int(r[ii])
                    for k in sc:
                        if r[k] not in ("", "0"):
                            stalls[hdr[k][6:]] = stalls.get(hdr[k][6:], 0) + int(r[k])
                ti = sum(mix.values()) or 1
                f.write("== %s\n   samples %d; executed warp instructions %d\n" % (kname[:90], tot, ti))
                f.write("   instruction mix (executed): " + ", ".join("%s %.1f%%" % (k, 100 * v / ti) for k, v in sorted(mix.items(), key=lambda x: -x[1])[:14]) + "\n")
                ts = sum(stalls.values()) or 1
                f.write("   stall samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / ts) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]) + "\n")
                f.write("   hottest instructions:\n")
                for r in sorted(data, key=lambda r: -int(r[si]))[:12]:
                    st = sorted(((hdr[k][6:], int(r[k])) for k in sc if r[k] not in ("", "0")), key=lambda x: -x[1])
                    f.write("     %5.2f%%  %-70s [%s]\n" % (100 * int(r[si]) / tot, r[1].strip()[:70], st[0][0] if st else ""))
            else:
                i += 1


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
