"""Summarise an .ncu-rep (read with `ncu -i` on the CPU box) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt ["note"]"""
import collections, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    for r in raw[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:72s} {r[i]} {units[i]}")
        lines.append("")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]))))
    tot, samp, text = collections.Counter(), collections.Counter(), {}
    stall = collections.Counter()
    cur_file, hdrs = None, None
    seen_fn = set()
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            fn = r[1]
        elif r[0] == "Line No":
            hdrs = r
            key_fn = (cur_file, fn)
            skip = key_fn in seen_fn      # each launch repeats the listing; keep the first
            seen_fn.add(key_fn)
        elif hdrs is not None and not skip:
            try:
                n = int(r[hdrs.index("Instructions Executed")]); sm = int(r[hdrs.index("# Samples")])
            except Exception:
                continue
            if not r[hdrs.index("Line No")].strip():
                continue                  # SASS rows of the mixed listing: already counted on their CUDA line
            k = (cur_file, r[hdrs.index("Line No")])
            tot[k] += n; samp[k] += sm; text[k] = r[hdrs.index("Source")][:100]
            for c in hdrs:
                if c.startswith("stall_") and "Not Issued" not in c:
                    try:
                        stall[c] += int(r[hdrs.index(c)])
                    except Exception:
                        pass
    T, S = sum(tot.values()) or 1, sum(samp.values()) or 1
    lines.append(f"top source lines of the first profiled launch (instructions {T}, samples {S})")
    for k, v in tot.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
        lines.append(f"  {k[0]:18s}:{k[1]:>5s}  inst {100*v/T:5.1f}%  samples {100*samp[k]/S:5.1f}%  {text[k]}")
    lines.append("")
    lines.append("warp stall reasons (all samples): " + ", ".join(f"{k[6:]} {100*v/max(sum(stall.values()),1):.1f}%" for k, v in stall.most_common(8)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
