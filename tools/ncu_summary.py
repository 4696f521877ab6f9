#!/usr/bin/env python
"""Summarises an .ncu-rep (one kernel launch) into the text kept under profiles/: duration, DRAM bytes, throughput
percentages, occupancy, registers, L1 data-pipe wavefronts and the warp-stall breakdown (per issued instruction and
per sampled instruction class).  Usage: tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import collections
import csv
import json
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    # global stores: what the output layouts of the skinning kernel differ in (wavefronts, requests, sectors per request)
    "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
    "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print(f"# ncu summary of {rep.split('/')[-1]}  (ncu --set full --clock-control none --import-source on, one launch)")
    print(f"kernel: {m.get('Kernel Name', ('?',))[0]}")
    for k in RAW_KEYS:
        if k in m:
            print(f"{k:80s} {m[k][0]:>18s} {m[k][1]}")
    print("\n# warp stall reasons (average warps stalled per issued instruction)")
    stalls = [(h, float(v[0])) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for h, v in sorted(stalls, key=lambda x: -x[1]):
        print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:8.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 2:
        h2 = srows[1]
        ix = {h: i for i, h in enumerate(h2)}
        data = srows[2:]
        total = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data) or 1
        print(f"\n# source page: {total} stall samples, {inst} warp instructions executed")
        ops, samp = collections.Counter(), collections.Counter()
        for r in data:
            s = r[ix["Source"]].split()
            if not s:
                continue
            op = (s[1] if s[0].startswith("@") else s[0])
            op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "STG", "LDG", "SYNCS", "UBLKCP")) else op.split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]] or 0)
            samp[op] += int(r[ix["# Samples"]] or 0)
        print("  opcode            instr%  samples%")
        for op, c in ops.most_common(18):
            print(f"  {op:16s} {100 * c / inst:6.1f}  {100 * samp[op] / total:6.1f}")
        print("  top stalled instructions:")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:8]:
            reasons = {k: int(r[ix[k]] or 0) for k in h2 if k.startswith("stall_") and "(Not Issued)" not in k}
            top = max(reasons, key=reasons.get)
            print(f"    {100 * int(r[ix['# Samples']] or 0) / total:5.1f}%  {r[ix['Source']][:60]:60s} {top}")
        for col in ("L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Excessive"):
            if col in ix:
                print(f"  {col}: {sum(int(r[ix[col]] or 0) for r in data)}")
        if "L1 Wavefronts Shared" in ix:
            wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data)
            ex = sum(int(r[ix["L1 Wavefronts Shared Excessive"]] or 0) for r in data)
            print(f"  shared-memory wavefronts: {wf} (excessive / bank-conflict: {ex})")
    if len(sys.argv) > 2:   # optional traffic json for bench.py
        rd, wr = float(m["dram__bytes_read.sum"][0]), float(m["dram__bytes_write.sum"][0])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        total_bytes = rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]
        json.dump({"kernel": m["Kernel Name"][0], "dram_bytes_per_launch": total_bytes, "characters": int(sys.argv[3]),
                   "vertices": int(sys.argv[4]), "source": rep.split("/")[-1]}, open(sys.argv[2], "w"))


if __name__ == "__main__":
    main()
