#!/usr/bin/env python3
"""Offline reading of an `ncu --set full --import-source on` report (no GPU needed): headline metrics, pipe
utilisation, stall reasons, the executed-instruction mix by opcode and the instructions that collected the most
stall samples. The opcode mix is what exposed the PRMT re-packing in the fp8 attention path (30 % of all
instructions) and the per-MMA replay loops of the first Marlin issuer.

    python tools/ncu_summary.py gpurun_out/attn_fp8_full.ncu-rep [--kernel-index 0] [--top 25]
"""
import argparse
import collections
import csv
import io
import subprocess


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


HEAD = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel-index", type=int, default=0)
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    raw = page(a.report, "raw")
    hdr, row = raw[0], raw[2 + a.kernel_index]
    d = dict(zip(hdr, row))
    print("== headline")
    for k in HEAD:
        if k in d:
            print(f"{k} = {d[k]}")
    print("== pipes (% of peak, active cycles)")
    for k in sorted(d):
        if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") or \
                k in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                      "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"):
            try:
                if float(d[k]) > 0.5:
                    print(f"{k} = {d[k]}")
            except ValueError:
                pass
    print("== stall reasons (warps per issue)")
    stalls = [(float(d[k]), k) for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for v, k in sorted(stalls, reverse=True)[:8]:
        print(f"{v:7.3f}  {k.split('stalled_')[1].split('_per_issue')[0]}")
    src = page(a.report, "source")
    if len(src) < 3:
        print("(no source page: capture with --import-source on)")
        return
    shdr = src[1]
    ix = {h: i for i, h in enumerate(shdr)}
    data = [r for r in src[2:] if len(r) == len(shdr)]
    mix, total = collections.Counter(), 0
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        parts = op.split(".")
        key = parts[0] + ("." + ".".join(parts[1:3]) if parts[0] in ("F2FP", "HMMA", "LDS", "STS", "LDG", "LDGSTS", "MUFU", "SYNCS", "UTCHMMA") and len(parts) > 1 else "")
        n = int(r[ix["Instructions Executed"]] or 0)
        mix[key] += n
        total += n
    print(f"== executed warp-instructions by opcode (total {total})")
    for k, v in mix.most_common(a.top):
        print(f"{100 * v / max(total, 1):5.1f}%  {v:12d}  {k}")
    print("== instructions with the most stall samples")
    samples = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: a.top]
    tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
    for r in samples:
        reasons = [(h[6:], int(r[ix[h]] or 0)) for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
        top = max(reasons, key=lambda x: x[1]) if reasons else ("", 0)
        print(f"{100 * int(r[ix['# Samples']] or 0) / max(tot_s, 1):5.1f}%  {top[0]:12s}  {r[ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main()
