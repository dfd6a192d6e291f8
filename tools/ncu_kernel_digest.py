"""Text digest of ONE launch of an `ncu --set full --import-source on` report: headline metrics, warp-state sample
breakdown, hottest SASS instructions, and the tcgen05 / TMA mnemonics present.

    python tools/ncu_kernel_digest.py <report.ncu-rep> <launch index> <out.txt> [title]
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, *args):
    txt = subprocess.run(["ncu", "-i", rep, *args, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def main():
    rep, idx, out_path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    title = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = ncu_csv(rep, "--page", "raw")
    rh, ru, row = raw[0], raw[1], raw[2 + idx]
    rc = {h: i for i, h in enumerate(rh)}
    src = ncu_csv(rep, "--page", "source", "--print-source", "sass", "-s", str(idx), "-c", "1")
    kname, head = src[0][1], src[1]
    col = {h: i for i, h in enumerate(head)}
    seen, body = set(), []
    for r in src[2:]:
        if len(r) == len(head) and r[col["# Samples"]].isdigit() and r[col["Address"]] not in seen:
            seen.add(r[col["Address"]])
            body.append(r)
    tot = sum(int(r[col["# Samples"]]) for r in body)
    stalls = [k for k in head if k.startswith("stall_") and "(Not" not in k]
    agg = {k: sum(int(r[col[k]]) for r in body if r[col[k]].isdigit()) for k in stalls}
    out = ["# ncu --set full digest" + (": " + title if title else ""), "# report: " + rep + ", launch " + str(idx), "kernel: " + kname[:140]]
    for k in METRICS:
        if k in rc:
            out.append(f"  {k:85s} {row[rc[k]]} {ru[rc[k]]}")
    out.append(f"warp-state samples: {tot}")
    for k, v in sorted(agg.items(), key=lambda t: -t[1])[:8]:
        out.append(f"  {k:28s} {v:7d}  {100 * v / max(tot, 1):5.1f}%")
    out.append("top instructions by samples (SASS, main stall reason):")
    for r in sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:14]:
        st = sorted(((int(r[col[k]]), k) for k in stalls if r[col[k]].isdigit() and int(r[col[k]]) > 0), reverse=True)
        out.append(f"  {int(r[col['# Samples']]):6d}  {r[col['Source']].strip()[:70]:70s} {st[0][1] if st else ''}")
    mn = lambda r: (r[col["Source"]].split()[1] if r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[0])
    special = sorted({mn(r) for r in body if any(t in r[col["Source"]] for t in ("UTC", "UTMA", "LDTM", "SYNCS"))})
    out.append("tcgen05 / TMA / mbarrier mnemonics present: " + ", ".join(special))
    open(out_path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
