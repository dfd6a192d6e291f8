"""Summarises an `ncu --set full` report of tools/forward_once.py (one launch per UNet layer) into JSON:
per-launch duration, tensor-pipe activity, DRAM bytes, plus the totals bench.py quotes as roofline.traffic.

    python tools/ncu_summary.py <report.ncu-rep> <out.json> [batch] [layer names, comma separated]
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else None
    names = sys.argv[4].split(",") if len(sys.argv) > 4 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}

    def val(r, key, kind):
        i = col[key]
        v = float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
        u = units[i]
        if kind == "bytes":
            return v * UNIT.get(u, 1.0)
        if kind == "ms":
            return v * UNIT.get(u, 1.0)
        return v

    layers = []
    for k, r in enumerate(body):
        layers.append({
            "layer": names[k] if names and k < len(names) else None,
            "kernel": r[col["Kernel Name"]],
            "ms": val(r, "gpu__time_duration.sum", "ms"),
            "dram_read_bytes": val(r, "dram__bytes_read.sum", "bytes"),
            "dram_write_bytes": val(r, "dram__bytes_write.sum", "bytes"),
            "pipe_tc_active_pct": val(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "pct")
            if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
            "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", "pct"),
            "dram_throughput_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "pct")
            if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
            "registers": val(r, "launch__registers_per_thread", "n"),
        })
    tc = [l for l in layers if "first_conv" not in l["kernel"]]
    summary = {
        "batch": batch,
        "note": "ncu replays every kernel cold-cache and serialised: durations here are not bench numbers; the DRAM bytes are what "
                "bench.py reports as roofline.traffic (sum over the tcgen05 launches of one forward pass)",
        "tcgen05_launches": len(tc),
        "tcgen05_dram_bytes": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in tc),
        "all_conv_dram_bytes": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in layers),
        "tcgen05_ms_under_ncu": sum(l["ms"] for l in tc),
        "layers": layers,
    }
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "layers"}))


if __name__ == "__main__":
    main()
