"""Builds profiles/traffic.json from ncu --set full summaries (tools/ncu_summary.py CSVs): per kernel the DRAM bytes per launch, the
issue-slot and lane utilisation, the tensor-pipe activity, and the hashes of the sources the capture was taken on.

    python tools/traffic_from_ncu.py --sha gpurun_out/source_sha.json --out profiles/traffic.json \
        k_env_rollout=profiles/r2_env_rollout_summary.csv k_nn_conv_tc3=profiles/r2_selfplay_summary.csv ...

Values are the mean over the captured launches of that kernel.  bench.py reports them as roofline.traffic / issue_roofline only while
the listed sources are unchanged."""
import argparse
import csv
import json

# which translation unit (and headers) a kernel is compiled from
SOURCES = {
    "k_env_rollout": ["az_env.cu", "az_game.cuh", "az_philox.h"],
    "k_mcts_sim": ["az_mcts.cu", "az_game.cuh", "az_philox.h"],
    "k_nn_conv_tc3": ["az_nn_tc.cu", "az_tc_ptx.cuh"],
    "k_nn_heads_tc": ["az_nn_tc.cu"],
    "k_nn_conv_tc": ["az_nn_tc.cu", "az_tc_ptx.cuh"],
}
COLS = {
    "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum", "duration": "gpu__time_duration.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "tensor_active_pct_of_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sha", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--launch", action="append", default=[], help="kernel=description of what one launch processes")
    ap.add_argument("pairs", nargs="+")
    a = ap.parse_args()
    sha = json.load(open(a.sha))
    launch = dict(x.split("=", 1) for x in a.launch)
    out = {"_comment": "per-launch ncu --set full numbers of the hot kernels (mean over the captured launches) and the sources they were "
                       "measured on; written by tools/traffic_from_ncu.py, read by bench.py (roofline.traffic, issue_roofline)"}
    for pair in a.pairs:
        kernel, path = pair.split("=", 1)
        rows = list(csv.reader(open(path)))
        hdr, data = rows[0], [r for r in rows[1:] if r and r[0].replace("void ", "").split("<")[0].strip() == kernel]
        if not data:
            raise SystemExit("no launch of %s in %s" % (kernel, path))
        ent = {"capture": path, "launches": len(data), "launch": launch.get(kernel, "")}
        for key, col in COLS.items():
            hit = [i for i, h in enumerate(hdr) if h.startswith(col + " [")]
            if not hit:
                continue
            unit = hdr[hit[0]].split("[")[1].rstrip("]")
            vals = [float(r[hit[0]].replace(",", "")) * SCALE.get(unit, 1.0) for r in data]
            ent[key + ("_ms" if key == "duration" else "")] = sum(vals) / len(vals)
        ent["sources"] = {f: sha[f] for f in SOURCES.get(kernel, []) if f in sha}
        out[kernel] = ent
    json.dump(out, open(a.out, "w"), indent=1)
    print("wrote", a.out, sorted(k for k in out if not k.startswith("_")))


if __name__ == "__main__":
    main()
