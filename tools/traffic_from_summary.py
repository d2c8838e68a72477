"""profiles/<tag>_traffic.json and <tag>_attention_pipe.json (what bench.py quotes as `roofline.traffic` and
`attention.tensor_pipe_pct_ncu`) from an ncu summary CSV made by tools/ncu_summary.py.
python tools/traffic_from_summary.py profiles/r02b_ncu_full_b148_fp16_summary.csv 148 fp16 r02b"""
import csv, json, os, sys

path, batch, precision, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
rows = list(csv.reader(open(path)))
hdr = rows[0]


def col(prefix):
    return next(i for i, h in enumerate(hdr) if h.startswith(prefix))


def to_bytes(v, unit_hdr):
    u = unit_hdr[unit_hdr.index("[") + 1:unit_hdr.index("]")].lower()
    return float(v) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


c_rd, c_wr, c_t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
c_tensor, c_xu, c_grid = col("sm__pipe_tensor_cycles_active"), col("sm__inst_executed_pipe_xu"), col("launch__grid_size")
src = f"{path} (ncu --set full --clock-control none, tools/profile_forward.py --batch {batch} --precision {precision})"
recs = [{"kernel": r[0], "bytes": to_bytes(r[c_rd], hdr[c_rd]) + to_bytes(r[c_wr], hdr[c_wr]), "ms": float(r[c_t]),
         "tensor": float(r[c_tensor]), "xu": float(r[c_xu]), "grid": int(float(r[c_grid]))} for r in rows[1:]]


def first(sub, pick=None):
    c = [r for r in recs if sub in r["kernel"]]
    return (max(c, key=pick) if pick else c[0]) if c else None


table = {}
for name, sub, pick, note in (
        ("bd_fps_grid", "fps_bucket_kernel", None, ""),
        ("bd_ball_query_grid_query", "bq_query_kernel", None, ""),
        ("bd_sa_mlp_tc", "sa_mlp_tc_kernel", lambda r: r["ms"], " — the SA2 launch"),
        ("bd_attention_tc", "attention_ws_kernel", lambda r: r["ms"], " — the 1024x1024 visual self-attention launch"),
        ("bd_linear_tc", "linear_tc_kernel<0, 0", lambda r: r["ms"], " — the largest plain launch"),
        ("bd_linear_ln_tc", "linear_tc_kernel<1, 0", lambda r: r["ms"], " — the largest linear + LayerNorm launch")):
    r = first(sub, pick)
    if r:
        table[name] = {"batch": batch, "precision": precision, "kernel": r["kernel"], "dram_bytes_per_launch": int(r["bytes"]),
                       "gpu_time_ms": r["ms"], "source": src + note}
att = sorted([r for r in recs if "attention_ws_kernel" in r["kernel"]], key=lambda r: -r["ms"])
pipe = {"source": src, "precision": precision, "batch": batch}
if att:
    pipe["attention_ws_kernel 1024x1024 (visual self-attention, K / V by tensor copy)"] = {
        "tensor_pipe_pct": att[0]["tensor"], "xu_pipe_pct": att[0]["xu"], "gpu_time_ms": att[0]["ms"]}
    seen, others = set(), []
    for r in att[1:]:
        key = (r["grid"], round(r["ms"], 2))
        if key not in seen:
            seen.add(key)
            others.append([r["ms"], r["tensor"], r["xu"]])
    pipe["other attention launches (gpu ms, tensor %, xu %)"] = others
    pipe["bf16x3 mode (round 1 capture, 32 scenes)"] = {"tensor_pipe_pct": 41.5,
                                                        "source": "profiles/r01f_ncu_full_attention_linear_b32_bf16x3_summary.csv"}
out = os.path.dirname(path)
json.dump(table, open(os.path.join(out, f"{tag}_traffic.json"), "w"), indent=1)
json.dump(pipe, open(os.path.join(out, f"{tag}_attention_pipe.json"), "w"), indent=1)
print(json.dumps(table, indent=1)[:600])
