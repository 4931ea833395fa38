"""DRAM traffic of the bench step from an ncu metrics CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch).

usage: python tools/traffic_report.py <ncu.csv> <steps profiled> <batch> <out.json> ["command that produced the csv"]

The CSV comes from `ncu --profile-from-start off --metrics ... --csv python bench.py --profile-steps N` (tools/probes/*): every kernel of
N whole pipeline steps (both conv stacks, pre/post-processing, tracker association), caches NOT flushed between launches
(--cache-control none), so a layer that finds its input in the 126 MB L2 is counted the way it runs in the bench.  The JSON this
writes is what bench.py reports as roofline.traffic (per launch of the conv/FC GEMM kernels) next to the algorithmic bytes of the
same launches (every operand and result touched exactly once: activations in, weights, activations out)."""
import csv, json, os, sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
GEMM_KERNELS = ("conv_gemm_v3_kernel", "conv_chain_v3_kernel", "fc_stream_kernel")


def parse(fn):
    launches = {}
    for r in csv.reader(open(fn, newline="")):
        if len(r) < 15 or not r[0].isdigit():
            continue
        d = launches.setdefault(int(r[0]), {"kernel": r[4].split("(")[0].split("<")[0].replace("void ", "").strip(), "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1.0)
    return [launches[k] for k in sorted(launches)]


def algorithmic_bytes(batch):
    """per step (one batch through YOLOv8l + UFLDv2-res34): bytes every GEMM launch must touch once -- fp16 activations in and out
    (interior pixels only), fp16 weights, fp32 outputs where the plan says so"""
    from adas_b200 import plan
    tot = 0
    n = 0
    for kind, build in (("yolov8", lambda W: plan.build_yolov8(W, "l")), ("ufldv2", lambda W: plan.build_ufldv2(W, "34"))):
        pb = build(plan.synth_weights(kind, 0))
        for typ, p, _ in pb.ops:
            if typ != 1:
                continue
            a_buf, _, Kc, ntaps, _, _, N, _, _, _, _, ob, _, _, tr = p[:15]
            s2 = p[16]
            ab, obf = pb.buffers[a_buf], pb.buffers[ob]
            def interior(b):            # (rows_per_img, C, dtype, H, W, _) -> interior pixels per image
                return (b[3] * b[4]) if b[3] > 0 else b[0]
            px_in, px_out = interior(ab), interior(obf)
            if tr:                      # FC: out[batch, N] = W[N, K] x[batch, K]
                tot += N * Kc * ntaps * 2 + batch * Kc * 2 + batch * N * (4 if obf[2] == 1 else 2)
            else:
                tot += batch * px_in * Kc * 2 + N * Kc * ntaps * 2 + batch * px_out * N * (4 if obf[2] == 1 else 2)
            n += 1
    return tot, n


def main():
    fn, steps, batch, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    cmd = sys.argv[5] if len(sys.argv) > 5 else ""
    L = parse(fn)
    by = {}
    for d in L:
        k = by.setdefault(d["kernel"], {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        k["launches"] += 1
        k["time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
        k["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
        k["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
    g = [v for k, v in by.items() if k in GEMM_KERNELS]
    gl = sum(v["launches"] for v in g)
    gb = sum(v["dram_read"] + v["dram_write"] for v in g)
    alg, n_ops = algorithmic_bytes(batch)
    all_b = sum(v["dram_read"] + v["dram_write"] for v in by.values())
    res = {
        "source": os.path.basename(fn), "command": cmd, "steps_profiled": steps, "batch": batch,
        "gemm_kernels": sorted(k for k in by if k in GEMM_KERNELS),
        "gemm_launches_per_step": gl / steps, "plan_gemm_ops_per_step": n_ops,
        "dram_bytes_per_launch": int(gb / max(gl, 1)), "gemm_dram_bytes_per_step": int(gb / steps),
        "algorithmic_bytes_per_step": int(alg), "algorithmic_bytes_per_launch": int(alg * steps / max(gl, 1)),
        "dram_over_algorithmic": round(gb / steps / alg, 3),
        "all_kernels_dram_bytes_per_step": int(all_b / steps),
        "per_kernel_per_step": {k: {"launches": v["launches"] / steps, "time_us_serialised": round(v["time_us"] / steps, 1),
                                    "dram_read_MB": round(v["dram_read"] / steps / 1e6, 2), "dram_write_MB": round(v["dram_write"] / steps / 1e6, 2)}
                                for k, v in sorted(by.items(), key=lambda x: -x[1]["time_us"])},
        "note": "ncu serialises the two networks' streams and replays each launch; caches are not flushed between launches (--cache-control none). "
                "Per-launch times are therefore not bench times; the byte counts are what the bench moves.",
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("gemm_launches_per_step", "dram_bytes_per_launch", "gemm_dram_bytes_per_step", "algorithmic_bytes_per_step", "dram_over_algorithmic", "all_kernels_dram_bytes_per_step")}))
    tot_t = sum(v["time_us"] for v in by.values())
    for k, v in sorted(by.items(), key=lambda x: -x[1]["time_us"])[:14]:
        print(f"  {k:34s} {v['launches'] / steps:6.1f}/step {v['time_us'] / steps:9.1f} us {100 * v['time_us'] / tot_t:5.1f}%  rd {v['dram_read'] / steps / 1e6:8.1f} MB  wr {v['dram_write'] / steps / 1e6:8.1f} MB")


if __name__ == "__main__":
    main()
