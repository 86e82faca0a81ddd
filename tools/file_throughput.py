"""BASELINE configs[3] from files: N synthetic 1024x768 PNG pages through the CLI (decode -> scale -> net -> integer
post-processing -> results), separator and heading mode, page-sharded over G GPUs (one process per GPU, no collective).
Reports pages/s from disk per mode: the pipelined CLI (decode prefetch + same-size batching) and, for comparison, the
reference-shaped loop (one page per call, one decode thread).  usage: python tools/file_throughput.py [N] [G] [out.json]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
import numpy as np

import __graft_entry__ as entry

entry.build()
from aru_b200.synth import synth_page, synth_pb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
OUT = sys.argv[3] if len(sys.argv) > 3 else None
H, W = 1024, 768


def make_inputs(d):
    os.makedirs(os.path.join(d, "page"), exist_ok=True)
    base = [synth_page(H, W, s) for s in range(16)]
    paths = []
    for i in range(N):
        p = os.path.join(d, f"page_{i:05d}.png")
        cv2.imwrite(p, np.roll(base[i % 16], 3 * (i // 16), axis=1))
        rng = np.random.default_rng(i)
        lines = []
        for c in range(2):
            x0, y = 30 + c * (W // 2), 20
            while y + 30 < H - 20:
                lh, lw = int(rng.integers(14, 30)), int(rng.integers(W // 5, W // 2 - 40))
                lines.append(f'<TextLine id="l{len(lines)}"><Coords points="{x0},{y} {x0 + lw},{y} {x0 + lw},{y + lh} {x0},{y + lh}"/></TextLine>')
                y += lh + int(rng.integers(6, 20))
        with open(os.path.join(d, "page", f"page_{i:05d}.xml"), "w") as f:
            f.write('<?xml version="1.0"?><PcGts xmlns="http://schema.primaresearch.org/PAGE/gts/pagecontent/2013-07-15">'
                    f'<Page imageFilename="page_{i:05d}.png" imageWidth="{W}" imageHeight="{H}"><TextRegion id="r1">'
                    + "".join(lines) + "</TextRegion></Page></PcGts>")
        paths.append(p)
    lst = os.path.join(d, "pages.lst")
    with open(lst, "w") as f:
        f.write("\n".join(paths) + "\n")
    return lst


def run(lst, pb, mode, d, tag, extra):
    out = os.path.join(d, f"out_{mode}_{tag}")
    cmd = [sys.executable, "-m", "aru_b200.run_net_post_processing", "--path_to_image_list", lst, "--path_to_pb", pb,
           "--mode", mode, "--fixed_height", str(H), "--dump_dir", out, "--dump_format", "none", "--gpus", str(G)] + extra
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(r.stderr[-2000:])
    ranks = [json.loads(l) for l in r.stderr.splitlines() if l.startswith('{"rank"')]
    secs = max(x["seconds"] for x in ranks)
    return {"mode": mode, "variant": tag, "gpus": G, "pages": N, "pages_per_s": round(sum(x["pages"] for x in ranks) / secs, 1),
            "slowest_rank_s": secs, "wall_s_with_start_up": round(wall, 2), "decode_threads": ranks[0]["decode_threads"],
            "batch_pages": ranks[0]["batch_pages"]}


if __name__ == "__main__":
    d = tempfile.mkdtemp(prefix="aru_files_")
    try:
        t0 = time.perf_counter()
        lst = make_inputs(d)
        results = {"workload": f"{N} synthetic {H}x{W} PNG pages on local disk, {G} GPU(s), results: manifest only",
                   "cpus": len(os.sched_getaffinity(0)), "make_inputs_s": round(time.perf_counter() - t0, 1), "runs": []}
        for mode, net in (("separator", "separator"), ("heading", "heading")):
            pb = os.path.join(d, net + ".pb")
            with open(pb, "wb") as f:
                f.write(synth_pb(net))
            results["runs"].append(run(lst, pb, mode, d, "pipelined", []))
            results["runs"].append(run(lst, pb, mode, d, "one_page_per_call", ["--decode_threads", "1", "--batch_pages", "1"]))
            print(json.dumps(results["runs"][-2]), flush=True)
            print(json.dumps(results["runs"][-1]), flush=True)
        if OUT:
            with open(OUT, "w") as f:
                json.dump(results, f, indent=1)
    finally:
        shutil.rmtree(d, ignore_errors=True)
