"""GPU-sharded counterpart of the reference's
``article_separation/image_segmentation/net_post_processing/run_net_post_processing.py``.

Same argparse surface (``--path_to_image_list --path_to_pb --mode {heading,separator} --num_processes
--fixed_height --scaling_factor --threshold``, defaults per reference :27-57: fixed height 1500 for the
separator net, 900 for the heading net).  What changes is the sharder (reference :59-82, a
``ProcessPoolExecutor`` of CPU workers with their own ``tf.Session``): here one *process per GPU* takes a
page shard (``sharding.shard_for_rank``), owns one engine and needs no collective.  Ranks come from
``torchrun`` (RANK / WORLD_SIZE / LOCAL_RANK) or from ``--gpus N``, which re-executes this module N times.

Two back ends per rank:
  * the reference's own, unmodified ``SeparatorNetPostProcessor`` / ``HeadingNetPostProcessor`` when the
    reference package (and its lxml / shapely / rasterio dependencies) is importable - ``net_boundary.install()``
    makes them call the B200 engine, and PAGE-XML is written exactly as before;
  * ``--dump_dir``: the path up to the polygon step - ``cv2.imread`` + ``scale_image`` on the host, then colour step ->
    net -> ``uint8(p*255)`` -> thresholded mask -> ``post_process`` (separator_net_post_processor.py:146-151, 25-99) on
    the device (``Engine.separator_pages``) - written as PNGs (probability, mask, horizontal, vertical) plus a
    per-rank manifest (a re-run skips pages already listed, and the parent checks every page was done: worker
    failures are silent in the reference because its futures are never awaited, reference :77,82).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from typing import List

import numpy as np

if __package__ in (None, ""):  # executed as a script: make ``aru_b200`` importable
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "aru_b200"

from . import net_boundary, sharding  # noqa: E402
from .engine import ARU_EUNSUP, EngineError  # noqa: E402


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="ARU-Net separator / heading inference on B200, page-sharded over GPUs")
    p.add_argument("--path_to_image_list", type=str, required=True, help="Path to the list file holding the image paths.")
    p.add_argument("--path_to_pb", type=str, required=True, help="Path to the TensorFlow pixel labelling graph.")
    p.add_argument("--num_processes", type=int, default=8,
                   help="Accepted for compatibility; parallelism is one process per GPU (--gpus / torchrun).")
    p.add_argument("--fixed_height", type=int, required=False, help="Input image height")
    p.add_argument("--scaling_factor", type=float, default=1.0, help="Scaling factor of images.")
    p.add_argument("--mode", type=str, required=True, choices=["heading", "separator"])
    p.add_argument("--threshold", type=float, default=0.05, help="Threshold for binarization of net output.")
    p.add_argument("--gpus", type=int, default=0, help="Spawn this many single-GPU ranks (0: use the torchrun env / 1 rank).")
    p.add_argument("--decode_threads", type=int, default=0,
                   help="--dump_dir: threads that decode images ahead of the GPU and encode the results behind it "
                        "(0 = the CPUs this rank may use, at most 16).")
    p.add_argument("--batch_pages", type=int, default=16,
                   help="--dump_dir: pages of one size that go to the GPU in one call.")
    p.add_argument("--dump_format", type=str, default="png", choices=["png", "none"],
                   help="--dump_dir: png = probability / mask images per page; none = manifest only (throughput runs).")
    p.add_argument("--device_cubic", action="store_true",
                   help="Enlarge small scans (scale > 1, cv2.INTER_CUBIC in the reference) on the device too: within one grey "
                        "level of OpenCV. Default: OpenCV on the host, as the reference does.")
    p.add_argument("--dump_dir", type=str, default=None,
                   help="Write uint8 probability maps / masks + a manifest here instead of running the PAGE-XML writers.")
    return p


def default_fixed_height(mode: str) -> int:
    return 900 if mode == "heading" else 1500  # reference :51-57


def _pixel_counts(paths: List[str]):
    try:
        from PIL import Image
    except Exception:
        return None
    out = []
    for p in paths:
        try:
            with Image.open(p) as im:
                out.append(im.size[0] * im.size[1])
        except Exception:
            return None
    return out


def _reference_available() -> bool:
    try:
        import importlib
        importlib.import_module("article_separation.image_segmentation.net_post_processing.region_to_page_writer")
        return True
    except Exception:
        return False


def run_rank(args, rank: int, world: int, local_rank: int) -> int:
    paths = net_boundary.load_image_paths(args.path_to_image_list)
    shard = sharding.shard_for_rank(paths, world, rank, _pixel_counts(paths))
    fixed_height = args.fixed_height if args.fixed_height is not None else default_fixed_height(args.mode)
    os.environ.setdefault("ARU_B200_DEVICE", str(local_rank))
    if args.dump_dir is None:
        if not _reference_available():
            raise SystemExit("the reference package (article_separation + lxml/shapely/rasterio) is not importable: "
                             "pass --dump_dir to write probability maps / masks instead of PAGE-XML")
        net_boundary.install()
        if args.mode == "separator":
            from article_separation.image_segmentation.net_post_processing.separator_net_post_processor import \
                SeparatorNetPostProcessor
            # install() ran before this import, when the class did not exist yet: bind the device post-processing now
            if not net_boundary.patch_separator_post_processor():
                raise SystemExit("could not bind SeparatorNetPostProcessor.post_process to the B200 engine")
            SeparatorNetPostProcessor(shard, args.path_to_pb, fixed_height, args.scaling_factor, args.threshold,
                                      gpu_devices="").run()
        else:
            from article_separation.image_segmentation.net_post_processing.heading_net_post_processor import \
                HeadingNetPostProcessor
            net_boundary.patch_separator_post_processor()   # RegionNetPostProcessor.apply_cc_analysis on the device
            HeadingNetPostProcessor(shard, args.path_to_pb, fixed_height, args.scaling_factor,
                                    {"net": 0.8, "stroke_width": 0.0, "text_height": 0.2}, 0.4,
                                    {"net_thresh": 1.0, "stroke_width_thresh": 1.0, "text_height_thresh": 0.9,
                                     "sw_th_thresh": 0.9}, 0.8).run(gpu_device="")
        return len(shard)

    return _run_rank_dump(args, shard, rank, world, fixed_height)


def _run_pages(eng, images, args, fixed_height):
    """A batch of decoded pages of ONE size (uint8 [n,H,W,3] as ``cv2.imread`` returns them) through
    ``load_and_scale_image``'s remaining steps, the net and the integer post-processing.  Shrinking (INTER_AREA, the usual
    case: scans are larger than the net input), the colour step, the net, uint8 / threshold and ``post_process`` run on the
    device; enlarging (INTER_CUBIC) is done by cv2 on the host as in the reference unless ``--device_cubic``.
    Returns (result dict of [n,...] arrays, scale, (H, W) of the net input)."""
    sc = net_boundary._scaling_factor(images.shape[1], images.shape[2], args.scaling_factor, fixed_height=fixed_height)
    want_sep = args.mode == "separator"
    kw = dict(threshold=args.threshold, want_u8=True, want_mask=True, u8_channels=1)
    if sc < 1.0 or (sc > 1.0 and args.device_cubic):
        h, w = eng.scaled_size(images.shape[1], images.shape[2], sc)
        try:
            return eng.separator_images(images, sc, want_separators=want_sep and h >= 50 and w >= 100, **kw), sc, (h, w)
        except EngineError as err:
            if err.code != ARU_EUNSUP:
                raise
    scaled = np.stack([net_boundary.scale_image(im, fixed_height, args.scaling_factor)[0] for im in images])
    h, w = scaled.shape[1:3]
    return eng.separator_pages(scaled, want_separators=want_sep and h >= 50 and w >= 100, **kw), sc, (h, w)


def _read_page_textlines(image_path):
    from . import page_textlines as T
    page_xml = T.page_path_for_image(image_path)
    return T.read_textlines(page_xml) if os.path.exists(page_xml) else []


def _run_heading_pages(eng, images, paths, args, fixed_height, textlines=None):
    """Heading mode for a batch of decoded pages of one size: the net's uint8 map and, for the pages that have a
    ``<dir>/page/<stem>.xml``, the network feature of every TextLine - ``get_net_prob_for_text_line``
    (heading_net_post_processor.py:247-270), box sums on the device.
    Returns ([(uint8 map [H,W,C], {text line id: probability} or None)], scale, (H, W))."""
    from . import page_textlines as T
    sc = net_boundary._scaling_factor(images.shape[1], images.shape[2], args.scaling_factor, fixed_height=fixed_height)
    per_page, boxes = [], []
    for i, path in enumerate(paths):
        lines = textlines[i] if textlines is not None else _read_page_textlines(path)
        boxed = [(lid, T.textline_box(pts, sc)) for lid, pts in lines if pts]
        per_page.append((lines, boxed, len(boxes)))
        boxes += [(i, y, y + h, x, x + w) for _, (x, y, w, h) in boxed]
    res = None
    if sc < 1.0 or (sc > 1.0 and args.device_cubic):
        try:
            res = eng.heading_images(images, sc, boxes, want_u8=True)
        except EngineError as err:
            if err.code != ARU_EUNSUP:
                raise
    if res is None:
        scaled = np.stack([net_boundary.scale_image(im, fixed_height, args.scaling_factor)[0] for im in images])
        res = eng.heading_pages(scaled, boxes, want_u8=True)
    sums, _, u8 = res
    out = []
    for i, (lines, boxed, b0) in enumerate(per_page):
        probs = None
        if lines:
            probs = {lid: 0 for lid, pts in lines if not pts}          # no surrounding polygon: 0 (head:259-260)
            for k, (lid, (x, y, w, h)) in enumerate(boxed):
                probs[lid] = T.net_prob(int(sums[b0 + k]), w, h)
        out.append((u8[i], probs))
    return out, sc, u8.shape[1:3]


def _run_rank_dump(args, shard, rank, world, fixed_height) -> int:
    """--dump_dir: decode -> GPU -> encode as a pipeline.  The reference reads, runs and writes one page at a time per worker
    process (run_net_post_processing.py:59-82, separator_net_post_processor.py:141); here a thread pool decodes ahead of
    the GPU (``cv2.imread`` releases the GIL), decoded pages of one size are batched into one device call (the engine
    micro-batches and overlaps copies with compute inside it), and the same pool encodes / writes the results behind."""
    import collections
    import itertools
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(args.dump_dir, exist_ok=True)
    manifest_path = os.path.join(args.dump_dir, f"manifest_rank{rank}.json")
    done = {}
    if os.path.exists(manifest_path):
        with open(manifest_path) as f:
            done = json.load(f).get("pages", {})
    todo = [p for p in shard if p not in done]
    graph = net_boundary.load_graph(args.path_to_pb)
    eng = graph.engine(net_boundary.resolve_device(""))          # binds this rank's threads to the GPU's NUMA node
    n_threads = args.decode_threads or max(1, min(16, len(os.sched_getaffinity(0))))
    batch = max(1, args.batch_pages)
    def write_png(name, arr):
        if args.dump_format == "png":
            cv2.imwrite(os.path.join(args.dump_dir, name), np.ascontiguousarray(arr))

    def save_manifest():
        with open(manifest_path, "w") as f:
            json.dump({"rank": rank, "world": world, "pages": done}, f)

    def flush(group, pool):
        """One device call for a group of same-size pages; returns (manifest entries, futures of their file writes)."""
        paths = [g[0] for g in group]
        images = np.stack([g[1] for g in group])
        entries, futs = {}, []
        if args.mode == "heading":
            pages, sc, shape = _run_heading_pages(eng, images, paths, args, fixed_height, [g[2] for g in group])
            for path, (u8, probs) in zip(paths, pages):
                stem = os.path.splitext(os.path.basename(path))[0]
                futs.append(pool.submit(write_png, stem + "_prob.png", u8[:, :, 0].copy()))
                if probs is not None:
                    with open(os.path.join(args.dump_dir, stem + "_textlines.json"), "w") as f:
                        json.dump(probs, f)
                entries[path] = {"scale": sc, "shape": list(shape)}
        else:
            r, sc, shape = _run_pages(eng, images, args, fixed_height)
            for i, path in enumerate(paths):
                stem = os.path.splitext(os.path.basename(path))[0]
                futs.append(pool.submit(write_png, stem + "_prob.png", r["u8"][i, :, :, 0].copy()))
                futs.append(pool.submit(write_png, stem + "_mask.png", r["mask"][i].copy()))
                if "horizontal" in r:
                    futs.append(pool.submit(write_png, stem + "_horizontal.png", r["horizontal"][i].copy()))
                    futs.append(pool.submit(write_png, stem + "_vertical.png", r["vertical"][i].copy()))
                entries[path] = {"scale": sc, "shape": list(shape)}
        return entries, futs

    in_flight = collections.deque()          # (manifest entries, write futures) of flushed groups, oldest first

    def retire(keep):
        """A page enters the manifest once its files are on disk (a re-run skips exactly the pages listed there)."""
        while len(in_flight) > keep:
            entries, futs = in_flight.popleft()
            for f in futs:
                f.result()
            done.update(entries)
            save_manifest()

    def decode(path):
        """``cv2.imread(path)`` as the reference calls it (helper.py:29) - except that a file which holds ONE 8-bit channel
        is kept as one channel: its B = G = R copy would go through the same resize per channel and BGR2GRAY
        ((3735 + 19235 + 9798) g + 2^14) >> 15 = g, i.e. identical results for a third of the bytes."""
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None or img.dtype != np.uint8 or img.ndim != 2:
            img = cv2.imread(path)
        # heading mode: the page's TextLines are read here too, off the thread that feeds the GPU
        return img, (_read_page_textlines(path) if args.mode == "heading" and img is not None else None)

    import time
    t_start = time.perf_counter()
    # three stages: the pool decodes ahead, this thread groups decoded pages by size, one worker thread owns the GPU
    # (stacking a batch, the device call and handing the results to the pool for encoding)
    with ThreadPoolExecutor(n_threads) as pool, ThreadPoolExecutor(1) as gpu:
        it = iter(todo)
        window = collections.deque((p, pool.submit(decode, p)) for p in itertools.islice(it, 2 * n_threads + batch))
        pending = collections.OrderedDict()                      # decoded pages waiting for a full batch, per image size
        flushing = collections.deque()
        while window:
            path, fut = window.popleft()
            nxt = next(it, None)
            if nxt is not None:
                window.append((nxt, pool.submit(decode, nxt)))
            image, lines = fut.result()
            if image is None:
                raise SystemExit(f"cannot read image {path}")
            group = pending.setdefault(image.shape, [])
            group.append((path, image, lines))
            if len(group) >= batch:
                flushing.append(gpu.submit(flush, pending.pop(image.shape), pool))
            elif sum(len(g) for g in pending.values()) > 4 * batch:   # many sizes in flight: run the oldest group
                flushing.append(gpu.submit(flush, pending.pop(next(iter(pending))), pool))
            while len(flushing) > 2:                                  # at most two batches queued behind the GPU
                in_flight.append(flushing.popleft().result())
            retire(keep=2)
        for shape in list(pending):
            flushing.append(gpu.submit(flush, pending.pop(shape), pool))
        while flushing:
            in_flight.append(flushing.popleft().result())
        retire(keep=0)
    save_manifest()
    dt = time.perf_counter() - t_start
    # one line per rank: pages from disk to results on disk, engine start-up excluded (tools/file_throughput.py adds them up)
    print(json.dumps({"rank": rank, "world": world, "mode": args.mode, "pages": len(todo), "seconds": round(dt, 4),
                      "pages_per_s": round(len(todo) / dt, 1) if dt > 0 else None, "decode_threads": n_threads,
                      "batch_pages": batch, "dump_format": args.dump_format}), file=sys.stderr, flush=True)
    return len(shard)


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    rank, world, local = sharding.rank_from_env()
    if args.gpus and args.gpus > 1 and "RANK" not in os.environ:
        procs = []
        for r in range(args.gpus):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(args.gpus), LOCAL_RANK=str(r))
            procs.append(subprocess.Popen([sys.executable, "-m", "aru_b200.run_net_post_processing"] + (argv or sys.argv[1:]),
                                          env=env))
        rcs = [p.wait() for p in procs]
        if any(rcs):
            raise SystemExit(f"rank exit codes {rcs}")
        if args.dump_dir:   # every page must be in exactly one manifest
            paths = net_boundary.load_image_paths(args.path_to_image_list)
            seen = []
            for r in range(args.gpus):
                with open(os.path.join(args.dump_dir, f"manifest_rank{r}.json")) as f:
                    seen += list(json.load(f)["pages"])
            if sorted(seen) != sorted(paths):
                raise SystemExit("page manifest does not cover the image list exactly once")
        return 0
    run_rank(args, rank, world, local)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
