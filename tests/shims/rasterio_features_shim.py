"""Stand-in for ``rasterio.features.shapes`` (GDAL polygonize), test infrastructure only.

The reference extracts separator / region contours with ``rasterio.features.shapes(mask, connectivity=8)``
(region_net_post_processor_base.py:186-197: ``apply_contour_detection`` / ``apply_contour_detection2``) and keeps the
shapes whose value is 255.  rasterio (GDAL) is not installed in this image, so the polygon-level equivalence test drives the
reference's real ``to_polygons`` / ``rescale_polygons`` over this re-implementation:

  * one ``(geojson, value)`` pair per connected region of equal value (4- or 8-connectivity);
  * ``geojson = {"type": "Polygon", "coordinates": [exterior, hole, ...]}``, every ring a closed list of ``(x, y)`` float
    pixel-CORNER coordinates (identity transform), vertices only where the outline turns.

Ring start vertex and orientation are this module's own (GDAL's are not documented); both sides of an equivalence
comparison go through the same code, and comparisons of emitted XML should canonicalise rings anyway.
"""
from __future__ import annotations

import numpy as np

# directions: 0 = +x, 1 = +y (down), 2 = -x, 3 = -y; the region is kept on the right-hand side of every edge
_DX = (1, 0, -1, 0)
_DY = (0, 1, 0, -1)


def _rings_of_component(comp: np.ndarray, x0: int, y0: int, connectivity: int):
    """comp: bool array (bounding box of one component, padded by one background pixel all round)."""
    h, w = comp.shape
    ys, xs = np.nonzero(comp)
    edges = {}        # (vx, vy) -> list of outgoing directions
    def add(vx, vy, d):
        edges.setdefault((vx, vy), []).append(d)
    up = ~comp[ys - 1, xs]
    right = ~comp[ys, xs + 1]
    down = ~comp[ys + 1, xs]
    left = ~comp[ys, xs - 1]
    for y, x in zip(ys[up], xs[up]):
        add(x, y, 0)                  # top edge, walking +x, pixel below
    for y, x in zip(ys[right], xs[right]):
        add(x + 1, y, 1)              # right edge, walking +y
    for y, x in zip(ys[down], xs[down]):
        add(x + 1, y + 1, 2)          # bottom edge, walking -x
    for y, x in zip(ys[left], xs[left]):
        add(x, y + 1, 3)              # left edge, walking -y
    rings = []
    while edges:
        start = next(iter(edges))
        d = edges[start].pop()
        if not edges[start]:
            del edges[start]
        ring = [start]
        vx, vy = start[0] + _DX[d], start[1] + _DY[d]
        while (vx, vy) != start or False:
            outs = edges.get((vx, vy))
            if outs is None:
                break
            if len(outs) == 1:
                nd = outs[0]
            else:
                # two diagonal pixels of the region meet in this vertex: 8-connectivity walks on to the diagonal
                # neighbour (left turn, the outline stays one ring), 4-connectivity stays on its own pixel (right turn)
                want = (d - 1) % 4 if connectivity == 8 else (d + 1) % 4
                nd = want if want in outs else outs[0]
            outs.remove(nd)
            if not outs:
                del edges[(vx, vy)]
            if nd != d:
                ring.append((vx, vy))
            d = nd
            vx, vy = vx + _DX[d], vy + _DY[d]
        # the start vertex may lie in the middle of a straight run: drop it if the ring does not turn there
        first_d = (ring[1][0] - ring[0][0], ring[1][1] - ring[0][1]) if len(ring) > 1 else (0, 0)
        last_d = (ring[0][0] - ring[-1][0], ring[0][1] - ring[-1][1])
        if len(ring) > 2 and (first_d[0] * last_d[1] - first_d[1] * last_d[0]) == 0:
            ring = ring[1:]
        pts = [(float(px + x0 - 1), float(py + y0 - 1)) for px, py in ring]
        pts.append(pts[0])
        area2 = sum(a[0] * b[1] - b[0] * a[1] for a, b in zip(pts[:-1], pts[1:]))
        rings.append((area2, pts))
    # region on the right with y pointing down: the exterior ring has positive shoelace sum, holes negative
    exterior = max(rings, key=lambda r: r[0])
    holes = sorted((r for r in rings if r is not exterior), key=lambda r: (r[1][0][1], r[1][0][0]))
    return [exterior[1]] + [r[1] for r in holes]


def shapes(source, mask=None, connectivity=4, transform=None):
    """Generator of ``(geojson_polygon, value)`` like ``rasterio.features.shapes`` (identity transform only)."""
    import cv2
    img = np.asarray(source)
    if img.ndim != 2:
        raise ValueError("shapes() takes a 2-D array")
    if connectivity not in (4, 8):
        raise ValueError("connectivity must be 4 or 8")
    for value in np.unique(img):
        sel = (img == value)
        if mask is not None:
            sel &= np.asarray(mask).astype(bool)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(sel.astype(np.uint8), connectivity=connectivity)
        for i in range(1, n):
            x, y, w, h = (int(v) for v in stats[i, :4])
            comp = np.zeros((h + 2, w + 2), bool)
            comp[1:-1, 1:-1] = labels[y:y + h, x:x + w] == i
            rings = _rings_of_component(comp, x, y, connectivity)
            yield {"type": "Polygon", "coordinates": rings}, float(value)
