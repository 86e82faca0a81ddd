"""Text-line boxes for the heading net's per-line feature, read from PAGE-XML with the standard library.

What ``HeadingNetPostProcessor.get_net_prob_for_text_line`` (heading_net_post_processor.py:247-270) does on the host
before it sums the network output, restated without the reference's lxml-based page parser:

  * a TextLine's surrounding polygon is its ``Coords/@points`` ("x1,y1 x2,y2 ...", PAGE 2013-07-15) - lines without
    coordinates score 0 (head:259-260);
  * ``Polygon.rescale(sc)`` truncates: ``(int(x * sc), int(y * sc))`` (python_util/geometry/point.py:11);
  * the bounding box is inclusive: ``Rectangle(min_x, min_y, max_x - min_x + 1, max_y - min_y + 1)``
    (python_util/geometry/polygon.py:82-92);
  * the feature is ``sum(net_output[y:y+height, x:x+width]) / (width * height)`` with ``net_output = u8[..., 0] / 255``
    - the sum runs on the device (``Engine.heading_pages``), the division here.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import List, Optional, Tuple


def page_path_for_image(image_path: str, page_folder_name: str = "page") -> str:
    """python_util/io/file_loader.py:23-36 (append_extension=False)."""
    d, name = os.path.dirname(image_path), os.path.basename(image_path)
    return os.path.join(d, page_folder_name, os.path.splitext(name)[0] + ".xml")


def _local(tag: str) -> str:
    return tag.rsplit("}", 1)[-1]


def read_textlines(page_xml_path: str) -> List[Tuple[str, Optional[List[Tuple[int, int]]]]]:
    """[(TextLine id, polygon points or None)] in document order."""
    out = []
    for el in ET.parse(page_xml_path).getroot().iter():
        if _local(el.tag) != "TextLine":
            continue
        pts = None
        for ch in el:
            if _local(ch.tag) == "Coords" and ch.get("points"):
                pts = [tuple(int(float(v)) for v in p.split(",")) for p in ch.get("points").split()]
                break
        out.append((el.get("id", ""), pts))
    return out


def textline_box(points, sc: float):
    """(x, y, width, height) of the rescaled polygon, as ``Polygon.rescale`` + ``get_bounding_box`` give it."""
    xs = [int(x * sc) for x, _ in points]
    ys = [int(y * sc) for _, y in points]
    return min(xs), min(ys), max(xs) - min(xs) + 1, max(ys) - min(ys) + 1


def net_prob(box_sum: int, width: int, height: int) -> float:
    return int(box_sum) / 255 / (width * height)
