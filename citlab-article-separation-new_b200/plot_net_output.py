"""Drop-in for the reference's ``article_separation/plot_net_output.py`` (the fourth consumer of the ``inImg`` -> ``output``
contract): apply a frozen ARU-Net graph to the images of a ``.lst`` file, print the class
statistics / accuracy and save (or show) the per-class maps, optionally blended over the page.

Same function names, arguments and printed lines as the reference (file:line cited per function); the ``tf.Session`` of
``plot_net_output.py:152-196`` is replaced by the B200 engine behind ``net_boundary.load_graph`` / ``get_net_output``, and
the per-pixel Python loop that builds the arg-max one-hot map and the class counts (``:216-225``, ~1 s per megapixel) by
its vectorised equivalent.  Three places where the reference cannot run as written are handled instead of reproduced and
are marked DEVIATION below."""
from __future__ import annotations

import colorsys
import os
import random
from argparse import ArgumentParser

import numpy as np

from . import net_boundary


def load_graph(frozen_graph_filename):
    """plot_net_output.py:16-37 -> the engine's graph handle (no TensorFlow)."""
    return net_boundary.load_graph(frozen_graph_filename)


def random_colors(N, bright=True):
    """plot_net_output.py:40-53."""
    brightness = 1.0 if bright else 0.7
    hsv = [(i / N, 1, brightness) for i in range(N)]
    colors = list(map(lambda c: colorsys.hsv_to_rgb(*c), hsv))
    random.shuffle(colors)
    return colors


def apply_mask(image, mask, color, alpha=0.5):
    """plot_net_output.py:56-69: blend ``color`` into ``image`` where ``mask == 255`` (in place, the image's dtype)."""
    for c in range(3):
        image[:, :, c] = np.where(mask == 255, image[:, :, c] * (1 - alpha) + alpha * color[c], image[:, :, c])
    return image


def plot_image_with_net_output(image, net_output):
    """plot_net_output.py:71-92 (the reference also draws ten random colours and prints them; only the fixed red is used)."""
    return apply_mask(image, net_output, (255, 50, 50), 0.5)


def plot_connected_components(image):
    """plot_net_output.py:95-105."""
    import cv2
    _, image_bin = cv2.threshold(image, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    _, labels, _, _ = cv2.connectedComponentsWithStats(255 - image_bin)
    return labels


def compute_accuracy(hyp_image, gt_image):
    """plot_net_output.py:108-116: share of equal pixels."""
    return np.sum(hyp_image == gt_image) / gt_image.size


def argmax_one_hot(out_img):
    """plot_net_output.py:212-225 without the per-pixel loop: the one-hot map of the arg-max class (first maximum wins, as
    ``np.argmax``) in the dtype of ``out_img`` and the pixel count per class."""
    n_class = out_img.shape[-1]
    winners = np.argmax(out_img, axis=3)
    one_hot = np.zeros_like(out_img)
    np.put_along_axis(one_hot, winners[..., None], 1, axis=3)
    counts = np.bincount(winners.reshape(-1), minlength=n_class)
    return one_hot, {"class_" + str(i): int(counts[i]) for i in range(n_class)}


def _scaling_factor(img_height, rescale, fixed_height):
    """plot_net_output.py:168-174."""
    if fixed_height:
        return (rescale if rescale and rescale != 1 else 1) * fixed_height / img_height
    return rescale if rescale else None


def _read_page(path, rescale, fixed_height):
    """plot_net_output.py:164-180: the BGR page, its gray version, the scaling factor and the unscaled size; both are area-resampled when
    0.1 < factor < 1 (DEVIATION: with neither ``rescale`` nor ``fixed_height`` the reference compares None with a float and
    raises TypeError; here the page is taken as it is)."""
    import cv2
    bgr = cv2.imread(path)
    if bgr is None:
        raise FileNotFoundError(path)
    full_shape = bgr.shape[:2]
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    sc = _scaling_factor(bgr.shape[0], rescale, fixed_height)
    if sc is not None and 0.1 < sc < 1.0:
        bgr, gray = (cv2.resize(a, None, fx=sc, fy=sc, interpolation=cv2.INTER_AREA) for a in (bgr, gray))
    return bgr, gray, sc, full_shape


def _read_ground_truth(path_to_img, n_class, sc):
    """plot_net_output.py:202-210: ``<dir>/C<n>/<name>_GT<i>.png`` per class, gray, resized (bilinear) by the factor."""
    import cv2
    folder, name = os.path.dirname(path_to_img), os.path.splitext(os.path.basename(path_to_img))[0]
    gts = []
    for i in range(n_class):
        gt = cv2.cvtColor(cv2.imread(os.path.join(folder, f"C{n_class}", f"{name}_GT{i}.png")), cv2.COLOR_BGR2GRAY)
        gts.append(cv2.resize(gt, None, fx=sc, fy=sc) if sc else gt)
    return gts


def _render(plane_u8, page_rgb, gt, with_img, with_gt):
    """plot_net_output.py:243-250: the class map, blended over the page and / or next to its ground truth."""
    out = plot_image_with_net_output(page_rgb.astype(np.uint32), plane_u8) if with_img else plane_u8
    if with_gt:
        out = np.concatenate((out, plot_image_with_net_output(page_rgb, gt) if with_img else gt), axis=1)
    return out


def plot_net_output(path_to_pb, path_to_img_lst, save_folder="", gpu_device="0", rescale=None, fixed_height=None,
                    mask_threshold=None, plot_with_gt=False, plot_with_img=False, show_plot=False,
                    calculate_accuracy=True, figsize=(16, 16), ax=None):
    """plot_net_output.py:131-283, same arguments and printed lines.  Returns the per-image accuracies (the reference
    prints them and returns None)."""
    import cv2
    graph = load_graph(path_to_pb)
    with open(path_to_img_lst) as f:
        paths = [line.rstrip() for line in f if line.strip()]
    accuracies = []
    for path_to_img in paths:
        page, gray, sc, (full_h, full_w) = _read_page(path_to_img, rescale, fixed_height)
        prob = net_boundary.get_net_output(gray / 255.0, graph, gpu_device)[None]      # [1, H, W, C], as sess.run returns it
        n_class = prob.shape[-1]
        print("Percentage of Pixels where the net is not 100% sure: ", np.sum((0 < prob) & (prob < 1)) / prob.size)
        if mask_threshold:
            prob = np.array((prob > 0.6), np.int32)                                     # the reference's fixed 0.6
        gts = _read_ground_truth(path_to_img, n_class, sc) if (plot_with_gt or calculate_accuracy) else None
        one_hot, counts = argmax_one_hot(prob)
        for class_name, count in counts.items():                                        # relative to the UNSCALED page, :226
            print(f"Percentage of pixels in {class_name}: {count / (full_w * full_h)}")
        name, ext = os.path.splitext(os.path.basename(path_to_img))
        accuracy = 0
        for cl in range(n_class):
            if calculate_accuracy:
                accuracy += compute_accuracy(one_hot[0, :, :, cl], gts[cl] / 255)
            if plot_with_img:
                page = cv2.cvtColor(page, cv2.COLOR_BGR2RGB)      # once per class, as the reference does (:244)
            shown = _render(np.uint8(prob[0, :, :, cl] * 255), page, gts[cl] if plot_with_gt else None, plot_with_img, plot_with_gt)
            if save_folder:
                out_u8 = np.uint8(shown)
                # DEVIATION: the reference converts RGB -> BGR unconditionally, which OpenCV rejects for the single-channel
                # map of plot_with_img=False; that map is written as it is
                cv2.imwrite(os.path.join(save_folder, f"{name}_OUT{cl}{ext}"),
                            cv2.cvtColor(out_u8, cv2.COLOR_RGB2BGR) if out_u8.ndim == 3 else out_u8)
            if show_plot:
                import matplotlib.pyplot as plt                   # DEVIATION: needed only when a window is asked for
                if plot_with_img:
                    _, ax = plt.subplots(1, figsize=figsize)
                    ax.set_ylim(page.shape[0] + 10, -10)
                    ax.set_xlim(-10, page.shape[1] + 10)
                    ax.axis('off')
                    ax.imshow(shown.astype(np.uint8))
                else:
                    plt.imshow(shown, cmap="gray")
                plt.show()
        accuracies.append(accuracy / n_class)
        print("Accuracy = ", accuracies[-1])
        print("+++++++++++++++++++++++")
    if accuracies:
        print("Overall Accuracy = ", sum(accuracies) / len(accuracies))
    return accuracies


def main(argv=None):
    ap = ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument('--path_to_tf_graph', required=True)
    ap.add_argument('--path_to_img_lst', required=True)
    ap.add_argument('--save_folder', default="")
    ap.add_argument('--rescale_factor', default=1.0, type=float)
    ap.add_argument('--fixed_height', default=0, type=int)
    ap.add_argument('--calculate_accuracy', action='store_true')
    ap.add_argument('--plot_with_img', action='store_true')
    ap.add_argument('--plot_with_gt', action='store_true')
    ap.add_argument('--gpu_device', default="0")
    a = ap.parse_args(argv)
    plot_net_output(a.path_to_tf_graph, a.path_to_img_lst, a.save_folder, gpu_device=a.gpu_device, rescale=a.rescale_factor,
                    fixed_height=a.fixed_height, plot_with_img=a.plot_with_img, plot_with_gt=a.plot_with_gt,
                    calculate_accuracy=a.calculate_accuracy)


if __name__ == '__main__':
    main()
