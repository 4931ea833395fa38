"""Seeded synthetic inputs shared by the golden generator and the tests (numpy Generator streams are stable)."""
import numpy as np


def frame(seed: int, h: int = 720, w: int = 1280, rects: int = 20) -> np.ndarray:
    """uint8 BGR frame: noise + flat rectangles (SURVEY 8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for _ in range(rects):
        x0, y0 = int(rng.integers(0, w - 40)), int(rng.integers(0, h - 40))
        bw, bh = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        img[y0:y0 + bh, x0:x0 + bw] = rng.integers(0, 256, 3, dtype=np.uint8)
    return img


def yolo_v8_head(seed: int, n_hot: int = 120, nc: int = 80, A: int = 8400) -> np.ndarray:
    """[4+nc, A] float32: boxes U(50,590)/U(20,200), background probs U(0,0.05), n_hot anchors with one prob U(0.4,0.95)."""
    rng = np.random.default_rng(seed)
    out = np.empty((4 + nc, A), np.float32)
    out[0:2] = rng.uniform(50, 590, (2, A))
    out[2:4] = rng.uniform(20, 200, (2, A))
    out[4:] = rng.uniform(0, 0.05, (nc, A))
    hot = rng.choice(A, n_hot, replace=False)
    # clustered boxes so the NMS has real overlaps (and its swap path triggers)
    centers = rng.uniform(100, 540, (max(n_hot // 6, 1), 2))
    for j, a in enumerate(hot):
        c = centers[j % len(centers)]
        out[0, a] = c[0] + rng.normal(0, 12)
        out[1, a] = c[1] + rng.normal(0, 12)
        out[2, a] = 120 + rng.normal(0, 15)
        out[3, a] = 90 + rng.normal(0, 15)
        out[4 + rng.integers(0, nc), a] = rng.uniform(0.4, 0.95)
    return out


def yolo_v5_head(seed: int, n_hot: int = 100, nc: int = 80, A: int = 25200) -> np.ndarray:
    """[A, 5+nc] float32 decoded v5 rows (cx,cy,w,h,obj,cls...)."""
    rng = np.random.default_rng(seed)
    out = np.empty((A, 5 + nc), np.float32)
    out[:, 0:2] = rng.uniform(50, 590, (A, 2))
    out[:, 2:4] = rng.uniform(20, 200, (A, 2))
    out[:, 4] = rng.uniform(0, 0.3, A)
    out[:, 5:] = rng.uniform(0, 0.2, (A, nc))
    hot = rng.choice(A, n_hot, replace=False)
    centers = rng.uniform(100, 540, (max(n_hot // 5, 1), 2))
    for j, a in enumerate(hot):
        c = centers[j % len(centers)]
        out[a, 0:2] = c + rng.normal(0, 10, 2)
        out[a, 2:4] = (110, 80) + rng.normal(0, 12, 2)
        out[a, 4] = rng.uniform(0.7, 0.99)
        out[a, 5 + rng.integers(0, nc)] = rng.uniform(0.6, 0.99)
    return out


def yolo_v5_lite_head(seed: int, n_hot: int = 100, nc: int = 80, in_hw=(640, 640)) -> np.ndarray:
    """[A, 5+nc] float32 SIGMOID-ONLY v5 rows (all values in (0,1)) as a YOLOv5-lite export emits them: the grid / anchor
    decode is left to YoloLiteParameters.lite_postprocess.  Hot rows decode to ~100 px boxes near shared centres."""
    rng = np.random.default_rng(seed)
    A = sum(3 * (in_hw[0] // s) * (in_hw[1] // s) for s in (8, 16, 32))
    out = np.empty((A, 5 + nc), np.float32)
    out[:, 0:4] = rng.uniform(0.05, 0.95, (A, 4))
    out[:, 4] = rng.uniform(0, 0.3, A)
    out[:, 5:] = rng.uniform(0, 0.2, (A, nc))
    hot = rng.choice(A, n_hot, replace=False)
    for a in hot:
        out[a, 2:4] = rng.uniform(0.55, 0.95, 2)
        out[a, 4] = rng.uniform(0.7, 0.99)
        out[a, 5 + rng.integers(0, nc)] = rng.uniform(0.6, 0.99)
    # clusters: neighbouring grid cells of one level fire together (overlapping boxes for the NMS)
    base = 3 * 80 * 80 + 40 * 40            # level 1 (stride 16), anchor 1
    for k, cell in enumerate(rng.choice(40 * 38, 12, replace=False)):
        for d in (0, 1, 40):
            a = base + cell + d
            out[a, 0:2] = rng.uniform(0.3, 0.7, 2)
            out[a, 2:4] = rng.uniform(0.6, 0.8, 2)
            out[a, 4] = rng.uniform(0.75, 0.99)
            out[a, 5 + (k % nc)] = rng.uniform(0.7, 0.99)
    return out


def ufld_heads(seed: int, ngr=200, ncr=72, ngc=100, ncc=81, nl=4, invalid_lanes=()):
    """4 head tensors [1,...] float32: loc ~ N(0,3) with a smooth ridge, exist logits mostly valid."""
    rng = np.random.default_rng(seed)
    loc_row = rng.normal(0, 3, (1, ngr, ncr, nl)).astype(np.float32)
    loc_col = rng.normal(0, 3, (1, ngc, ncc, nl)).astype(np.float32)
    for i in range(nl):
        ridge = np.clip((np.linspace(0.2, 0.8, ncr) + 0.1 * i) * ngr, 0, ngr - 1).astype(int)
        ridge[:2] = (0, ngr - 1)                                      # argmax at the grid edges
        loc_row[0, ridge, np.arange(ncr), i] += 12
        ridge = np.clip((np.linspace(0.7, 0.3, ncc) + 0.05 * i) * ngc, 0, ngc - 1).astype(int)
        ridge[:2] = (ngc - 1, 0)
        loc_col[0, ridge, np.arange(ncc), i] += 12
    ex_row = rng.normal(0, 1, (1, 2, ncr, nl)).astype(np.float32)
    ex_col = rng.normal(0, 1, (1, 2, ncc, nl)).astype(np.float32)
    ex_row[0, 1] += 1.5
    ex_col[0, 1] += 1.5
    for i in invalid_lanes:
        ex_row[0, 1, :, i] -= 6
        ex_col[0, 1, :, i] -= 6
    ex_row[0, 0, 5, :] = ex_row[0, 1, 5, :]                          # exact tie -> argmax 0 (invalid)
    return [loc_row, loc_col, ex_row, ex_col]


def ufld_v1_head(seed: int, griding: int = 100, rows: int = 56, invalid_lanes=()):
    """UFLD v1 output [1, griding+1, rows, 4] float32: a smooth ridge per lane, the last ("no lane") bin winning on some rows, an
    exact tie between the ridge and the no-lane bin (argmax -> first = the ridge), lanes in `invalid_lanes` mostly empty."""
    rng = np.random.default_rng(seed)
    out = rng.normal(0, 2, (1, griding + 1, rows, 4)).astype(np.float32)
    for l in range(4):
        ridge = np.clip((np.linspace(0.15, 0.85, rows) + 0.07 * l) * griding, 0, griding - 1).astype(int)
        ridge[:2] = (0, griding - 1)
        out[0, ridge, np.arange(rows), l] += 10
        none = rng.random(rows) < (0.95 if l in invalid_lanes else 0.15)
        out[0, griding, none, l] += 25
        out[0, griding, 7, l] = out[0, :griding, 7, l].max()          # tie with the ridge: argmax keeps the lower index
    return out


def nms_case(seed: int, n: int):
    """xywh float32 boxes + float64 confs with heavy overlap."""
    rng = np.random.default_rng(seed)
    k = max(n // 5, 1)
    centers = rng.uniform(100, 1000, (k, 2))
    b = np.empty((n, 4), np.float32)
    for j in range(n):
        c = centers[j % k]
        wh = rng.uniform(60, 180, 2)
        b[j] = (c[0] + rng.normal(0, 15) - wh[0] / 2, c[1] + rng.normal(0, 15) - wh[1] / 2, wh[0], wh[1])
    confs = rng.uniform(0.4, 0.99, n).astype(np.float32).astype(np.float64)
    return b, confs


def track_sequence(seed: int, frames: int = 40, objects: int = 8):
    """Per-frame (boxes_xyxy int list, scores list, labels list): linear motion, drop-outs, low scores, clutter, births."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(100, 1000, (objects, 2))
    vel = rng.uniform(-12, 12, (objects, 2))
    size = rng.uniform(40, 160, (objects, 2))
    label = [f"class{int(c)}" for c in rng.integers(0, 3, objects)]
    born = rng.integers(0, frames // 3, objects)
    seq = []
    for f in range(frames):
        boxes, scores, labels = [], [], []
        for o in range(objects):
            if f < born[o]:
                continue
            p = pos[o] + vel[o] * (f - born[o])
            if rng.random() < 0.12:
                continue                                           # missed detection
            s = rng.uniform(0.55, 0.95) if rng.random() > 0.2 else rng.uniform(0.15, 0.45)
            jit = rng.normal(0, 2.0, 4)
            x1, y1 = p[0] - size[o, 0] / 2 + jit[0], p[1] - size[o, 1] / 2 + jit[1]
            x2, y2 = p[0] + size[o, 0] / 2 + jit[2], p[1] + size[o, 1] / 2 + jit[3]
            boxes.append([int(x1), int(y1), int(x2), int(y2)])
            scores.append(float(np.float32(s)))
            labels.append(label[o] if rng.random() > 0.05 else "class9")
        for _ in range(int(rng.integers(0, 3))):                   # clutter
            x, y = rng.uniform(0, 1100, 2)
            boxes.append([int(x), int(y), int(x + rng.uniform(30, 90)), int(y + rng.uniform(30, 90))])
            scores.append(float(np.float32(rng.uniform(0.12, 0.8))))
            labels.append("class5")
        seq.append((boxes, scores, labels))
    return seq


def ego_lanes(seed: int, img_w: int = 1280, img_h: int = 720):
    """Two plausible ego-lane polylines in the frontal view: lists of (int x, int y) converging towards the horizon, with a mild
    seeded curvature (inputs of the bird-view geometry tests)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(18, 40))
    ys = np.linspace(img_h * rng.uniform(0.55, 0.7), img_h - 1, n)
    t = (ys - ys[0]) / (ys[-1] - ys[0])
    bend = rng.uniform(-120, 120)
    centre = img_w * rng.uniform(0.42, 0.58) + bend * (1 - t) ** 2
    half = img_w * (0.04 + rng.uniform(0.16, 0.24) * t)
    jit = rng.normal(0, 1.5, (2, n))
    left = [(int(x), int(y)) for x, y in zip(centre - half + jit[0], ys)]
    right = [(int(x), int(y)) for x, y in zip(centre + half + jit[1], ys)]
    return left, right
