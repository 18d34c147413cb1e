#!/usr/bin/env python
"""CPU emulation (numpy) of the flat index on config 2 with two point orders: the 60-bit Morton
order and the aligned KD order (recursive splits at fixed, 8*2^k-aligned positions along the widest
dimension).  Prints boxes tested per level, leaves and points evaluated per query.  Not part of the
product or the tests; the numbers are quoted in DESIGN.md section 4."""
import sys, os, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigmap_b200 import host as H
from oracle.oracle import Port

REF_BP = int(os.environ.get("REF_BP", 4_600_000))
model = H.load_pore_model()
ref = H.sim_reference(20251017, [REF_BP])
pos, val = H.build_point_cloud(ref, model[0])
N = len(val); W = N - 5
win = np.lib.stride_tricks.sliding_window_view(val, 6)[:W]
vmin, vmax = val.min(), val.max(); inv = 1.0 / (vmax - vmin)


def morton_order():
    q = np.clip(((win - vmin) * inv * 1024).astype(np.int64), 0, 1023)
    def spread(v):
        r = np.zeros_like(v)
        for i in range(10): r |= ((v >> i) & 1) << (6 * i)
        return r
    code = np.zeros(W, np.int64)
    for d in range(6): code |= spread(q[:, d]) << (5 - d)
    return np.argsort(code, kind='stable')


def kd_order(rule="widest"):
    idx = np.arange(W, dtype=np.int64)
    s = 8
    while s < W: s *= 2
    level = 0
    while s > 8:
        seg = np.arange(W) // s
        starts = np.arange(0, W, s)
        pts = win[idx]
        if rule == "widest":
            ext = np.maximum.reduceat(pts, starts, axis=0) - np.minimum.reduceat(pts, starts, axis=0)
            dim = ext.argmax(1)
        else:
            dim = np.full(len(starts), level % 6)
        key = pts[np.arange(W), dim[seg]]
        # segments with <= s/2 points need no split, but sorting them is harmless
        o = np.lexsort((key, seg))
        idx = idx[o]
        s //= 2
        level += 1
    return idx


def rd(x):
    h = x.astype(np.float16); h = np.where(h.astype(np.float32) > x, np.nextafter(h, np.float16(-np.inf)), h); return h.astype(np.float32)
def ru(x):
    h = x.astype(np.float16); h = np.where(h.astype(np.float32) < x, np.nextafter(h, np.float16(np.inf)), h); return h.astype(np.float32)


def build(order):
    pts = win[order]
    nl = (W + 7) // 8
    pad = np.full((nl * 8 - W, 6), 1e18, np.float32)
    P = np.concatenate([pts, pad]).reshape(nl, 8, 6)
    real = np.concatenate([np.ones(W, bool), np.zeros(nl * 8 - W, bool)]).reshape(nl, 8)
    lo = np.where(real[:, :, None], P, np.inf).min(1); hi = np.where(real[:, :, None], P, -np.inf).max(1)
    levels = [(rd(lo), ru(hi))]
    while len(levels[-1][0]) > 8:
        l, h = levels[-1]; n = len(l); m = (n + 7) // 8
        lp = np.concatenate([l, np.full((m * 8 - n, 6), np.inf, np.float32)]).reshape(m, 8, 6).min(1)
        hp = np.concatenate([h, np.full((m * 8 - n, 6), -np.inf, np.float32)]).reshape(m, 8, 6).max(1)
        levels.append((lp, hp))
    return P, levels


reads = H.sim_reads(20251018, ref, 40, model=model)
port = Port(); Q = []
for r in range(40):
    pa = port.raw_to_pa(reads.read(r), 8192.0, 10.0, 1437.976685)
    f = port.generate_events(pa[:4000])
    for p in range(2, len(f) - 5, 2): Q.append(f[p:p + 6])
Q = np.stack(Q)[::7]
r2 = np.float32(0.08)


def run(name, order):
    P, levels = build(order)
    stats = []
    for qv in Q:
        front = np.arange(len(levels[-1][0]))
        counts = []
        for k in range(len(levels) - 1, -1, -1):
            l, h = levels[k]
            t = np.maximum(np.maximum(l[front] - qv, qv - h[front]), 0)
            ok = (t * t).sum(1) <= r2 * 1.005
            surv = front[ok]
            counts.append((len(front), len(surv)))
            if k > 0:
                front = (surv[:, None] * 8 + np.arange(8)).ravel()
                front = front[front < len(levels[k - 1][0])]
            else:
                leaves = surv
        d2 = ((P[leaves] - qv) ** 2).sum(2)
        hits = (d2 < r2).sum()
        stats.append(([c[0] for c in counts], len(leaves), hits))
    tested = np.array([s[0] for s in stats]); lv = np.array([s[1] for s in stats]); ht = np.array([s[2] for s in stats])
    print(f'== {name}: levels', [len(l[0]) for l in levels])
    print('  queries', len(Q), 'boxes tested per level (top..leaf boxes):', tested.mean(0).round(1), 'sum', tested.sum(1).mean().round(1))
    print('  leaves visited mean', lv.mean().round(1), 'p50', np.median(lv), 'p90', np.percentile(lv, 90), 'max', lv.max(), ' hits mean', ht.mean().round(1))
    steps = np.ceil(tested / 64).sum(1) + np.ceil(lv / 8)
    print('  steps/query (64 boxes or 8 leaves per step): mean', steps.mean().round(2), ' max frontier', tested.max())


t0 = time.time(); run('morton', morton_order()); print('  t', round(time.time() - t0, 1))
for rule in os.environ.get("RULES", "widest,cycle").split(","):
    t0 = time.time(); run('kd-' + rule, kd_order(rule)); print('  t', round(time.time() - t0, 1))
