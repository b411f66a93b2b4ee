"""ORACLE (test infrastructure): numpy restatement of the reference's host tail after the packed top-k kernel.
REF mistralrs-core/src/sampler.rs:1172-1273 (probabilities, nucleus cut, min-p, weighted draw), :532-534 (cutoff on the
mass renormalised over the kept set); the draw restates rand's WeightedIndex: the count of running weights (all but the
last) that are <= the chosen weight.  All arithmetic in float32, in the reference's order."""
import numpy as np

f32 = np.float32


def pack_row(logits, k, inv_t):
    """What the device kernel hands over: top-k raw values (value desc, index asc), their indices, denom, max."""
    x = np.asarray(logits, dtype=np.float32)
    order = np.lexsort((np.arange(x.size), -x.astype(np.float64)))[:k]
    s = x.astype(np.float64) * inv_t
    gm = s.max()
    return np.concatenate([x[order], order.astype(np.float32), [np.exp(s - gm).sum(), gm]]).astype(np.float32)


def sample_row(packed, packed_k, row_k, inv_t, top_p, min_p, u):
    row_k = min(row_k, packed_k)
    vals, idx = packed[:row_k], packed[packed_k:packed_k + row_k]
    denom, gmax = packed[2 * packed_k], packed[2 * packed_k + 1]
    report = (np.exp(vals * f32(inv_t) - gmax, dtype=np.float32) / denom).astype(np.float32)
    if row_k == 1:
        return int(idx[0]), float(np.log(report[0])), report, report
    w = report.copy()
    if 0.0 < top_p < 1.0:
        kept = f32(0)
        for p in w:
            kept = f32(kept + p)
        cutoff, run = f32(f32(top_p) * kept), f32(0)
        for i in range(row_k):
            if run >= cutoff:
                w[i] = 0
            else:
                run = f32(run + w[i])
    if 0.0 < min_p < 1.0:
        thr = f32(w[0] * f32(min_p))
        w[thr >= w] = 0
    total = f32(0)
    for p in w:
        total = f32(total + p)
    if not total > 0:
        raise ValueError("all zero")
    chosen = f32(u * float(total))
    run, cum = f32(0), []
    for p in w[:-1]:
        run = f32(run + p)
        cum.append(run)
    sel = int(sum(c <= chosen for c in cum))
    while sel > 0 and w[sel] == 0:
        sel -= 1
    return int(idx[sel]), float(np.log(report[sel])), report, w
