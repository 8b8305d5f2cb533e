"""Oracle: the reference's thresholded best-path decoder and the Keras/TF CTC decoders.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Parity unpinned by reference fixtures.

Follows:
  * /root/reference/multimodal_fusion/sequence_decoding.py:21-69 ``decode_batch`` (threshold 0.5,
    22-class map :26-29, ignore list :32, MLF format :35-36,:60-65);
    /root/reference/audio_network/sequence_decoding.py:19-69 (threshold 0.75, 44-word map,
    ``_audio`` suffix :61).  The reference is Python 2: ``zip`` returns a *list snapshot*, so the
    loop at :45-48 iterates the original pairs while ``list.remove`` deletes the FIRST equal
    element -- restated literally with ``list(zip(...))``.
  * Keras 2.1.4 ``K.ctc_decode`` -> TF 1.12.1 ``ctc_greedy_decoder`` / ``CTCBeamSearchDecoder``
    (not called by the reference; named by BASELINE.json config 5).  SURVEY.md A.6.
"""
import itertools
import math

import numpy as np

MAP_GEST_FUSION = {0: "oov", 1: "VA", 2: "VQ", 3: "PF", 4: "FU", 5: "CP", 6: "CV",
                   7: "DC", 8: "SP", 9: "CN", 10: "FN", 11: "OK", 12: "CF", 13: "BS",
                   14: "PR", 15: "NU", 16: "FM", 17: "TT", 18: "BN", 19: "MC",
                   20: "ST", 21: "sil"}
IGNORE_LIST = [228, 298, 299, 300, 303, 304, 334, 343, 373, 375]


def decode_ids_literal(sample, threshold, drop_frames=2):
    """sequence_decoding.py:41-50 for one sample, literally (Python-2 zip snapshot).
    sample: (T, C) float32 probabilities.  Returns list of int class ids (blank kept)."""
    out_prob = list(np.max(sample[drop_frames:], 1))
    out_best = list(np.argmax(sample[drop_frames:], 1))
    for p, s in list(zip(out_prob, out_best)):
        if p < threshold:
            out_prob.remove(p)
            out_best.remove(s)
    return [int(k) for k, _ in itertools.groupby(out_best)]


def decode_ids_closed_form(sample, threshold, drop_frames=2):
    """Closed form of the above: for each class s delete the FIRST n_low[s]
    occurrences of s, n_low[s] = #{t: best[t]==s and conf[t] < threshold}; then collapse."""
    y = np.asarray(sample)[drop_frames:]
    if y.shape[0] == 0:
        return []
    best = np.argmax(y, 1)
    conf = np.max(y, 1)
    low = conf.astype(np.float64) < float(threshold)
    C = y.shape[1]
    n_low = np.bincount(best[low], minlength=C)
    seen = np.zeros(C, dtype=np.int64)
    kept = []
    for t in range(best.shape[0]):
        s = best[t]
        if seen[s] >= n_low[s]:
            kept.append(int(s))
        seen[s] += 1
    return [k for k, _ in itertools.groupby(kept)]


def decode_batch(pred_out, f_list, threshold=0.5, map_gest=None, mlf_path=None,
                 ignore_list=IGNORE_LIST, name_fmt="Sample%s"):
    """decode_batch restated (returns list[list[str]]; writes the MLF if mlf_path is given)."""
    map_gest = MAP_GEST_FUSION if map_gest is None else map_gest
    ret = []
    lines = ["#!MLF!#\n"]
    for j in range(pred_out.shape[0]):
        ids = decode_ids_literal(pred_out[j], threshold)
        outstr = [map_gest[i] for i in ids]
        ret.append(outstr)
        f_num = f_list[j]
        if int(f_num) in ignore_list:
            continue
        lines.append('"*/%s.rec"\n' % (name_fmt % format(f_num, "05")))
        for cl in outstr:
            lines.append("%s\n" % cl)
        lines.append(".\n")
    if mlf_path is not None:
        with open(mlf_path, "w") as of:
            of.writelines(lines)
    return ret


# ------------------------------------------------------------------ Keras/TF decoders
def ctc_greedy(probs, seq_len, eps=1e-8):
    """K.ctc_decode(greedy=True): inputs = log(p+eps); per frame argmax (first max wins);
    emit c_t iff c_t != blank and c_t != c_{t-1}; score = -sum_t max_c inputs.  probs (T, C)."""
    x = np.log(np.asarray(probs, dtype=np.float32)[:seq_len] + np.float32(eps))
    blank = x.shape[1] - 1
    out = []
    prev = -1
    score = np.float32(0)
    for t in range(x.shape[0]):
        c = int(np.argmax(x[t]))
        score = np.float32(score - x[t, c])
        if c != blank and c != prev:
            out.append(c)
        prev = c
    return out, float(score)


NEG = float("-inf")


def _lse2(a, b):
    if a == NEG:
        return b
    if b == NEG:
        return a
    m = max(a, b)
    return m + math.log1p(math.exp(min(a, b) - m))


def ctc_beam_search(probs, seq_len, beam_width=100, eps=1e-8, merge_repeated=True,
                    top_paths=1):
    """TF CTCBeamSearchDecoder restated in its batch ("parallel") form.

    inputs = log(p+eps), per step lp = log_softmax(inputs).  State per beam entry:
    p_blank, p_label, p_total (log).  One step: (1) every current leaf is updated
    (label term extended from its parent only if the parent is still a leaf); (2) every
    (leaf, non-blank label) pair whose child is not already a leaf is a candidate with
    p_total = lp[c] + (c == leaf.label ? leaf.old_blank : leaf.old_total); (3) the new leaf set is
    the top `beam_width` of (updated leaves ++ candidates) by p_total, ties broken by position
    in that list (leaves in descending previous p_total, then candidates by (leaf rank, label)).
    This equals TF's sequential TopN loop whenever p_totals are distinct.
    Returns list of (labels, log_prob) for the top_paths best leaves."""
    x = np.log(np.asarray(probs, dtype=np.float64)[:seq_len] + eps)
    C = x.shape[1]
    blank = C - 1
    # node pool: parent, label
    parent = [-1]
    label = [-1]
    # leaves: dict node -> [old(b,l,t), new(b,l,t)]
    leaves = [(0, (0.0, NEG, 0.0))]  # (node id, newp=(blank,label,total)), sorted by total desc
    for t in range(x.shape[0]):
        row = x[t]
        m = row.max()
        lp = row - (m + math.log(np.exp(row - m).sum()))
        active = {n: pr for n, pr in leaves}
        updated = []
        child_of = set()
        for n, old in leaves:
            ob, ol, ot = old
            nl = NEG
            if parent[n] >= 0:
                nl = ol
                pn = parent[n]
                if pn in active:
                    pb, pl_, pt = active[pn]
                    prev = pb if label[n] == label[pn] else pt
                    nl = _lse2(nl, prev)
                nl = nl + lp[label[n]] if nl != NEG else NEG
                child_of.add((pn, label[n]))
            nb = ot + lp[blank]
            updated.append((n, (nb, nl, _lse2(nb, nl)), False, None))
        cands = list(updated)
        for n, old in leaves:
            ob, ol, ot = old
            if ot == NEG:
                continue
            for c in range(C):
                if c == blank or (n, c) in child_of:
                    continue
                prev = ob if c == label[n] else ot
                if prev == NEG:
                    continue
                tot = lp[c] + prev
                cands.append((None, (NEG, tot, tot), True, (n, c)))
        order = sorted(range(len(cands)), key=lambda i: (-cands[i][1][2], i))[:beam_width]
        new_leaves = []
        for i in order:
            n, pr, is_new, pc = cands[i]
            if pr[2] == NEG:
                continue
            if is_new:
                parent.append(pc[0])
                label.append(pc[1])
                n = len(parent) - 1
            new_leaves.append((n, pr))
        leaves = new_leaves
    res = []
    for n, pr in leaves[:top_paths]:
        seq = []
        k = n
        while k != 0:  # node 0 is the root (empty prefix)
            seq.append(label[k])
            k = parent[k]
        seq = seq[::-1]
        if merge_repeated:
            seq = [k for k, _ in itertools.groupby(seq)]
        res.append((seq, pr[2]))
    return res


def brute_force_best_labelling(probs, eps=1e-8):
    """Most probable LABELLING (sum over alignments) by exhaustive enumeration; tiny cases only."""
    x = np.log(np.asarray(probs, dtype=np.float64) + eps)
    x = x - np.log(np.exp(x).sum(axis=1, keepdims=True))
    y = np.exp(x)
    T, C = y.shape
    blank = C - 1
    tot = {}
    for path in itertools.product(range(C), repeat=T):
        out = []
        prev = None
        for c in path:
            if c != prev and c != blank:
                out.append(c)
            prev = c
        pr = 1.0
        for t, c in enumerate(path):
            pr *= y[t, c]
        key = tuple(out)
        tot[key] = tot.get(key, 0.0) + pr
    best = max(tot.items(), key=lambda kv: kv[1])
    return list(best[0]), math.log(best[1]), tot
