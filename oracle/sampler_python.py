"""ORACLE — test infrastructure only.  The SimANS negative sampler restated line for line from
/root/reference/SimANS/utils/MARCO_until_new.py:179-202 (tau form) and /root/reference/SimANS/utils/util_wiki.py:613-639
(a, b form): CPython `random.choices` rounds, `set` union, `list(set)[:k]`.  Returns candidate INDICES."""
import math
import random


def simans_sample_reference(neg_scores, pos_score, k, *, tau=None, a=None, b=0.0, rng=random):
    n = len(neg_scores)
    if pos_score == 0:
        return list(range(n))[-k:]
    neg_candidates, weights = [], []
    for i, s in enumerate(neg_scores):
        if tau is not None:
            weights.append(math.exp(-abs(s - pos_score) * tau))
        else:
            weights.append(math.exp(-((s - pos_score + b) ** 2) * a))
        neg_candidates.append(i)
    neg_ids_list = set()
    while len(neg_ids_list) < k:
        neg_ids_list = neg_ids_list.union(rng.choices(neg_candidates, weights=weights, k=k))
        new_c, new_w = [], []
        for c, w in zip(neg_candidates, weights):
            if c not in neg_ids_list:
                new_c.append(c)
                new_w.append(w)
        neg_candidates, weights = new_c, new_w
    return list(neg_ids_list)[0:k]


def simans_sample_mod128(neg_scores, pos_score, k, *, tau=None, a=None, b=0.0, rng=random):
    """Same rounds as `simans_sample_reference` with the truncation made explicit: when the rounds overshoot k, keep the
    k chosen candidates with the smallest (i & 127, i) — the iteration order of a 128-slot CPython set of small ints,
    which is what `list(set)[:k]` yields in ~95 % of the overshooting draws.  This is the rule the CUDA kernel implements."""
    n = len(neg_scores)
    if pos_score == 0:
        return list(range(n))[-k:]
    if tau is not None:
        weights = [math.exp(-abs(s - pos_score) * tau) for s in neg_scores]
    else:
        weights = [math.exp(-((s - pos_score + b) ** 2) * a) for s in neg_scores]
    cands = list(range(n))
    chosen = []
    while len(chosen) < k:
        for c in rng.choices(cands, weights=weights, k=k):
            if c not in chosen:
                chosen.append(c)
        keep = [(c, w) for c, w in zip(cands, weights) if c not in chosen]
        cands, weights = [c for c, _ in keep], [w for _, w in keep]
    if len(chosen) > k:
        chosen = sorted(chosen, key=lambda v: (v & 127, v))[:k]
    return chosen
