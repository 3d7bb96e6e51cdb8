"""ORACLE — test infrastructure only (never imported by the product path).

numpy float64 restatement of the exact inner-product search the SimANS ANN refresh performs with faiss
(/root/reference/SimANS/co_training/co_training_generate.py:359-384, 412-427): faiss.IndexFlatIP.add(P) then
.search(Q, k) -> (D [nq, k] inner products in descending order, I [nq, k] row numbers).  faiss is a third-party
dependency absent from /root/reference and from this image (no version pinned: README "conda install faiss-gpu");
IndexFlatIP's published contract is brute-force maximum inner product, which is what is restated here.  Ties (which
faiss leaves unspecified) are broken towards the lower row number.  "parity unpinned" by executable faiss; pinned by
definition (argsort of the exact score matrix)."""
from __future__ import annotations

import numpy as np


def flat_ip_search(P, Q, k):
    P, Q = np.asarray(P, np.float64), np.asarray(Q, np.float64)
    S = Q @ P.T
    n = P.shape[0]
    order = np.argsort(-S, axis=1, kind="stable")[:, :k]  # stable: equal scores keep ascending row order
    D = np.take_along_axis(S, order, axis=1)
    I = order.astype(np.int64)
    if n < k:
        D = np.concatenate([D, np.full((Q.shape[0], k - n), -np.inf)], axis=1)
        I = np.concatenate([I, np.full((Q.shape[0], k - n), -1, np.int64)], axis=1)
    return D, I, S
