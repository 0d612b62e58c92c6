"""ORACLE — test infrastructure only (never imported by the product path).

Numpy restatement of the data-parallel FIRST STAGE of the reference's per-site decoder, the part ``c3b_decode_stage1``
(clair3_b200/csrc/decode.cu) runs on the GPU.  Reference lines (paths relative to HKU-BAL/Clair3):

* head slicing              ``clair3/CallVariants.py:1072,1082``  (``param.label_shape_cum`` = 21, 24, 57, 90)
* early-out test + product  ``possible_outcome_probabilites_from``, ``clair3/CallVariants.py:519-534`` (no indel heads) and
                            ``:565-576`` (with ``add_indel_length``): ``homo_reference >= 0.5 and gt21[ref+ref] >= 0.5`` (and both
                            ``variant_length[0 + index_offset] >= 0.5``) -> ``[homo_Ref_probability]``; ``output_from`` then
                            reports a homozygous-reference call (``:690-695``) which ``output_with`` drops unless
                            ``is_show_reference`` (``:1182-1186``)
* ``homo_Ref_probability``  float32 products in the reference's evaluation order (``:527`` / ``:569-572``)
* QUAL                      ``quality_score_from`` before its ``round(.., 2)``, ``clair3/CallVariants.py:375-381``
* reference-base gt21 index ``gt21_enum_from_label(ref + ref)``, ``clair3/task/gt21.py:29-61`` (AA 0, CC 4, GG 7, TT 9)

Pinned by ``tests/golden/decode_stage1.npz``, minted by ``tests/golden/make_decode_golden.py`` from the reference's own
``possible_outcome_probabilites_from`` / ``quality_score_from`` (``tests/test_oracle.py::test_decode_oracle_matches_reference``).
"""
from __future__ import annotations

from math import e, log

import numpy as np

LABEL_CUM = (21, 24, 57, 90)           # shared/param_p.py label_shape_cum
HOMO_REFERENCE = 0                     # clair3/task/genotype.py:7
VL_OFFSET = 16                         # clair3/task/variant_length.py:6
REF_GT21 = {"A": 0, "C": 4, "G": 7, "T": 9}
PHRED_TRANS = -10 * log(e, 10)         # clair3/CallVariants.py:27


def ref_gt21_from_bases(bases):
    return np.array([REF_GT21[b] for b in bases], dtype=np.uint8)


def decode_stage1(y, ref_gt21):
    """y: float32 [B,24|90]; ref_gt21: uint8 [B].  Returns the dict ``Clair3_X.decode_stage1`` returns (numpy)."""
    y = np.asarray(y, dtype=np.float32)
    B, out_dim = y.shape
    nh = 4 if out_dim == 90 else 2
    bounds = (0,) + LABEL_CUM[:nh]
    argmax = np.zeros((B, nh), dtype=np.int32)
    maxprob = np.zeros((B, nh), dtype=np.float32)
    for h in range(nh):
        seg = y[:, bounds[h]:bounds[h + 1]]
        argmax[:, h] = seg.argmax(1)                       # first maximum
        maxprob[:, h] = seg.max(1)
    homo_ref = y[:, 21 + HOMO_REFERENCE]
    gt_ref = y[np.arange(B), ref_gt21.astype(np.int64)]
    early = (homo_ref >= 0.5) & (gt_ref >= 0.5)
    if nh == 4:
        v1, v2 = y[:, 24 + VL_OFFSET], y[:, 57 + VL_OFFSET]
        early &= (v1 >= 0.5) & (v2 >= 0.5)
        prob = ((v1 * v2) * homo_ref) * gt_ref             # float32, reference order  :567-572
    else:
        prob = homo_ref * gt_ref                            # :527
    prob = prob.astype(np.float32)
    ratio = ((np.float32(1.0) - prob) + np.float32(1e-10)) / (prob + np.float32(1e-10))    # float32 scalars (NumPy >= 2 promotion)
    qual = np.array([max(PHRED_TRANS * log(float(r)) + 10, 0) for r in ratio], dtype=np.float64)
    idx = np.nonzero(~early)[0].astype(np.int32)
    return {"is_ref": early.astype(np.uint8), "ref_prob": prob, "argmax": argmax, "maxprob": maxprob, "qual": qual,
            "nonref_idx": idx, "n_nonref": np.array([len(idx)], dtype=np.int32)}


def pileup_windows(cols, starts, positions=33):
    """The host-side window slicing ``c3b_forward_windows`` replaces (``preprocess/CreateTensorPileupFromCffi.py:362-394``):
    site b = rows [starts[b], starts[b]+33) of the per-column count matrix, zero rows where the window overhangs it."""
    cols = np.asarray(cols)
    out = np.zeros((len(starts), positions, cols.shape[1]), dtype=cols.dtype)
    for b, s in enumerate(np.asarray(starts, dtype=np.int64)):
        lo, hi = max(int(s), 0), min(int(s) + positions, cols.shape[0])
        if hi > lo:
            out[b, lo - int(s):hi - int(s)] = cols[lo:hi]
    return out
