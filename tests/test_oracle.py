"""Pin the oracle (numpy restatement + torch timing port) against golden vectors minted from the
real reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN_FA, GOLDEN_PILEUP, golden_case
from oracle import clair3_oracle as orc
from oracle import torch_port


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("name", GOLDEN_PILEUP)
def test_numpy_oracle_pileup_matches_reference(name):
    z, meta, sd, x = golden_case(name)
    taps = {}
    y = orc.pileup_forward(sd, x, meta["add_indel_length"], taps=taps)
    assert y.shape == z["y"].shape
    assert np.abs(y - z["y"]).max() < 2e-5          # fp64 math vs fp32 reference
    n = z["tap_lstm1"].shape[0]
    assert _relerr(taps["lstm1"][:n], z["tap_lstm1"]) < 1e-5
    assert _relerr(taps["lstm2"][:n], z["tap_lstm2"]) < 1e-5
    assert _relerr(taps["l4_pre"], z["tap_l4_pre"]) < 1e-5


@pytest.mark.parametrize("name", GOLDEN_FA)
def test_numpy_oracle_fa_matches_reference(name):
    z, meta, sd, x = golden_case(name)
    taps = {}
    y = orc.fa_forward(sd, x, meta["add_indel_length"], taps=taps)
    assert y.shape == z["y"].shape
    assert np.abs(y - z["y"]).max() < 2e-5
    assert _relerr(taps["spp"], z["tap_spp"]) < 1e-5
    assert _relerr(taps["l4_pre"], z["tap_l4_pre"]) < 1e-5
    for k in ("conv1", "res_block1", "conv3", "res_block2", "conv5", "res_block3"):
        if "tap_" + k in z.files:
            assert _relerr(taps[k][:1], z["tap_" + k]) < 1e-5, k


@pytest.mark.parametrize("name", GOLDEN_PILEUP + GOLDEN_FA)
def test_torch_port_matches_reference(name):
    z, meta, sd, x = golden_case(name)
    if meta["kind"] == "pileup":
        y = torch_port.PileupPort(sd, meta["add_indel_length"])(x).numpy()
    else:
        y = torch_port.FullAlignmentPort(sd, meta["add_indel_length"])(x).numpy()
    assert np.abs(y - z["y"]).max() < 1e-5


def test_outputs_are_probabilities_and_peaked():
    z, meta, sd, x = golden_case("p24")
    y = z["y"]
    assert np.allclose(y[:, :21].sum(1), 1, atol=1e-5) and np.allclose(y[:, 21:].sum(1), 1, atol=1e-5)
    # trained-like: the synthetic heads are not uniform
    assert y[:, :21].max(1).mean() > 0.3


def test_depth_rescale_truncates_toward_zero():
    x = np.array([[[-7, 7, 300]]], dtype=np.int32)
    out = orc.depth_rescale_with(x, [288])          # scale 2.0 -> -3 (trunc), 3, 150
    assert out.tolist() == [[[-3, 3, 150]]]
    assert orc.depth_rescale_with(x, [200]).tolist() == x.tolist()   # <= 1.5*144: untouched


@pytest.mark.parametrize("out_dim", [24, 90])
def test_decode_oracle_matches_reference(out_dim):
    """oracle/decode_oracle.py vs the reference's own possible_outcome_probabilites_from / quality_score_from
    (fixture minted by tests/golden/make_decode_golden.py): flags and float32 products bit-exact, QUAL to double rounding."""
    import os

    from conftest import GOLDEN_DIR
    from oracle import decode_oracle as dec
    z = np.load(os.path.join(GOLDEN_DIR, "decode_stage1.npz"))
    y, ref_gt21 = z["y%d" % out_dim], z["ref_gt21_%d" % out_dim]
    assert np.array_equal(dec.ref_gt21_from_bases(str(z["bases%d" % out_dim])), ref_gt21)
    d = dec.decode_stage1(y, ref_gt21)
    assert np.array_equal(d["is_ref"], z["early%d" % out_dim])
    assert np.array_equal(d["ref_prob"], z["prob%d" % out_dim])                  # bit-exact float32
    assert np.allclose(d["qual"], z["qual%d" % out_dim], rtol=1e-13, atol=1e-13)
    assert np.array_equal(np.round(d["qual"], 2), z["qual_rounded%d" % out_dim])
    assert np.array_equal(d["nonref_idx"], np.nonzero(z["early%d" % out_dim] == 0)[0])
    assert 0 < d["n_nonref"][0] < len(y)


def test_pileup_windows_restatement():
    from oracle import decode_oracle as dec
    cols = np.arange(100 * 18, dtype=np.int64).reshape(100, 18) + 1
    w = dec.pileup_windows(cols, [0, 10, 67, -5, 90, 200])
    assert np.array_equal(w[1], cols[10:43]) and np.array_equal(w[2], cols[67:100])
    assert (w[3][:5] == 0).all() and np.array_equal(w[3][5:], cols[0:28])        # head overhang -> zero rows
    assert np.array_equal(w[4][:10], cols[90:100]) and (w[4][10:] == 0).all()    # tail overhang
    assert (w[5] == 0).all()
