"""Host logic of clair3_b200.launcher (the replacement of the reference's .npy + GNU-parallel GPU launcher,
clair3/CallVariantsFromCffiGPU.py:141-199): file-list cut, batching, reference-format shards, dropping of early-out
reference calls.  CPU tests use a stub model; the GPU test drives the real module end to end."""
import os

import numpy as np
import pytest

from clair3_b200 import launcher, synth
from oracle import decode_oracle as dec


def test_split_file_list_is_the_reference_cut():
    files = ["f%d" % i for i in range(10)]
    # reference: each = ceil(10 / 4) = 3 -> [0:3] [3:6] [6:9] [9:10]  (CallVariantsFromCffiGPU.py:141-149)
    assert launcher.split_file_list(files, 4) == [files[0:3], files[3:6], files[6:9], files[9:10]]
    assert launcher.split_file_list(files, 3) == [files[0:4], files[4:8], files[8:10]]
    assert launcher.split_file_list(files[:2], 4) == [files[0:1], files[1:2], [], []]
    assert sum(launcher.split_file_list(files, 8), []) == files
    with pytest.raises(ValueError):
        launcher.split_file_list(files, 0)


def _write_files(tmp_path, sizes, seed=0):
    r = np.random.default_rng(seed)
    prefixes, all_pos, all_alt = [], [], []
    for i, n in enumerate(sizes):
        prefix = str(tmp_path / ("pileup_chr20_%d" % i))
        x = synth.pileup_inputs(n, seed=seed + i, dtype=np.int8)
        np.save(prefix, x)                                           # .npy appended by numpy, like the reference (:447)
        with open(prefix + ".info", "w") as f:
            for j in range(n):
                seq = "".join(r.choice(list("ACGTN"), size=33, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
                pos = "chr20:%d:%s" % (1000 * i + j, seq)
                alt = "%d-X%s 3" % (int(r.integers(5, 90)), "ACGT"[j % 4])
                f.write("%s\t%s\n" % (pos, alt))
                all_pos.append(pos)
                all_alt.append(alt)
        prefixes.append(prefix)
    lst = str(tmp_path / "pileup_gpu_chunk_0")
    with open(lst, "w") as f:
        for p in prefixes:
            f.write(os.path.basename(p) + "\n")                      # relative to the list's directory (:110-113)
    return lst, prefixes, all_pos, all_alt


class _StubModel:
    """Module-protocol stand-in: deterministic 'probabilities' from the tensor, decode via the oracle restatement."""
    out_dim = 24

    def predict_stream(self, batches, streams=8):
        for x in batches:
            x = np.asarray(x).astype(np.float32)
            z = np.stack([x[:, 3 + (k % 30), k % 18] * 0.05 + 0.01 * k for k in range(24)], axis=1)
            rows = np.arange(len(x))
            z[rows, np.array([0, 4, 7, 9])[rows % 4]] += 8.0         # a confident homozygous gt21 call on every site ...
            z[rows[::2], 21] += 6.0                                  # ... and a confident 0/0 genotype on every other one
            y = np.concatenate([np.exp(z[:, :21]) / np.exp(z[:, :21]).sum(1, keepdims=True),
                                np.exp(z[:, 21:]) / np.exp(z[:, 21:]).sum(1, keepdims=True)], axis=1).astype(np.float32)
            yield y

    def decode_stage1(self, y, ref_gt21):
        return dec.decode_stage1(y, ref_gt21)


@pytest.mark.parametrize("drop", [False, True])
def test_rank_shards_in_reference_format_and_order(tmp_path, drop):
    sizes = [1500, 0 + 37, 1024, 2050, 5]
    lst, prefixes, all_pos, all_alt = _write_files(tmp_path, sizes)
    files = launcher.read_file_list(lst)
    assert files == prefixes
    model = _StubModel()
    world = 2
    got_pos, got_y = [], []
    expect_y = np.concatenate(list(model.predict_stream(np.load(p + ".npy") for p in prefixes)))
    for rank in range(world):
        shard = str(tmp_path / ("pred_%d" % rank))
        read, written = launcher.run_rank(model, launcher.split_file_list(files, world)[rank], shard, "pileup", drop_ref_calls=drop)
        pred = np.load(shard + ".prediction", mmap_mode="r")         # the reference's replay reader (CallVariants.py:1636-1638)
        pos = np.load(shard + ".position", mmap_mode="r")
        alt = np.load(shard + ".alt_info", mmap_mode="r")
        assert pred.dtype == np.float32 and pred.shape == (written, 24)
        assert pos.dtype == np.dtype("S100") and pos.shape == (written, 1) and alt.dtype == np.dtype("S2000")
        got_pos += [p[0].decode() for p in pos]
        got_y.append(np.array(pred))
        assert all(a[0].decode().split("-")[0].isdigit() for a in alt)
    got_y = np.concatenate(got_y)
    if not drop:
        assert got_pos == all_pos and np.array_equal(got_y, expect_y)
    else:
        gt = launcher.center_ref_gt21(all_pos)
        known = gt != 255
        early = dec.decode_stage1(expect_y, np.where(known, gt, 0).astype(np.uint8))["is_ref"].astype(bool) & known
        assert 0 < early.sum() < len(early)
        keep = np.nonzero(~early)[0]
        assert got_pos == [all_pos[i] for i in keep] and np.array_equal(got_y, expect_y[keep])


def test_center_ref_base():
    seq = "A" * 16 + "G" + "T" * 16
    assert launcher.center_ref_gt21(["chr1:100:" + seq, "chr1:5:C", "HLA:x:y:7:" + "N" * 33, "chr2:9:t"]).tolist() == [7, 4, 255, 9]


@pytest.mark.gpu
def test_launcher_single_rank_end_to_end(tmp_path, monkeypatch):
    import torch

    from clair3_b200.model import Clair3_P
    lst, prefixes, all_pos, all_alt = _write_files(tmp_path, [1300, 200, 1024])
    sd = synth.pileup_state_dict(False, seed=9)
    torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, str(tmp_path / "pileup.pt"))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    launcher.main(["--file_list", lst, "--chkpnt_fn", str(tmp_path / "pileup"), "--pileup", "--out_prefix", str(tmp_path / "pred"),
                   "--drop_ref_calls"])
    m = Clair3_P(False, True, 18)
    m.to(torch.device("cuda"))
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = np.concatenate([np.load(p + ".npy") for p in prefixes])
    y = m(torch.from_numpy(x)).numpy()
    gt = launcher.center_ref_gt21(all_pos)
    known = gt != 255
    early = dec.decode_stage1(y, np.where(known, gt, 0).astype(np.uint8))["is_ref"].astype(bool) & known
    keep = np.nonzero(~early)[0]
    pred = np.load(str(tmp_path / "pred_0.prediction"))
    pos = np.load(str(tmp_path / "pred_0.position"))
    assert [p[0].decode() for p in pos] == [all_pos[i] for i in keep]
    assert np.abs(pred - y[keep]).max() < 1e-5
    assert "--input_probabilities" in open(str(tmp_path / "pred_0.decode_cmd")).read()
