"""On-disk formats: the feature-file batch loader (misc.py:257-302) and extract_binary's compound datasets
(binary_vectors.py:12-135).  CPU: mappings of arrays stand in for HDF5 files (h5py is optional and absent here); the exchanges
run on the emulated kernels.  tests/test_gpu_parity.py repeats the extraction on the CUDA library."""
import numpy as np
import pytest
import torch

from multimodalgame_b200 import formats as fmt, model as M
from tests import emu_util, golden_util as gu, surface_util as su


@pytest.fixture()
def emulated_library():
    M._LIB_OVERRIDE = emu_util.emu_library()
    M._BINDINGS.clear(); M._LAST_BINDING.clear()
    yield
    M._LIB_OVERRIDE = None
    M._BINDINGS.clear(); M._LAST_BINDING.clear()


def feature_store(n, feat_dim, n_classes, seed=0):
    """A mapping with the datasets of a reference feature file: Target, Location, avgpool_512 (n, F, 1, 1), fc, layer4_2."""
    rng = np.random.RandomState(seed)
    return {"Target": rng.randint(0, n_classes, size=n).astype(np.int64),
            "Location": np.array(["img_%04d.jpg" % i for i in range(n)], dtype="S50"),
            "avgpool_512": rng.randn(n, feat_dim, 1, 1).astype(np.float32),
            "fc": rng.randn(n, 10).astype(np.float32),
            "layer4_2": rng.randn(n, 4, 2, 2).astype(np.float32)}


def test_compound_dtypes_follow_the_reference():
    c = fmt.bin_vec_dtype(32)
    assert c.names == ("ExampleId", "AgentId", "Index", "Target", "Rank", "BinaryProb", "BinaryVec")      # binary_vectors.py:24-30
    assert c["ExampleId"] == np.dtype("<U50") and c["AgentId"] == np.dtype("<U1") and c["Index"] == np.dtype("i")
    assert c["BinaryProb"].shape == (32,) and c["BinaryVec"].base == np.float32
    p = fmt.preds_dtype(30)
    assert p.names == ("ExampleId", "AgentId", "Index", "Target", "Rank", "Predictions", "StopProb", "StopVec", "StopMask")
    assert p["Predictions"].shape == (30,) and p["StopProb"].shape == (1,) and p["StopMask"].base == np.float32


def test_load_hdf5_batches_like_the_reference():
    st = feature_store(23, 16, 5)
    plain = list(fmt.load_hdf5(st, 8, 0, shuffle=False))
    assert len(plain) == 2 and all(b["target"].dtype == torch.int64 and b["target"].shape == (8,) for b in plain)     # 23 // 8
    assert plain[0]["avgpool_512"].shape == (8, 16) and plain[0]["avgpool_512"].dtype == torch.float32               # .float().squeeze()
    assert np.array_equal(plain[1]["example_ids"], st["Location"][8:16])
    full = list(fmt.load_hdf5(st, 8, 0, shuffle=False, truncate_final_batch=True))
    assert [len(b["target"]) for b in full] == [8, 8, 7]
    # shuffled: seed 11 + random_seed, Python 2's shuffle loop, indices sorted inside each batch (misc.py:269-282)
    order = fmt.py2_shuffle(range(23), 11 + 3)
    assert sorted(order) == list(range(23)) and order != list(range(23))
    # regression pin of the Python-2 loop `j = int(random() * (i + 1))` on the seed-14 Mersenne-Twister stream
    assert order == [16, 7, 6, 0, 3, 14, 21, 11, 8, 22, 1, 20, 19, 9, 17, 10, 12, 4, 5, 18, 13, 15, 2]
    sh = list(fmt.load_hdf5(st, 8, 3, shuffle=True, truncate_final_batch=True, map_labels=lambda t: int(t) + 100))
    for i, b in enumerate(sh):
        idx = sorted(order[8 * i:8 * i + 8])
        assert np.array_equal(b["example_ids"], st["Location"][idx])
        assert b["target"].tolist() == [int(t) + 100 for t in st["Target"][idx]]
        assert torch.equal(b["fc"], torch.from_numpy(st["fc"][idx]))


def run_extract_case(device, tmp_path):
    """extract_binary over a small development set: rows, order, Rank and the values against the exchange it wraps."""
    z, cfg = gu.load("eval_adaptive")
    su.set_flags(cfg)
    params = gu.params_at(z, "P0")
    mods = su.build_modules(cfg, params, device)
    D, F, M_ = cfg.n_classes, cfg.img_feat_dim, cfg.sender_out_dim
    st = feature_store(10, F, D, seed=4)
    st["Target"][:] = 2                                     # "Rank only works if there is one target" (binary_vectors.py:98-100)
    desc = torch.from_numpy(z["desc"])
    M.FLAGS.binary_output = str(tmp_path / "binary.h5")
    M.FLAGS.img_feat = "avgpool_512"
    out = fmt.extract_binary(M.FLAGS, fmt.load_hdf5, M.exchange, st, 4, 0, False, device != "cpu", 2, mods["sender"],
                             mods["receiver"], dict(desc=desc), int, None)
    saved, close = fmt.open_store(M.FLAGS.binary_output)
    comm, preds = np.asarray(saved["Communication"]), np.asarray(saved["Predictions"])
    close()
    assert comm.dtype.names == fmt.bin_vec_dtype(M_).names and preds.dtype.names == fmt.preds_dtype(D).names
    # the same conversations, batch by batch, straight through exchange()
    rows_c = rows_p = 0
    for batch in fmt.load_hdf5(st, 4, 0, False, truncate_final_batch=True):
        data, target = batch["avgpool_512"].to(device), batch["target"].to(device)
        s, sen_w, rec_w, y, _, _ = M.exchange(mods["sender"], mods["receiver"], None, None,
                                              dict(data=data, target=target, desc=desc.to(device), train=False,
                                                   break_early=not cfg.fixed_exchange))
        bsz = len(target)
        for t in range(len(y)):
            sen = comm[rows_c:rows_c + bsz]; rec = comm[rows_c + bsz:rows_c + 2 * bsz]; pr = preds[rows_p:rows_p + bsz]
            rows_c += 2 * bsz; rows_p += bsz
            ids = [i.decode() if isinstance(i, bytes) else str(i) for i in batch["example_ids"]]
            as_str = lambda a: [v.decode() if isinstance(v, bytes) else str(v) for v in a]
            assert as_str(sen["ExampleId"]) == ids and set(as_str(sen["AgentId"])) == {"S"} and set(as_str(rec["AgentId"])) == {"R"}
            assert set(sen["Index"]) == {2 * t} and set(rec["Index"]) == {2 * t + 1} and set(pr["Index"]) == {2 * t + 1}
            assert np.array_equal(sen["BinaryVec"], sen_w[0][t].cpu().numpy()) and np.allclose(sen["BinaryProb"], sen_w[1][t].cpu().numpy())
            assert np.array_equal(rec["BinaryVec"], rec_w[0][t].cpu().numpy())
            yy = y[t].detach().cpu().numpy()
            assert np.allclose(pr["Predictions"], yy)
            assert np.array_equal(pr["Rank"], np.abs(yy.argsort(1) - D)[:, 2])
            assert np.array_equal(pr["StopVec"][:, 0], s[1][t].cpu().numpy().reshape(-1))
            assert np.array_equal(pr["StopMask"][:, 0], s[0][t].cpu().numpy().reshape(-1).astype(np.float32))
    assert rows_c == len(comm) and rows_p == len(preds) and len(comm) == 2 * len(preds)
    return out


def test_extract_binary_datasets(emulated_library, tmp_path):
    run_extract_case("cpu", tmp_path)
