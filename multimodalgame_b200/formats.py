"""On-disk formats either side of the path: the feature files `misc.load_hdf5` reads (misc.py:257-302)
and the compound datasets `extract_binary` writes (binary_vectors.py:12-135), with the reference's dataset names, field
names, dtypes and row order.

h5py is optional.  A *store* is anything that maps dataset names to array-likes: an open `h5py.File`, a dict of NumPy
arrays, or the `.npz` container `NpzStore` writes when h5py is not installed (same dataset names and compound dtypes, so a
file converts one-to-one).  Nothing here is on the GPU path: the exchanges themselves run through `model.exchange`.
"""
import os
import random

import numpy as np
import torch

FEATURE_KEYS = ("layer4_2", "avgpool_512", "fc")          # misc.py:291-296


def bin_vec_dtype(sender_out_dim):
    """Rows of the "Communication" dataset (binary_vectors.py:24-30)."""
    return np.dtype([("ExampleId", np.str_, 50), ("AgentId", np.str_, 1), ("Index", "i"), ("Target", "i"), ("Rank", "i"),
                     ("BinaryProb", np.float32, (sender_out_dim,)), ("BinaryVec", np.float32, (sender_out_dim,))])


def preds_dtype(n_classes):
    """Rows of the "Predictions" dataset (binary_vectors.py:35-44)."""
    return np.dtype([("ExampleId", np.str_, 50), ("AgentId", np.str_, 1), ("Index", "i"), ("Target", "i"), ("Rank", "i"),
                     ("Predictions", np.float32, (n_classes,)), ("StopProb", np.float32, (1,)), ("StopVec", np.float32, (1,)),
                     ("StopMask", np.float32, (1,))])


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


class NpzStore(object):
    """Minimal write store used when h5py is absent: datasets are collected in memory and written as one `.npz` on close
    (compound rows as NumPy structured arrays).  `open_store(path)` reads it back as a plain dict."""

    def __init__(self, path):
        self.path, self.data = path, {}

    def append(self, name, rows):
        self.data[name] = rows if name not in self.data else np.concatenate([self.data[name], rows])

    def close(self):
        np.savez(self.path, **self.data)
        if not self.path.endswith(".npz") and os.path.exists(self.path + ".npz"):
            os.replace(self.path + ".npz", self.path)          # keep the caller's file name (np.savez appends .npz)


class H5Store(object):
    """h5py-backed write store: resizable compound datasets exactly as binary_vectors.py:31-46 creates them."""

    def __init__(self, path):
        import h5py
        self.f, self.sets = h5py.File(path, "w"), {}

    def append(self, name, rows):
        if name not in self.sets:
            self.sets[name] = self.f.create_dataset(name, (0,), maxshape=(None,), dtype=_h5_dtype(rows.dtype))
        ds = self.sets[name]
        ds.resize(ds.shape[0] + len(rows), axis=0)
        ds[-len(rows):] = rows.astype(_h5_dtype(rows.dtype))

    def close(self):
        self.f.close()


def _h5_dtype(dt):
    """h5py has no fixed-width unicode type: np.str_ fields are stored as fixed-width bytes of the same length (what h5py made
    of the reference's Python 2 `np.str_`, which WAS a byte string)."""
    return np.dtype([(n, ("S%d" % (dt[n].itemsize // 4)) if dt[n].kind == "U" else dt[n]) for n in dt.names])


def open_write_store(path):
    return H5Store(path) if _have_h5py() else NpzStore(path)


def open_store(source):
    """A readable store from a path (HDF5 when h5py is installed, else the `.npz` container) or from a mapping (returned as is).
    Returns (store, close_fn)."""
    if not isinstance(source, (str, bytes, os.PathLike)):
        return source, (lambda: None)
    path = os.path.expanduser(source)
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic.startswith(b"\x89HDF"):
        if not _have_h5py():
            raise ImportError("%s is an HDF5 file and h5py is not installed; convert it or pass a mapping of arrays" % path)
        import h5py
        f = h5py.File(path, "r")
        return f, f.close
    z = np.load(path, allow_pickle=False)
    return z, z.close


def py2_shuffle(order, seed):
    """`random.seed(seed); random.shuffle(order)` as Python 2 executed it (misc.py:269-271): the Mersenne-Twister stream is the
    same in Python 3 for an int seed, but `shuffle` itself changed (it no longer uses `int(random() * (i + 1))`), so the
    reference's batch composition is only reproduced by the old loop."""
    rng = random.Random(seed)
    order = list(order)
    for i in reversed(range(1, len(order))):
        j = int(rng.random() * (i + 1))
        order[i], order[j] = order[j], order[i]
    return order


def load_hdf5(hdf5_file, batch_size, random_seed, shuffle, truncate_final_batch=False, map_labels=int):
    """misc.load_hdf5 (misc.py:257-302): yields batch dicts {target (LongTensor), example_ids, layer4_2, avgpool_512, fc} in
    the reference's order (shuffled with seed 11 + random_seed, indices sorted inside a batch, optional short final batch).
    `hdf5_file`: a path or any mapping with the datasets "Target", "Location" and the feature sets that exist."""
    store, close = open_store(hdf5_file)
    try:
        dataset_size = int(store["Target"].shape[0])
        order = list(range(dataset_size))
        if shuffle:
            order = py2_shuffle(order, 11 + random_seed)
        num_batches = dataset_size // batch_size
        if truncate_final_batch and dataset_size - num_batches * batch_size > 0:
            num_batches += 1
        names = set(store.keys()) if hasattr(store, "keys") else set(FEATURE_KEYS)
        for i in range(num_batches):
            idx = sorted(order[i * batch_size:(i + 1) * batch_size])
            batch = dict()
            batch["target"] = torch.LongTensor([map_labels(t) for t in np.asarray(store["Target"])[idx]])
            batch["example_ids"] = np.asarray(store["Location"])[idx]
            for key in FEATURE_KEYS:
                if key in names:
                    batch[key] = torch.from_numpy(np.asarray(store[key])[idx]).float().squeeze()
            yield batch
    finally:
        close()


def communication_records(exchange_out, example_ids, target, n_classes):
    """The rows `extract_binary` appends for ONE batch (binary_vectors.py:85-135), in its order: per exchange step the sender's
    message rows ("S", Index 2t), then the receiver's ("R", Index 2t + 1), and the receiver's prediction rows.
    `exchange_out`: what `exchange()` returns.  Returns (communication_rows, prediction_rows) as structured arrays."""
    s, sen_w, rec_w, y, _, _ = exchange_out
    s_masks, s_feats, s_probs = s
    sen_feats, sen_probs = sen_w
    rec_feats, rec_probs = rec_w
    tgt = target.detach().cpu().numpy()
    bsz = tgt.shape[0]
    assert len(set(tgt.tolist())) == 1, "Rank only works if there is one target"      # binary_vectors.py:98-100
    single = int(tgt[0])
    M = int(sen_feats[0].shape[1])
    cdt, pdt = bin_vec_dtype(M), preds_dtype(n_classes)
    comm, preds = [], []
    ids = np.asarray(example_ids).astype(np.str_)
    cpu = lambda t: t.detach().cpu().numpy()
    for t, (zb, zp, wb, wp, yy, sf, sp, sm) in enumerate(zip(sen_feats, sen_probs, rec_feats, rec_probs, y, s_feats, s_probs,
                                                             s_masks)):
        np_preds = cpu(yy)
        rank = np.abs(np_preds.argsort(1) - n_classes)[:, single]                     # binary_vectors.py:101 (as written)
        for agent, index, probs, vec in (("S", 2 * t, zp, zb), ("R", 2 * t + 1, wp, wb)):
            rows = np.zeros(bsz, dtype=cdt)
            rows["ExampleId"], rows["AgentId"], rows["Index"], rows["Target"], rows["Rank"] = ids, agent, index, tgt, rank
            rows["BinaryProb"], rows["BinaryVec"] = cpu(probs), cpu(vec)
            comm.append(rows)
        rows = np.zeros(bsz, dtype=pdt)
        rows["ExampleId"], rows["AgentId"], rows["Index"], rows["Target"], rows["Rank"] = ids, "R", 2 * t + 1, tgt, rank
        rows["Predictions"] = np_preds
        rows["StopProb"], rows["StopVec"], rows["StopMask"] = cpu(sp).reshape(bsz, 1), cpu(sf).reshape(bsz, 1), cpu(sm).reshape(bsz, 1)
        preds.append(rows)
    return np.concatenate(comm), np.concatenate(preds)


def extract_binary(FLAGS, load_hdf5, exchange, dev_file, batch_size, epoch, shuffle, cuda, top_k, sender, receiver, desc_dict,
                   map_labels, file_name, store=None):
    """binary_vectors.extract_binary with the reference's signature (binary_vectors.py:12-13): runs eval-mode exchanges over
    the development set and writes the "Communication" and "Predictions" datasets to FLAGS.binary_output.  `exchange` is
    `multimodalgame_b200.model.exchange` (the fused kernels); `store` overrides the output container (tests)."""
    desc = desc_dict["desc"]
    out = store if store is not None else open_write_store(FLAGS.binary_output)
    dev_loader = load_hdf5(dev_file, batch_size, epoch, shuffle, truncate_final_batch=True, map_labels=map_labels)
    for batch in dev_loader:
        target, data = batch["target"], batch[FLAGS.img_feat]
        if cuda:
            data, target, desc = data.cuda(), target.cuda(), desc.cuda()
        exchange_args = dict(data=data, target=target, desc=desc, desc_set=desc_dict.get("desc_set", None),
                             desc_set_lens=desc_dict.get("desc_set_lens", None), train=False,
                             break_early=not FLAGS.fixed_exchange)
        if getattr(FLAGS, "attn_extra_context", False):
            exchange_args["data_context"] = batch[FLAGS.data_context]
        res = exchange(sender, receiver, None, None, exchange_args)
        comm, preds = communication_records(res, batch["example_ids"], target, int(desc.shape[0]))
        out.append("Communication", comm)
        out.append("Predictions", preds)
    if store is None:
        out.close()
    return out
