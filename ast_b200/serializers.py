"""chainer.serializers.save_npz / load_npz with the reference's key set:
numpy.savez_compressed, keys = slash-joined link paths, BN persistents avg_mean / avg_var / N
included, LSTM h/c and mask_pad_id not saved.  Reference-trained seq2seq_<N>.model files load."""
import numpy as np


def save_npz(file, obj, compression=True):
    arrays = obj.state_arrays()
    with open(file, "wb") as f:
        (np.savez_compressed if compression else np.savez)(f, **arrays)


def load_npz(file, obj, path="", strict=True):
    with np.load(file) as npz:
        arrays = {k[len(path):] if path and k.startswith(path) else k: npz[k] for k in npz.files}
    feat_dim = None
    if getattr(obj, "_engine", None) is None:
        # lazily-shaped model (in_channels: null): the checkpoint fixes the feature dimension
        r_in = arrays["L0_enc/upward/W"].shape[1]
        l0 = obj.cfg["cnn_config"]["cnn_layers"][0]
        c_last = obj.cfg["cnn_config"]["cnn_layers"][-1]["out_channels"]
        fp = r_in // c_last
        feat_dim = (fp - 1) * l0["stride"][1] + l0["ksize"][1]
        obj._build(feat_dim)
    obj.load_state(arrays)
    return obj
