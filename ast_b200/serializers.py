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
    if getattr(obj, "_engine", None) is None:
        # lazily-shaped model (in_channels: null).  The checkpoint fixes only the number of CNN frequency positions
        # F' = L0_enc input width / C1, not the feature dimension D (any D with (D + 2p - kw) // sw + 1 == F' has the same
        # parameter shapes: D = 39 and D = 40 both give F' = 3 with the shipped kw = sw = 13).  Shape the model for the
        # smallest such D so that links are addressable (copy_params.py:26-65); the first encode() re-shapes it for the
        # real D of the data if that differs (SpeechEncoderDecoder._require).
        r_in = arrays["L0_enc/upward/W"].shape[1]
        l0 = obj.cfg["cnn_config"]["cnn_layers"][0]
        c_last = obj.cfg["cnn_config"]["cnn_layers"][-1]["out_channels"]
        fp = r_in // c_last
        obj._build((fp - 1) * l0["stride"][1] + l0["ksize"][1] - 2 * l0["pad"][1], provisional=True)
    obj.load_state(arrays)
    return obj
