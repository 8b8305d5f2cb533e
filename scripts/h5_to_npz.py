"""Convert a Keras 2.1.x weight file (`model.save_weights('..._weights_best.h5')`, the files
/root/reference/multimodal_fusion/multimodal.py:63-66 loads) into the ordered .npz that
`mgr_b200.keras_io.load_weights` reads.  Needs h5py, i.e. runs on the machine that has the reference's
environment (h5py is not part of the B200 image).

    python scripts/h5_to_npz.py sp_ctc_lstm_weights_best.h5 sp.npz

Layers are taken in `f.attrs['layer_names']` order and, inside a layer, in `g.attrs['weight_names']` order --
the order Keras itself uses for topological loading; layers without weights are skipped."""
import sys

import numpy as np


def main(src, dst):
    import h5py
    out = {}
    with h5py.File(src, "r") as f:
        g0 = f["model_weights"] if "model_weights" in f else f      # model.save() vs model.save_weights()
        i = 0
        for layer in g0.attrs["layer_names"]:
            layer = layer.decode() if isinstance(layer, bytes) else layer
            g = g0[layer]
            for w in g.attrs["weight_names"]:
                w = w.decode() if isinstance(w, bytes) else w
                out["%03d|%s" % (i, w)] = np.asarray(g[w], dtype=np.float32)
                i += 1
    np.savez(dst, **out)
    print("wrote %d arrays to %s" % (len(out), dst))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
