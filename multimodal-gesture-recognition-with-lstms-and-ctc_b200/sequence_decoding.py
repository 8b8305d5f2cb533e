"""`decode_batch` and `ctc_decode` on the B200 decode kernels.

`decode_batch(pred_out, f_list)` mirrors /root/reference/multimodal_fusion/sequence_decoding.py:21-69
(threshold 0.5, 22-gesture map, `final_ctc_recout.mlf`); `decode_batch_speech` mirrors
/root/reference/audio_network/sequence_decoding.py:19-69 (threshold 0.75, 44-word map supplied by
the caller, `Sample%05d_audio`, `ctc_recout.mlf`).  Same return value (list[list[str]]) and the
same MLF side effect; the numeric part (argmax / count-based confidence filter / collapse) runs
in `gr_ctc_bestpath_ref_f32`.  `ctc_decode` mirrors Keras `K.ctc_decode` (greedy / beam 100).
"""
import numpy as np
import torch

from . import ops

MAP_GEST = {0: "oov", 1: "VA", 2: "VQ", 3: "PF", 4: "FU", 5: "CP", 6: "CV",
            7: "DC", 8: "SP", 9: "CN", 10: "FN", 11: "OK", 12: "CF", 13: "BS",
            14: "PR", 15: "NU", 16: "FM", 17: "TT", 18: "BN", 19: "MC",
            20: "ST", 21: "sil"}
IGNORE_LIST = [228, 298, 299, 300, 303, 304, 334, 343, 373, 375]


def _to_device(pred_out):
    t = torch.as_tensor(pred_out)
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.pin_memory().cuda(non_blocking=True) if torch.cuda.is_available() else t.cuda()
    return t.contiguous()


def decode_ids(pred_out, threshold=0.5, drop_frames=2):
    """(N,T,C) softmax -> list of int id lists (blank kept), the reference's filter semantics."""
    ids, lens = ops.bestpath_ref(_to_device(pred_out), threshold, drop_frames)
    ids = ids.cpu().numpy()
    lens = lens.cpu().numpy()
    return [ids[j, :lens[j]].tolist() for j in range(ids.shape[0])]


def decode_batch(pred_out, f_list, threshold=0.5, map_gest=None, mlf_path="final_ctc_recout.mlf",
                 ignore_list=IGNORE_LIST, name_fmt="Sample%s"):
    map_gest = MAP_GEST if map_gest is None else map_gest
    all_ids = decode_ids(pred_out, threshold)
    ret = []
    lines = ["#!MLF!#\n"]
    for j, ids in enumerate(all_ids):
        outstr = [map_gest[i] for i in ids]
        ret.append(outstr)
        f_num = f_list[j]
        if int(f_num) in ignore_list:
            continue
        lines.append('"*/%s.rec"\n' % (name_fmt % format(f_num, "05")))
        lines.extend("%s\n" % cl for cl in outstr)
        lines.append(".\n")
    if mlf_path is not None:
        with open(mlf_path, "w") as of:
            of.writelines(lines)
    return ret


def decode_batch_speech(pred_out, f_list, map_gest, threshold=0.75, mlf_path="ctc_recout.mlf", ignore_list=()):
    return decode_batch(pred_out, f_list, threshold=threshold, map_gest=map_gest, mlf_path=mlf_path,
                        ignore_list=ignore_list, name_fmt="Sample%s_audio")


def ctc_decode(y_pred, input_length, greedy=True, beam_width=100, top_paths=1, merge_repeated=True, eps=1e-8):
    """Keras `K.ctc_decode`: returns ([decoded (N, max decoded length) int64 padded -1] * top_paths,
    log_prob (N, top_paths)) -- the dense form of the SparseTensor Keras builds, so the width is the longest decoded
    sequence, not T.  merge_repeated defaults to TF 1.12's ctc_beam_search_decoder default (Keras 2.1.4 does not pass it).
    Beam log_prob: every frame is fully normalised (log_softmax of log(p + eps)) before the search; the pinned TF
    1.12 kernel is believed to subtract only the per-frame maximum, which shifts log_prob by sum_t(lse_t - max_t) and
    cannot change the decoded labels (a per-frame constant) -- DESIGN.md section 2."""
    p = _to_device(y_pred)
    N = p.shape[0]
    sl = torch.as_tensor(input_length).reshape(N).to(device=p.device, dtype=torch.int32)

    def trim(ids2d, lens1d):
        w = max(int(lens1d.max().item()) if lens1d.numel() else 0, 1)
        return ids2d[:, :w].to(torch.int64).contiguous()
    if greedy:
        ids, lens, score = ops.greedy(p, sl, eps)
        return [trim(ids, lens)], score.reshape(N, 1)
    ids, lens, logp = ops.beam(p, sl, beam_width, top_paths, merge_repeated, eps)
    return [trim(ids[:, k], lens[:, k] if lens.dim() == 2 else lens) for k in range(top_paths)], logp
