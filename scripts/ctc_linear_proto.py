"""fp32 NumPy model of the linear-domain CTC kernel (ctc_lin.cu): scaled alpha from the left, scaled
"gamma" (= beta with the emission folded in) from the right, power-of-two renormalisation every R steps,
meet in the middle, occupancies on the way through the other half, per-row consistency check
(sum_u occupancy == 1).  Used to settle the numerics on CPU before spending GPU time; compares against
oracle/ctc_ref.py.  Run: python scripts/ctc_linear_proto.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import ctc_ref
from helpers import random_probs, random_labels, peaky_probs

f32 = np.float32
TARGET = 50
R = 4


def expo(x):  # floor(log2 x) of a positive normal fp32
    return int((np.float32(x).view(np.uint32) >> 23) & 0xFF) - 127


D_LIP = 54
EMPTY = -(1 << 20)


def pow2(k):
    """fp32 2^k with the kernel's clamp: k < -126 -> 0, k > 127 -> 2^127."""
    k = np.asarray(k)
    return np.where(k < -126, f32(0), np.exp2(np.minimum(k, 127).astype(np.float64))).astype(f32)


def run_dir(pe, labs, blank, Tn, K, norm_rows):
    """Recursion in processing order over rows pe[0..Tn).  Lane j owns blank states b[jK..jK+K) and label
    states l[jK..jK+K) and one exponent E[j] (block floating point per lane); values flow from lane j-1 to
    lane j rescaled by 2^(E[j-1]-E[j]).  norm_rows[i] says whether the states are renormalised after
    computing row i.  Returns per-row state arrays (scaled) and per-row exponent vectors."""
    L = len(labs)
    NL = 32
    S = NL * K
    b = np.zeros(S, f32); l = np.zeros(S, f32)
    pl_idx = np.zeros(S, np.int64); vl = np.zeros(S, bool); vl[:L] = True
    pl_idx[:L] = labs
    skip = np.zeros(S, bool)
    if L > 1:
        skip[1:L] = np.asarray(labs[1:]) != np.asarray(labs[:-1])
    E = np.zeros(NL, np.int64)
    lane_of = np.arange(S) // K
    rows_b, rows_l, exps = [], [], []
    for i in range(Tn):
        pb = pe[i, blank]
        pl = np.where(vl, pe[i, pl_idx], f32(0)).astype(f32)
        if i == 0:
            b[0] = pb
            if L: l[0] = pl[0]
        else:
            # value arriving from the previous lane: l[jK-1] * 2^(E[j-1]-E[j])
            fac = np.zeros(NL, f32); fac[1:] = pow2(E[:-1] - E[1:])
            prevl = np.zeros(S, f32); prevl[1:] = l[:-1]
            first = (np.arange(S) % K) == 0
            prevl = np.where(first, (prevl * fac[lane_of]).astype(f32), prevl)
            tb = (b + prevl).astype(f32)
            nb = (tb * pb).astype(f32)
            nl = ((l + np.where(skip, tb, b)).astype(f32) * pl).astype(f32)
            b, l = nb, nl
        assert np.all(np.isfinite(b)) and np.all(np.isfinite(l))
        if norm_rows[i]:
            m = np.maximum(b.reshape(NL, K).max(axis=1), l.reshape(NL, K).max(axis=1))
            e = np.array([expo(x) if x > 0 else 0 for x in m])
            own = np.where(m > 0, E + e - TARGET, EMPTY)
            # Lipschitz floor: E[j] >= E[j-1] - D  (prefix max of own[i] - D*(j-i))
            newE = own.copy()
            pm = EMPTY
            for j in range(1, NL):
                pm = max(pm, own[j - 1])
                newE[j] = max(own[j], pm - D_LIP)
            sc = pow2(E - newE)
            b = (b * sc[lane_of]).astype(f32); l = (l * sc[lane_of]).astype(f32)
            E = newE
        rows_b.append(b.copy()); rows_l.append(l.copy()); exps.append(E.copy())
    return rows_b, rows_l, exps


def ctc_linear(pe, labs, Tn, check_tol=1e-3, R=8):
    """pe: (Tn, C) fp32 = p + eps (unnormalised).  Returns loss, dz = q - occ (Tn, C), flagged, worst check."""
    pe = pe.astype(f32)
    C = pe.shape[1]; blank = C - 1; L = len(labs)
    K = (L + 1 + 31) // 32
    Z = pe.sum(axis=1, dtype=f32)
    tstar = (Tn // 2) & ~(R - 1)
    labs = list(labs)
    t = np.arange(Tn)
    norm0 = (t % R == 0)                       # dir 0, processing index i = t
    norm1 = ((Tn - 1 - t) % R == R - 1); norm1[0] = True   # dir 1, i = Tn-1-t: natural t%R == R-1, and the init row
    ab, al, ae = run_dir(pe, labs, blank, Tn, K, norm0)
    gb, gl, ge = run_dir(pe[::-1], labs[::-1], blank, Tn, K, norm1)
    S = 32 * K
    lane_of = np.arange(S) // K
    kb = np.arange(L + 1); kl = np.arange(L)
    def row(t):
        a_b, a_l = ab[t][:L + 1], al[t][:L]
        Ea_b, Ea_l = ae[t][lane_of[kb]], ae[t][lane_of[kl]]
        i = Tn - 1 - t
        g_b, g_l = gb[i][L - kb], gl[i][L - 1 - kl]
        Eg_b, Eg_l = ge[i][lane_of[L - kb]], ge[i][lane_of[L - 1 - kl]]
        return a_b, a_l, Ea_b, Ea_l, g_b, g_l, Eg_b, Eg_l
    a_b, a_l, Ea_b, Ea_l, g_b, g_l, Eg_b, Eg_l = row(tstar)
    with np.errstate(divide="ignore"):
        vb = np.log2(a_b).astype(f32) + np.log2(g_b).astype(f32) - np.log2(pe[tstar, blank])
        vlv = np.log2(a_l).astype(f32) + np.log2(g_l).astype(f32) - np.log2(pe[tstar, labs]) if L else np.zeros(0, f32)
    Eall = np.concatenate([Ea_b + Eg_b, Ea_l + Eg_l]); vall = np.concatenate([vb, vlv]).astype(f32)
    ok = np.isfinite(vall)
    if not ok.any():
        return np.inf, None, True, 1.0
    M = int(Eall[ok].max())
    v = np.where(ok, vall + (Eall - M).astype(f32), f32(-1e30)).astype(f32)
    m = v.max(); ssum = np.exp2((v - m).astype(f32)).astype(f32).sum(dtype=f32)
    logP2 = float(M) + float(m) + float(np.log2(f32(ssum)))
    loss = -(logP2 * np.log(2.0)) + float(np.log(Z.astype(np.float64)).sum())
    dz = np.zeros((Tn, C), np.float64)
    flagged = False
    kfl = int(np.floor(logP2)); mant = f32(2.0 ** -(logP2 - kfl))
    worst = 0.0
    for t in range(Tn):
        a_b, a_l, Ea_b, Ea_l, g_b, g_l, Eg_b, Eg_l = row(t)
        def occ_of(a, g, Ea, Eg, rpe):
            kexp = np.clip(Ea + Eg - kfl, -252, 252)
            k1 = kexp >> 1; k2 = kexp - k1
            c1 = pow2(k1); c2 = (pow2(k2) * mant).astype(f32)
            return ((a * c1).astype(f32) * (g * c2).astype(f32)).astype(f32) * rpe
        ob = occ_of(a_b, g_b, Ea_b, Eg_b, f32(1) / pe[t, blank]).astype(f32)
        occ = np.zeros(C, f32)
        occ[blank] = ob.sum(dtype=f32)
        if L:
            ol = occ_of(a_l, g_l, Ea_l, Eg_l, (f32(1) / pe[t, labs]).astype(f32)).astype(f32)
            np.add.at(occ, labs, ol)
        if not np.all(np.isfinite(occ)):
            return loss, None, True, np.inf
        s = float(occ.sum(dtype=f32))
        worst = max(worst, abs(s - 1))
        if not abs(s - 1) <= check_tol:
            flagged = True
        dz[t] = pe[t] / Z[t] - occ
    return loss, dz, flagged, worst


def compare(p, labels, il, ll, eps=1e-8, name=""):
    B, T, C = p.shape
    ref_loss, ref_g = ctc_ref.ctc_lambda_func((p, labels, il, ll), want_grad=True)
    worst_l = worst_g = worst_chk = 0; nflag = 0
    for b in range(B):
        Tn = int(il[b, 0])
        seq = ctc_ref.prepare_label_sequence(labels[b, :ll[b, 0]], C)
        pe = (p[b, 2:2 + Tn].astype(f32) + f32(eps)).astype(f32)
        out = ctc_linear(pe, seq, Tn)
        loss, dz, flagged = out[0], out[1], out[2]
        if flagged:
            nflag += 1
            if dz is None: continue
        worst_chk = max(worst_chk, out[3])
        g = dz / pe.astype(np.float64)
        rg = ref_g[b, 2:2 + Tn]
        scale = np.abs(rg).max(axis=1, keepdims=True) + 1e-12
        worst_g = max(worst_g, (np.abs(g - rg) / scale).max())
        worst_l = max(worst_l, abs(loss - ref_loss[b, 0]) / abs(ref_loss[b, 0]))
    print("%-40s loss rel %.2e  grad (row-scaled) %.2e  check dev %.2e  flagged %d/%d" % (name, worst_l, worst_g, worst_chk, nflag, B))


if __name__ == "__main__":
    for (B, T, C, Lmax) in [(4, 12, 5, 3), (8, 50, 22, 10), (3, 33, 44, 40), (5, 130, 22, 35), (2, 70, 6, 33), (2, 300, 22, 150)]:
        rng = np.random.default_rng(B * 1000 + T)
        p, _ = random_probs(rng, B, T, C)
        il = rng.integers((T - 2) // 2 + 1, T - 1, size=(B, 1)); il[0, 0] = T - 2
        labels, ll = random_labels(rng, B, Lmax, C, T_avail=il[:, 0])
        compare(p, labels, il, ll, name="random B%d T%d C%d L%d" % (B, T, C, Lmax))
    rng = np.random.default_rng(5)
    B, T, C, Lmax = 6, 1002, 22, 40
    for scale in (2.0, 6.0, 20.0):
        p, _ = random_probs(rng, B, T, C, scale=scale)
        il = np.full((B, 1), T - 2); labels, ll = random_labels(rng, B, Lmax, C)
        compare(p, labels, il, ll, name="random T1000 L40 logit-scale %g" % scale)
    for sharp in (4.0, 12.0, 40.0):
        p = peaky_probs(rng, B, T, C, sharp=sharp)
        il = np.full((B, 1), T - 2); labels, ll = random_labels(rng, B, Lmax, C)
        compare(p, labels, il, ll, name="peaky(sharp %g) vs random labels" % sharp)
