"""ORACLE (test infrastructure, not product code) -- stage 3, EM VP localisation.

A float64 numpy restatement of the reference's expectation-maximisation stage
(reference vp_localisation.py + probability_functions.py), written from the
algorithm, vectorised where the reference uses Python loops / joblib pools.
Every function cites the reference file:line it follows.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product path never does.

PARITY STATUS: *pinned* -- oracle/make_golden.py imports the reference's own
modules from /root/reference (after the five mechanical Python-3 patches of
SURVEY.md section 8(c), applied in a temp dir) and stores its inputs/outputs
under tests/golden/; tests/test_oracle_em.py checks this restatement against
those vectors.  Third-party pieces that stay unpinned: scikit-learn's
AgglomerativeClustering (reference pins 0.18, here 1.9) and LAPACK SVD.
"""
from collections import namedtuple

import numpy as np

PI = np.pi

PDFParams = namedtuple("PDFParams", "means weights sigma")
PDF = namedtuple("PDF", "v lv vl l lvsq angles")


# --------------------------------------------------------------------------
# segment-pair geometry  (vp_localisation.py:700-762)
# --------------------------------------------------------------------------
def _psd(ax, ay, bx, by, px, py):
    """Point-to-segment distance, vp_localisation.py:743-758 (broadcasting).
    Segment a->b, point p.  Note the reference squares the *norm* (:747)."""
    dx, dy = bx - ax, by - ay
    nrm = np.sqrt(dx * dx + dy * dy)
    with np.errstate(divide="ignore", invalid="ignore"):
        param = ((px - ax) * dx + (py - ay) * dy) / np.square(nrm)
    cx = np.where(param < 0, ax, np.where(param > 1, bx, ax + param * dx))
    cy = np.where(param < 0, ay, np.where(param > 1, by, ay + param * dy))
    ex, ey = cx - px, cy - py
    return np.sqrt(ex * ex + ey * ey)


def segment_distance(A, B):
    """line_distance_closest, vp_localisation.py:727-740.  A (K,4) broadcast
    against B (L,4) -> (K,L)."""
    a = [A[:, k][:, None] for k in range(4)]
    b = [B[:, k][None, :] for k in range(4)]
    d1 = _psd(a[0], a[1], a[2], a[3], b[0], b[1])
    d2 = _psd(a[0], a[1], a[2], a[3], b[2], b[3])
    d4 = _psd(b[0], b[1], b[2], b[3], a[0], a[1])
    d5 = _psd(b[0], b[1], b[2], b[3], a[2], a[3])
    return np.minimum(np.minimum(d1, d2), np.minimum(d4, d5))


def cosangle(A, B, f):
    """lines_points_cosangle, vp_localisation.py:715-724 -> (K,L)."""
    v1x, v1y = (A[:, 0] - A[:, 2])[:, None], (A[:, 1] - A[:, 3])[:, None]
    v2x, v2y = (B[:, 0] - B[:, 2])[None, :], (B[:, 1] - B[:, 3])[None, :]
    n1 = np.sqrt(v1x * v1x + v1y * v1y)
    n2 = np.sqrt(v2x * v2x + v2y * v2y)
    c = np.abs((v1x * v2x + v1y * v2y) / (n1 * n2))
    dphi = np.abs(np.arccos(np.clip(c, -1, 1)))
    return np.cos(np.clip(f * dphi, -PI / 2, PI / 2))


def seg_length(lp):
    """line_length, vp_localisation.py:761-762."""
    dx, dy = lp[:, 0] - lp[:, 2], lp[:, 1] - lp[:, 3]
    return np.sqrt(dx * dx + dy * dy)


def proximity(A, B, sigma, d=None):
    """lines_proximity, vp_localisation.py:708-712 -> (K,L)."""
    sg = sigma * np.minimum(seg_length(A)[:, None], seg_length(B)[None, :])
    if d is None:
        d = segment_distance(A, B)
    return np.exp(-(d * d) / (2 * sg * sg))


def calc_lsim(lp, sigma=0.1, chunk=512):
    """calc_lsim + lines_similarity, vp_localisation.py:87-108, 700-705.
    Symmetric (N,N), zero diagonal."""
    N = lp.shape[0]
    out = np.zeros((N, N))
    for i0 in range(0, N, chunk):
        A = lp[i0:i0 + chunk]
        out[i0:i0 + chunk] = cosangle(A, lp, 9) * proximity(A, lp, sigma)
    np.fill_diagonal(out, 0.0)
    return out


def line_rating_knn(lp, k1=10, k2=3, sigma=1, chunk=512):
    """vp_localisation.py:34-84.  ldist diag = 4 (:82); k1 nearest by
    distance, of those the k2 with largest cosangle(f=9), score = sum(prox *
    cos)/k2."""
    N = lp.shape[0]
    k1 = min(k1, N)
    k2 = min(k2, N)
    lscore = np.zeros(N)
    for i0 in range(0, N, chunk):
        A = lp[i0:i0 + chunk]
        K = A.shape[0]
        d_true = segment_distance(A, lp)
        d = d_true.copy()
        d[np.arange(K), np.arange(i0, i0 + K)] = 4                  # :82
        near = np.argsort(d, axis=1)[:, 0:k1]                       # :47-48
        rows = np.arange(K)[:, None]
        cosphi = cosangle(A, lp, 9)[rows, near]                     # :55
        best = np.argsort(cosphi, axis=1)[:, ::-1][:, 0:k2]         # :57-59
        nb = near[rows, best]
        prox = proximity(A, lp, sigma, d=d_true)                    # :65 recomputes the true distance
        lscore[i0:i0 + K] = np.sum(prox[rows, nb] * cosphi[rows, best], axis=1)
    return lscore / k2


def lines_angles(lp):
    """vp_localisation.py:765-776."""
    vx, vy = lp[:, 0] - lp[:, 2], lp[:, 1] - lp[:, 3]
    vx = vx / np.sqrt(vx * vx + vy * vy)
    phi = np.abs(np.arccos(np.clip(vx, -1, 1)))
    return np.where(phi > PI / 2, PI - phi, phi)


# --------------------------------------------------------------------------
# initialisation  (E0, E1, E2)
# --------------------------------------------------------------------------
def find_maxima(resp):
    """vp_localisation.py:13-31 incl. the `a-1 > 0` / `b-1 > 0` border quirk:
    the neighbour at index 0 (and the out-of-range ones) compare as 0."""
    B, A = resp.shape
    z = np.zeros_like(resp)
    vu = z.copy(); vu[:, :-1] = resp[:, 1:]
    vd = z.copy(); vd[:, 2:] = resp[:, 1:-1]
    vl = z.copy(); vl[2:, :] = resp[1:-1, :]
    vr = z.copy(); vr[:-1, :] = resp[1:, :]
    return ((resp > vu) & (resp > vd) & (resp > vl) & (resp > vr)).astype(np.float64)


def _index_to_angle(index, shape):
    """coordinate_conversion.py:4-20."""
    return np.array([(index[0] - 0.5 * shape[0] + 0.5) * PI / shape[0],
                     (index[1] - 0.5 * shape[1] + 0.5) * PI / shape[1]])


def _angle_to_point(angle):
    """coordinate_conversion.py:38-50."""
    p = np.array([np.sin(angle[0]) * np.cos(angle[1]), np.sin(angle[1]),
                  np.cos(angle[0]) * np.cos(angle[1])])
    return p * np.sign(p[2])


def find_initial_vps(sphere_image, resp, num_max):
    """vp_localisation.py:111-165.  Raises ValueError (np.vstack([])) when no
    candidate survives, like the reference."""
    sphere = sphere_image[::-1, :]
    rA, rB = resp.shape
    sA, sB = sphere_image.shape
    maxima = find_maxima(resp).flatten()
    flat = resp.flatten()
    idx = np.where(maxima == 1)[0]
    # Equal responses (a saturated sigmoid gives exact float32 ties) are ordered by numpy.argsort's default
    # sort in the reference (:123), i.e. by whatever the numpy build does with ties; the stable order is used
    # here and in the kernel (csrc/em_core.cuh, init_prior_and_vps): of equal maxima the later cell wins.
    order = np.argsort(flat[idx], kind="stable")[::-1]
    maxima[idx[order[num_max:]]] = 0
    maxima = maxima.reshape(resp.shape)
    vps = []
    for ra in range(rA):
        for rb in range(rB):
            if maxima[ra, rb] != 1:
                continue
            r0, c0 = ra * sA // rA, rb * sB // rB
            sl = sphere[r0:(ra + 1) * sA // rA, c0:(rb + 1) * sB // rB]
            mx = sl.max()
            rr, cc = np.nonzero((sl >= mx) & (sl > 0))
            if rr.size == 0:
                continue
            avg_r = rr.sum() / float(rr.size)
            avg_c = cc.sum() / float(cc.size)
            idx2 = np.array([avg_c + c0, avg_r + r0])            # :157-158
            vps.append(_angle_to_point(_index_to_angle(idx2, sphere_image.shape)))
    return np.vstack(vps)


def pdf_params(resp, confidence=1.282):
    """probability_functions.py:62-96."""
    A, B = resp.shape
    sigma = PI / (confidence * A)
    alphas = np.tile(np.linspace(-(A - 1.0) / A * PI / 2, (A - 1.0) / A * PI / 2, A), (B, 1)).flatten()
    betas = np.tile(np.linspace(-(B - 1.0) / B * PI / 2, (B - 1.0) / B * PI / 2, B), (A, 1)).T.flatten()
    weights = resp.flatten()
    order = np.argsort(weights)[::-1]
    weights[order[100:]] = 0
    weights /= np.sum(weights)
    weights /= (2 * PI * sigma * sigma)
    return PDFParams(means=np.stack([alphas, betas], axis=1), weights=weights, sigma=sigma)


# --------------------------------------------------------------------------
# E-step  (E5)
# --------------------------------------------------------------------------
def calc_angles(v):
    """probability_functions.py:252-259."""
    beta = np.arcsin(v[:, 1])
    inner = np.maximum(np.minimum(v[:, 0] / np.cos(beta), 1), -1)
    return np.stack([np.arcsin(inner), beta], axis=1)


def calc_pdf(pdfpar, x, y):
    """probability_functions.py:8-40, incl. the duplicated 4th/5th wrapped
    copy (:25-26).  Accumulates over components in index order like :38."""
    resp = np.zeros(x.shape[0])
    k = -0.5 / (pdfpar.sigma * pdfpar.sigma)
    for n in np.nonzero(pdfpar.weights > 0)[0]:
        mx, my = pdfpar.means[n, 0], pdfpar.means[n, 1]
        d1 = (x - mx) ** 2 + (y - my) ** 2
        d2 = (x - mx + PI) ** 2 + (y + my) ** 2
        d3 = (x - mx - PI) ** 2 + (y + my) ** 2
        d4 = (x + mx) ** 2 + (y - my - PI) ** 2
        p = (((np.exp(d1 * k) + np.exp(d2 * k)) + np.exp(d3 * k)) + np.exp(d4 * k)) + np.exp(d4 * k)
        resp += p * pdfpar.weights[n]
    return resp


def calc_lvsq_angle(v, lp):
    """probability_functions.py:157-176 -> (N,M).  v is (M,3)."""
    vx = (v[:, 0] / v[:, 2])[None, :]
    vy = (v[:, 1] / v[:, 2])[None, :]
    mx = (0.5 * (lp[:, 0] + lp[:, 2]))[:, None]
    my = (0.5 * (lp[:, 1] + lp[:, 3]))[:, None]
    ax, ay = mx - vx, my - vy
    bx, by = (lp[:, 0] - lp[:, 2])[:, None], (lp[:, 1] - lp[:, 3])[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        c = (ax * bx + ay * by) / (np.sqrt(ax * ax + ay * ay) * np.sqrt(bx * bx + by * by))
    return (1 - np.abs(c)) ** 2


def calc_probabilities(pdfpar, v, lp, s):
    """probability_functions.py:99-147 (angle mode).  Mutates s (floor 1e-200,
    :139) like the reference."""
    angles = calc_angles(v)
    p_v = calc_pdf(pdfpar, angles[:, 0], angles[:, 1])
    lvsq = calc_lvsq_angle(v, lp)
    np.maximum(s, 1e-200, out=s)
    with np.errstate(under="ignore"):
        p_lv = np.exp(-(lvsq / (2 * s)[None, :])) * (1.0 / np.sqrt(2 * PI * s))[None, :]
    p_l = np.maximum(np.dot(p_lv, p_v), 1e-12)
    with np.errstate(under="ignore"):
        p_vl = (p_lv * p_v[None, :]).T / p_l[None, :]
    return PDF(v=p_v, lv=p_lv, vl=p_vl, l=p_l, lvsq=lvsq, angles=angles)


def calc_lvsq_single(vp, lp):
    """probability_functions.py:212-224, vectorised over lines with a
    per-line VP (vp (N,3), lp (N,4))."""
    vx, vy = vp[:, 0] / vp[:, 2], vp[:, 1] / vp[:, 2]
    ax, ay = 0.5 * (lp[:, 0] + lp[:, 2]) - vx, 0.5 * (lp[:, 1] + lp[:, 3]) - vy
    bx, by = lp[:, 0] - lp[:, 2], lp[:, 1] - lp[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        c = (ax * bx + ay * by) / (np.sqrt(ax * ax + ay * ay) * np.sqrt(bx * bx + by * by))
    return (1 - np.abs(c)) ** 2


# --------------------------------------------------------------------------
# M-step pieces  (E6, E7, E9)
# --------------------------------------------------------------------------
def weight_matrix(p_vl, lweight, lsim, bias=0.001, lsim_colsum=None):
    """vp_localisation.py:515-524 as one (M,N)x(N,N) product."""
    w_ = p_vl * lweight[None, :]
    if lsim_colsum is None:
        lsim_colsum = np.sum(lsim, axis=0)
    with np.errstate(under="ignore"):
        return (w_ + (bias * lweight)[None, :] * np.dot(w_, lsim)) / (1 + bias * lweight * lsim_colsum)[None, :]


def calc_new_vanishing_point(l, w):
    """vp_localisation.py:453-479 (thin SVD: same right singular vectors)."""
    if np.size(w) == 0 or np.max(w) == 0:
        return None
    try:
        _, _, Vt = np.linalg.svd((w / np.max(w))[:, None] * l, full_matrices=False)
    except np.linalg.LinAlgError:
        return None
    if Vt.shape[0] < 3:
        # fewer than 3 rows: the reference's full SVD still returns a 3x3 V
        _, _, Vt = np.linalg.svd((w / np.max(w))[:, None] * l, full_matrices=True)
    vp = Vt[2, :].copy()
    vp /= np.linalg.norm(vp, ord=2)
    vp *= np.sign(vp[2])
    return vp


def calc_vp_line_counts(vp, lp, s, decision_metric, lweights, thresh):
    """vp_localisation.py:482-512 (angle mode)."""
    M = vp.shape[0]
    assoc = np.argmax(decision_metric, axis=0)
    dist = calc_lvsq_single(vp[assoc], lp)
    with np.errstate(invalid="ignore"):
        outlier = (dist > thresh * np.sqrt(s[assoc])) | (lweights == 0)
    assoc = np.where(outlier, -1, assoc)
    counts = np.zeros(M)
    counts_weighted = np.zeros(M)
    for m in range(M):
        sel = assoc == m
        counts[m] = np.count_nonzero(sel)
        counts_weighted[m] = np.sum(lweights[sel])
    return counts, counts_weighted, assoc


def _s_update(lvsq_col, pvl_row):
    """vp_localisation.py:301-304: exp(log(sum(lvsq*p_vl)) - log(sum(p_vl)))."""
    with np.errstate(divide="ignore", invalid="ignore", under="ignore"):
        return np.exp(np.log(np.sum(lvsq_col * pvl_row)) - np.log(np.sum(pvl_row)))


# --------------------------------------------------------------------------
# split / merge  (E10, E11)
# --------------------------------------------------------------------------
def average_linkage_two_clusters(D):
    """What the reference obtains from sklearn.cluster.AgglomerativeClustering(
    linkage='average', connectivity=D, n_clusters=2, affinity='precomputed')
    (vp_localisation.py:574-578), via scikit-learn itself (unpinned version)."""
    import warnings
    import sklearn.cluster as cluster
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = cluster.AgglomerativeClustering(linkage="average", connectivity=D, n_clusters=2,
                                                metric="precomputed")
        model.fit_predict(D)
    return model.labels_


def split_best_vp(v_cur, s, lp, l, w, lweight, langles, min_diff, clusterer=average_linkage_two_clusters):
    """vp_localisation.py:527-630.  Returns (v_cur, s, n_added); the caller
    keeps v_next in step (appended rows are zero there, :626)."""
    M = v_cur.shape[0]
    assoc = np.argmax(w, axis=0)
    wg = w[assoc, np.arange(w.shape[1])] / w.max()
    std = np.full(M, np.nan)
    for m in range(M):
        sel = (assoc == m) & (wg > 0)
        if np.any(sel):
            std[m] = np.std(langles[sel])
    order = np.argsort(std)[::-1]                                # :546-547 (NaN first)
    worst = None
    for m in range(M):
        lines_w = np.where(assoc == order[m])[0]
        vp = v_cur[m, :] / v_cur[m, 2]                           # :557 quirk: row m, not order[m]
        if lines_w.size > 8 and (-1 < vp[0] < 1) and (-1 < vp[1] < 1):
            worst = order[m]
            break
    if worst is None:
        return v_cur, s, 0
    stdd = s[worst] / 2
    lpw = lp[lines_w]
    D = 1 - cosangle(lpw, lpw, 2)
    np.fill_diagonal(D, 0.0)
    labels = clusterer(D)
    lw = l[lines_w] * lweight[lines_w][:, None]
    new_vps = []
    for c in range(2):
        ls = lw[labels == c]
        if ls.shape[0] < 3:
            continue
        _, _, Vt = np.linalg.svd(ls)
        vp = Vt[2, :] / np.linalg.norm(Vt[2, :], ord=2)
        if vp[2] < 0:
            vp = -vp
        new_vps.append(vp)
    too_similar = True
    for c in range(len(new_vps)):
        for d in range(c + 1, len(new_vps)):
            cp = np.clip(np.dot(new_vps[c], new_vps[d]), -1, 1)
            if np.abs(np.arccos(np.clip(np.abs(cp), -1, 1))) > min_diff:
                too_similar = False
    if too_similar:
        return v_cur, s, 0
    v_cur = v_cur.copy()
    s = s.copy()
    v_cur[worst] = new_vps[0]
    s[worst] = stdd
    added = 0
    for vp in new_vps[1:]:
        v_cur = np.vstack([v_cur, vp[None, :]])
        s = np.append(s, stdd)
        added += 1
    return v_cur, s, added


def vp_angles(v):
    """calc_angle_to_other_vp for every VP, vp_localisation.py:687-697."""
    c = np.clip(np.dot(v, v.T), -1, 1)
    ang = np.abs(np.arccos(np.clip(np.abs(c), -1, 1)))
    np.fill_diagonal(ang, PI)
    return ang


def merge_vps(v_other, v_at, s, l, thresh, lweight, lsim, colsum, wbias, pdfpar, lp, max_stdd=0.01):
    """vp_localisation.py:633-684.  v_at is the VP row-set the reference
    indexes with `i`; v_other is the adjacent history slice that is deleted
    in step.  Returns (v_other, v_at, s)."""
    while v_at.shape[0] > 1:
        ang = vp_angles(v_at)
        j, k = np.unravel_index(ang.argmin(), ang.shape)
        if not ang[j, k] < thresh:
            break
        p = calc_probabilities(pdfpar, v_at, lp, s)
        w = weight_matrix(p.vl, lweight, lsim, wbias, colsum)
        new_vp = calc_new_vanishing_point(l, w[j, :] + w[k, :])
        pv = p.vl[k, :] + p.vl[j, :]
        with np.errstate(divide="ignore", invalid="ignore", under="ignore"):
            s[k] = np.exp(np.log(np.sum(0.5 * (p.lvsq[:, j] + p.lvsq[:, k]) * pv)) - np.log(np.sum(pv)))
        if new_vp is None or s[k] > max_stdd:
            break
        v_at[k, :] = new_vp
        v_at = np.delete(v_at, j, axis=0)
        v_other = np.delete(v_other, j, axis=0)
        s = np.delete(s, j, axis=0)
    return v_other, v_at, s


# --------------------------------------------------------------------------
# the EM driver  (vp_localisation.py:168-450)
# --------------------------------------------------------------------------
def expectation_maximisation(l, lp, cnn_response, num_iter=100, sphere_image=None,
                             init_vp=None, do_merge=True, do_split=True, do_iterations=True,
                             distance_measure="angle", use_weights=True, wbias=1, num_init_vp=25,
                             split_merge_freq=10, merge_thresh=1e-3, outlier_thresh=1.96 ** 2,
                             final_convergence=5e-3, s_thresh=1e-200, num_min_lines=3,
                             clusterer=average_linkage_two_clusters, trace=None):
    """Same signature/defaults/result keys as the reference.  Mutates `l`
    (row normalisation, :186/:226) like the reference.  Only
    distance_measure="angle" (the only value any caller passes)."""
    assert distance_measure == "angle"
    N = l.shape[0]
    lsim = calc_lsim(lp, sigma=1) if use_weights else np.zeros((N, N))          # :177-180
    colsum = np.sum(lsim, axis=0)
    l /= np.sqrt(np.sum(l * l, axis=1))[:, None]                                 # :186
    max_stdd = 1e-6                                                              # :197
    result = {"vp_assoc": None, "vp": None, "counts": None, "count_id": None,
              "decision_metric": None, "iterations": 0}
    v0 = find_initial_vps(sphere_image, cnn_response, num_init_vp)              # :208
    pdfpar = pdf_params(cnn_response)                                            # :210
    if init_vp is not None:
        v0 = init_vp / np.sqrt(np.sum(init_vp * init_vp, axis=1))[:, None]
    langles = lines_angles(lp)
    s_init = pdfpar.sigma * 1e-6                                                 # :219
    l /= np.sqrt(np.sum(l * l, axis=1))[:, None]                                 # :226
    llen = seg_length(lp)
    if use_weights:
        lweight = llen * np.clip(line_rating_knn(lp, k2=4), 0.2, 1)              # :230-233
    else:
        lweight = np.ones(N)
    cur = v0.copy()
    s = np.ones(cur.shape[0]) * s_init

    def estep(v):
        return calc_probabilities(pdfpar, v, lp, s)

    def wmat(p):
        return weight_matrix(p.vl, lweight, lsim, wbias, colsum)

    p = estep(cur)
    w = wmat(p)
    counts, _, _ = calc_vp_line_counts(cur, lp, s, w, lweight, outlier_thresh)   # :247
    keep = ~(counts < 3)
    cur, s = cur[keep], s[keep]                                                  # :250-251
    nxt = np.zeros_like(cur)
    if trace is not None:
        trace.append(("init", cur.copy(), s.copy()))

    for i in range(num_iter):
        M = cur.shape[0]
        if M == 0:                                                               # :258
            return result
        if i % split_merge_freq == 0 and 0 < i < 100 and do_split:               # :262
            p = estep(cur)
            w = wmat(p)
            cur, s, added = split_best_vp(cur, s, lp, l, w, lweight, langles, merge_thresh, clusterer)
            if added:
                nxt = np.vstack([nxt, np.zeros((added, 3))])
        M = cur.shape[0]
        p = estep(cur)                                                           # :273
        max_err = 0.0
        rem = []
        w = wmat(p)                                                              # :282
        for m in range(M):
            if not do_iterations:
                break
            nv = calc_new_vanishing_point(l, w[m, :])                            # :292
            if nv is None:
                rem.append(m)
                continue
            nxt[m] = nv
            s[m] = _s_update(p.lvsq[:, m], p.vl[m, :])                           # :301-304
            s[m] = np.maximum(np.minimum(s[m], max_stdd), s_thresh)              # :306-307
            if np.isnan(s[m]):
                rem.append(m)
            else:
                err = np.arccos(np.minimum(np.abs(np.dot(cur[m], nxt[m])), 1.0))
                max_err = np.maximum(max_err, err)
                if err > 1.5:
                    rem.append(m)
        if not do_iterations:
            nxt = cur.copy()
        rem = np.array(rem, dtype=int)
        cur, nxt, s = np.delete(cur, rem, 0), np.delete(nxt, rem, 0), np.delete(s, rem, 0)
        p = estep(cur)                                                           # :332 (index i)
        if trace is not None:
            trace.append(("iter", i, nxt.copy(), s.copy(), float(max_err)))

        if max_err < final_convergence or i == num_iter - 1 or not do_iterations:   # :335
            if do_merge:
                cur, nxt, s = merge_vps(cur, nxt, s, l, merge_thresh * 10, lweight, lsim, colsum,
                                        wbias, pdfpar, lp)                       # :339
            if trace is not None:
                trace.append(("final_merged", cur.copy(), nxt.copy(), s.copy()))
            p = estep(cur)                                                       # :344 (index i, sic)
            w = wmat(p)
            rem = []
            assoc = np.argmax(w, axis=0)
            if trace is not None:
                trace.append(("final_assoc", np.bincount(assoc, minlength=cur.shape[0])))
            for m in range(cur.shape[0]):
                sel = assoc == m
                if not np.any(sel):
                    continue
                w[m, sel] /= np.max(w[m, sel])                                   # :358
                nv = calc_new_vanishing_point(l[sel, :], w[m, sel])
                if nv is None:
                    rem.append(m)
                    continue
                nxt[m] = nv
                s[m] = np.minimum(_s_update(p.lvsq[:, m], p.vl[m, :]), max_stdd)   # :374-377
                if np.isnan(s[m]) or s[m] < s_thresh:
                    rem.append(m)
                else:
                    err = np.arccos(np.minimum(np.abs(np.dot(cur[m], nxt[m])), 1.0))
                    if err > 1.5:
                        rem.append(m)
            rem = np.array(rem, dtype=int)
            if trace is not None:
                trace.append(("final_refit_removed", rem.copy(), nxt.copy(), s.copy()))
            cur, nxt, s = np.delete(cur, rem, 0), np.delete(nxt, rem, 0), np.delete(s, rem, 0)
            p = estep(cur)                                                       # :398
            dm = wmat(p)
            if dm.size <= 0:
                return result
            good = np.unique(np.argmax(dm, axis=0))                              # :406-408
            if trace is not None:
                trace.append(("final_good", good.copy(), np.bincount(np.argmax(dm, axis=0), minlength=cur.shape[0])))
            cur, nxt, s = cur[good], nxt[good], s[good]
            p = estep(nxt)                                                       # :415 (index i+1)
            dm = wmat(p)
            counts, cw, assoc = calc_vp_line_counts(nxt, lp, s, dm, lweight, outlier_thresh)
            vidx = 0
            while vidx < nxt.shape[0]:                                           # :423-437
                if counts[vidx] < num_min_lines:
                    cur, nxt, s = np.delete(cur, vidx, 0), np.delete(nxt, vidx, 0), np.delete(s, vidx)
                    p = estep(nxt)
                    dm = wmat(p)
                    counts, cw, assoc = calc_vp_line_counts(nxt, lp, s, dm, lweight, outlier_thresh)
                else:
                    vidx += 1
            return {"vp_assoc": assoc, "vp": nxt, "counts": counts, "counts_weighted": cw,
                    "count_id": None, "decision_metric": dm, "iterations": i, "distribution": p,
                    "sigma": s}

        if i % split_merge_freq == 0 and 0 < i <= 100 + split_merge_freq and do_merge:   # :444
            cur, nxt, s = merge_vps(cur, nxt, s, l, merge_thresh, lweight, lsim, colsum, wbias, pdfpar, lp)
        cur, nxt = nxt, np.zeros_like(nxt)
    return result
