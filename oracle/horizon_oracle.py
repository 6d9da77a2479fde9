"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's horizon /
orthogonal-triplet selection (calc_horizon.py:19-225), SURVEY.md section 8(f) row N1.

Pinned: tests/golden/horizon_cases.npz holds inputs and outputs of the reference's own
calc_horizon.calculate_horizon_and_ortho_vp (imported unmodified from /root/reference by
oracle/make_golden_horizon.py); tests/test_oracle_horizon.py checks this restatement
against them.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.

The triplets are scored as arrays instead of one Python loop iteration each; every
expression keeps the reference's operation order.
"""
import itertools

import numpy as np


def _in_image(v):
    """VPinImage (calc_horizon.py:11-16) for rows of v."""
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y = v[:, 0] / v[:, 2], v[:, 1] / v[:, 2]
    return (x <= 1) & (x >= -1) & (y <= 1) & (y >= -1)


def _horizon_points(hlin):
    """:219-222 (and :172-175)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        p1 = np.cross(hlin, np.array([1.0, 0.0, 1.0]))
        p2 = np.cross(hlin, np.array([-1.0, 0.0, 1.0]))
        return p1 / p1[..., 2:3], p2 / p2[..., 2:3]


def calculate_horizon_and_ortho_vp(em_result, maxbest=10, theta_vmin=np.pi / 10., theta_z=np.pi / 4.):
    vps = np.asarray(em_result["vp"], np.float64).reshape(-1, 3).copy()          # :22
    counts = np.asarray(em_result["counts"])                                      # :25
    M = vps.shape[0]
    nb = int(min(maxbest, M))                                                      # :28
    zen = np.abs(vps[:, 1]) > np.sin(theta_z)                                      # :31
    best = np.argsort(counts, kind="stable")[::-1][:nb]                            # :34-36 (see csrc/horizon.cu on ties)
    up = np.array([0.0, 1.0, 0.0])
    if nb > 2:
        tri = np.array(list(itertools.combinations(range(nb), 3)), dtype=int)      # :45-51: i < j < k, lexicographic
        ia, ib, ic = best[tri[:, 0]], best[tri[:, 1]], best[tri[:, 2]]
        Va, Vb, Vc = vps[ia], vps[ib], vps[ic]
        dot = lambda p, q: p[:, 0] * q[:, 0] + p[:, 1] * q[:, 1] + p[:, 2] * q[:, 2]
        AB, BC, AC = np.abs(dot(Va, Vb)), np.abs(dot(Vb, Vc)), np.abs(dot(Va, Vc))   # :77-83
        num_zenith = zen[ia].astype(int) + zen[ib] + zen[ic]                       # :86-95
        zenith = np.where(zen[ic][:, None], Vc, np.where(zen[ib][:, None], Vb, Va))   # the last candidate wins
        num_central = _in_image(Va).astype(int) + _in_image(Vb) + _in_image(Vc)    # :98-104
        ya, yb, yc = np.abs(Va[:, 1]), np.abs(Vb[:, 1]), np.abs(Vc[:, 1])
        za = (ya > yb) & (ya > yc)                                                 # :108-128
        zb = ~za & (yb > ya) & (yb > yc)
        sel = lambda a, b, c: np.where(za[:, None], a, np.where(zb[:, None], b, c))
        zVP, h1, h2 = sel(Va, Vb, Vc), sel(Vb, Va, Va), sel(Vc, Vc, Vb)
        cnt = counts.astype(np.float64)
        selc = lambda a, b, c: np.where(za, a, np.where(zb, b, c))
        h1c, h2c = selc(cnt[ib], cnt[ia], cnt[ia]), selc(cnt[ic], cnt[ic], cnt[ib])
        zid = selc(ia, ib, ic)
        with np.errstate(divide="ignore", invalid="ignore"):
            z0, z1 = zVP[:, 1] * 1.0 - zVP[:, 2] * 0.0, zVP[:, 2] * 0.0 - zVP[:, 0] * 1.0     # :131
            zn = np.sqrt(z0 * z0 + z1 * z1)                                        # :132
            l1, l2 = z0 / zn, z1 / zn
            p1, p2 = h1 / h1[:, 2:3], h2 / h2[:, 2:3]
            e1, e2 = np.array([0.0, 0.0, 1.0]) - p1, np.array([0.0, 0.0, 1.0]) - p2
            d1, d2 = np.sqrt(dot(e1, e1)), np.sqrt(dot(e2, e2))                    # :144-145
            h3 = ((h1[:, 0] * l2 - h1[:, 1] * l1) / h1[:, 2] * (d2 * h1c) + (h2[:, 0] * l2 - h2[:, 1] * l1) / h2[:, 2] * (d1 * h2c)) \
                / ((d1 * h2c) + (d2 * h1c))                                        # :149
            hlin = np.stack([-l2, l1, h3], axis=1)
            hvec = p1 - p2                                                         # :154
            hn = np.sqrt(dot(hvec, hvec))
            hang = np.arccos(np.abs(hvec[:, 0] * 1.0 + hvec[:, 1] * 0.0 + hvec[:, 2] * 0.0) / hn)   # :155
            hP1, hP2 = _horizon_points(hlin)
            zl = np.sqrt(dot(zenith, zenith))
            cosphi = np.abs(dot(hvec / hn[:, None], zenith / zl[:, None]))         # :167
            ortho = np.where(num_zenith == 1, 1 - np.clip(1.0 * cosphi, 0, 1), 0.0)   # :165-168
            zenith_pos = np.where(zVP[:, 1] > 0, 1, -1)                            # :170
            hor_pos = np.where((hP1[:, 1] + hP2[:, 1]) / 2 < 0, 1, -1)             # :171
            costh = np.cos(theta_vmin)                                             # :54
            ok = (AB < costh) & (BC < costh) & (AC < costh) & (num_zenith == 1) & (num_central <= 1) & \
                 (hang < 30 * np.pi / 180) & (zenith_pos * hor_pos == 1)           # :177-180
        weight = cnt[ia] + cnt[ib] + cnt[ic]                                       # :183
        score = ok.astype(np.float64) * weight * ortho                             # :186
        # first triplet whose score exceeds everything before it, starting from -1 (:57, :191-197)
        cand = np.where(score > -1)[0]
        win = 0 if cand.size == 0 else int(cand[np.argmax(score[cand])])
        hVP1, hVP2, z, hl = h1[win], h2[win], zVP[win], hlin[win]
        combo = best[tri[win]]                                                     # :199
        solutions = [{"score": score[i], "zVP_id": int(zid[i]), "horizon": hlin[i]} for i in np.where(score > 0)[0]]
    else:
        solutions = []
        z = up
        if nb > 1:                                                                 # :200-205
            hVP1, hVP2, combo = vps[0], vps[1], np.array([0, 1])
            hl = np.cross(hVP1, hVP2)
        else:
            hVP1 = vps[0] if nb > 0 else np.array([-1.0, 0.0, 0.0])                # :206-217
            hVP2 = vps[0] if nb > 0 else np.array([1.0, 0.0, 0.0])
            combo = np.array([0, 0])
            hl = np.cross(np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 1.0]))
    hP1, hP2 = _horizon_points(np.asarray(hl, np.float64))
    return hP1, hP2, np.asarray(z, np.float64), np.asarray(hVP1, np.float64), np.asarray(hVP2, np.float64), np.asarray(combo)


def horizon_error(hP1, hP2, true_horizon, scale, image_height):
    """benchmark.py:247-253: distance of the estimated horizon from the ground truth at the image borders."""
    with np.errstate(divide="ignore", invalid="ignore"):
        thP1, thP2 = _horizon_points(np.asarray(true_horizon, np.float64))
    return np.maximum(np.abs(hP1[1] - thP1[1]), np.abs(hP2[1] - thP2[1])) / 2 * scale * 1.0 / image_height
