"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

CPU restatement (NumPy, float64, whole-array ops exactly like the reference executes them) of SPOMSO's algorithm
for the SDF hot path, expressed over the same flattened op list the CUDA interpreter runs. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the package
aegolius_b200 never does (tests/test_no_oracle_in_product.py enforces it).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself).
This oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
tests/golden/make_golden.py (imports /root/reference/Code/spomso, runs obj.create(co)) and committed as
tests/golden/*.npz; tests/test_oracle_golden.py checks oracle == reference to <= 1e-12 * extent on every fixture.

Every function cites the reference lines it restates (paths relative to Code/spomso/spomso/cores/).
"""
from __future__ import annotations

import numpy as np

# opcode numbers: kept literal here so the oracle does not import the product package
(END, SAVE_P, LOAD_P, PUSH_V) = (0, 1, 2, 3)
(AFFINE, TRANSLATE, SCALE_P, ELONGATE, TWIST, BEND, ABSX_SUB, SYMMETRY, ROTSYM, REVOLVE, AXIS_REVOLVE, REP_INF,
 REP_FIN, LIN_INST, CURVE_INST, ZERO_Z) = range(8, 24)
(NEXT_AFFINE, NEXT_TRANSLATE, NEXT_LOAD) = (24, 25, 26)  # fused PUSH_V + LOAD_P + transform (program.py peephole)
(ROUND, ABS, NEG, SIGN, ONION, CONCENTRIC, SCALE_V, EXTRUDE_BEGIN, EXTRUDE_END, POLY_SIGN) = range(32, 42)
(PP_SIGMOID, PP_POS_SIGMOID, PP_CAPPED_EXP, PP_HARD_BIN, PP_LINEAR, PP_RELU, PP_SMOOTH_RELU, PP_SLOWSTART,
 PP_GAUSS_BOUNDARY, PP_GAUSS_FALLOFF) = range(48, 58)
(C_UNION, C_INTERSECT, C_SUBTRACT, C_SUM, C_DIFF, C_SMIN2, C_SMIN3, C_SMAX3, C_SSUB3, C_BOLTZ_INT,
 C_BOLTZ_SUB) = range(64, 75)
(P_SPHERE, P_CYLINDER, P_BOX, P_TORUS, P_CHAINLINK, P_BRAID, P_ARC3D, P_PLANE, P_UPLANE, P_SEGMENT, P_CONE,
 P_OINF_CONE, P_INF_CONE, P_SOLID_ANGLE, P_TRIANGLE3D, P_QUAD3D, P_SEGLINE, P_AXIS, P_POINT_CLOUD) = range(96, 115)
P_FIELD = 115  # output of a grid-stencil stage (conv_averaging / conv_edge_detection as modifications)
(P_CIRCLE, P_NEU_CIRCLE, P_BOX2D, P_SEGMENT2D, P_RBOX2D, P_TRIANGLE2D, P_ARC, P_SECTOR, P_INF_SECTOR, P_NGON,
 P_SEGLINE2D, P_POLYGON2D) = range(128, 140)


def _norm(*c):
    return np.sqrt(sum(x * x for x in c))


# ---- combine.py:12-34 ------------------------------------------------------------------------------------------
def _smin2(x, y, a):  # combine.py:12-18
    h = np.maximum(a - np.abs(x - y), 0.0) / a
    return np.minimum(x, y) - h * h * a / 4.0


def _smin3(x, y, a):  # combine.py:20-26
    h = np.maximum(a - np.abs(x - y), 0.0) / a
    return np.minimum(x, y) - h * h * h * a / 6.0


def _boltz(x, y, a):  # combine.py:29-34
    with np.errstate(over="ignore", invalid="ignore"):
        e1 = np.exp(x / a)
        e2 = np.exp(y / a)
        return (x * e1 + y * e2) / (e1 + e2)


# ---- sdf_3D.py / sdf_2D.py helpers -----------------------------------------------------------------------------
def _segment(pa, ba, bb):  # sdf_3D.py:111-118, sdf_2D.py:31-38
    h = np.clip(sum(p * b for p, b in zip(pa, ba)) / bb, 0, 1)
    return _norm(*[p - b * h for p, b in zip(pa, ba)])


def _arc_core(x, y, C, S, R, ea):  # sdf_2D.py:85-102 / sdf_3D.py:78-96
    xr = C * x + S * y
    yr = np.abs(-S * x + C * y)
    phi = np.arctan2(yr, xr)
    psi = np.clip(phi, 0, ea)
    return xr - R * np.cos(psi), yr - R * np.sin(psi)


def _sector_core(x, y, radius, ad, cad, sad):  # sdf_2D.py:105-129 after rotation & fold; sdf_3D.py:160-183
    phi = np.arctan2(y, x)
    psi = np.clip(phi, 0, ad)
    length = _norm(x - radius * np.cos(psi), y - radius * np.sin(psi))
    t = np.clip(x * cad + y * sad, 0, radius)
    m = _norm(x - cad * t, y - sad * t)
    msk = (_norm(x, y) <= radius) * (phi <= ad)
    out = np.minimum(m, length)
    return np.where(msk, -out, out)


def _tri_edge_sq(co_v, s, ss):  # one edge term of sdf_3D.py:203-206
    h = np.clip(sum(a * b for a, b in zip(s, co_v)) / ss, 0, 1)
    t = [a * h - b for a, b in zip(s, co_v)]
    return sum(x * x for x in t)


def run(prog, co, return_state=False, return_margin=False, fields=None):
    """Evaluates the program on fp64 coordinates co (3,N); returns the field (N,) float64.
    Follows apply_ec_transforms (transformations.py:232-242) + the closure chain semantics (see program.py).

    With return_margin=True also returns, per point, the distance (in the local units of the op) to the nearest
    DISCONTINUOUS branch boundary taken on the way (nearest-instance ties, repetition cell edges, sector edges,
    sign / binarisation thresholds ...). Where that margin is below the comparison band the reference's own choice
    is decided by rounding (or, for KD-tree ties, by scipy's traversal order), so parity tests exclude those points
    (SURVEY §7.3 item 2/7)."""
    co = np.asarray(co, dtype=np.float64)
    x, y, z = co[0].copy(), co[1].copy(), co[2].copy()
    acc = np.zeros(co.shape[1])
    margin = np.full(co.shape[1], np.inf)
    P, V = {}, {}
    A = prog.args
    with np.errstate(divide="ignore", invalid="ignore"):
        for op in prog.ops:
            code, a, b, o = int(op["opcode"]), int(op["a"]), int(op["b"]), int(op["arg"])
            if code == END:
                break
            elif code == SAVE_P:
                P[a] = (x, y, z)
            elif code == LOAD_P:
                x, y, z = P[a]
            elif code == PUSH_V:
                V[a] = acc
            elif code in (NEXT_AFFINE, NEXT_TRANSLATE, NEXT_LOAD):
                if b:
                    V[b - 1] = acc
                x, y, z = P[a]
                if code == NEXT_AFFINE:
                    m = A[o:o + 12]
                    x, y, z = (m[0] * x + m[1] * y + m[2] * z + m[9], m[3] * x + m[4] * y + m[5] * z + m[10],
                               m[6] * x + m[7] * y + m[8] * z + m[11])
                elif code == NEXT_TRANSLATE:
                    x, y, z = x + A[o], y + A[o + 1], z + A[o + 2]
            # ---- coordinate ops ----
            elif code == AFFINE:  # transformations.py:238-240 folded: M = R^T/s, b = -R^T t
                m = A[o:o + 12]
                x, y, z = (m[0] * x + m[1] * y + m[2] * z + m[9], m[3] * x + m[4] * y + m[5] * z + m[10],
                           m[6] * x + m[7] * y + m[8] * z + m[11])
            elif code == TRANSLATE:
                x, y, z = x + A[o], y + A[o + 1], z + A[o + 2]
            elif code == SCALE_P:
                x, y, z = x * A[o], y * A[o], z * A[o]
            elif code == ELONGATE:  # modifications.py:91-93
                x = x - np.minimum(np.maximum(x, A[o]), A[o + 3])
                y = y - np.minimum(np.maximum(y, A[o + 1]), A[o + 4])
                z = z - np.minimum(np.maximum(z, A[o + 2]), A[o + 5])
            elif code == TWIST:  # modifications.py:516-522
                c, s = np.cos(A[o] * z), np.sin(A[o] * z)
                x, y = c * x - s * y, s * x + c * y
            elif code == BEND:  # modifications.py:545-573
                r, ha, c, s, thr, rs, r1c, rha = A[o:o + 8]
                qy = y - r
                phi = np.arctan2(x, -qy)
                margin = np.minimum(margin, np.where(qy > 0, np.abs(x), np.inf))
                ny_ = -r + _norm(x, qy)
                nx_ = r * phi
                mask1 = thr <= np.abs(nx_)
                sg = np.sign(x)
                w0 = x - rs * sg
                w1 = y - r1c
                pos = x >= 0
                wr0 = np.where(pos, c * w0 + s * w1, c * w0 - s * w1)
                wr1 = np.where(pos, -s * w0 + c * w1, s * w0 + c * w1)
                wr0 = wr0 + rha * sg
                x, y = np.where(mask1, wr0, nx_), np.where(mask1, wr1, ny_)
            elif code == ABSX_SUB:  # modifications.py:991-993
                x = np.abs(x) - A[o]
            elif code == SYMMETRY:  # modifications.py:951
                if a == 0:
                    x = np.abs(x)
                elif a == 1:
                    y = np.abs(y)
                else:
                    z = np.abs(z)
            elif code == ROTSYM:  # modifications.py:1023-1029
                ang, rad = A[o], A[o + 1]
                phi = np.arctan2(y, x)
                phi = np.where(phi < 0, 2 * np.pi + phi, phi)
                u = np.mod(phi, ang)
                rr = _norm(x, y)
                margin = np.minimum(margin, rr * np.minimum(u, ang - u))
                phi = u - ang / 2
                x, y = rr * np.cos(phi) - rad, rr * np.sin(phi)
            elif code == REVOLVE:  # modifications.py:427-431
                x, y, z = _norm(x, z) - A[o], y, np.zeros_like(z)
            elif code == AXIS_REVOLVE:  # modifications.py:455-467
                rad, c, s = A[o:o + 3]
                xr, yr = c * x + s * y, -s * x + c * y
                m = _norm(xr, z)
                x, y, z = c * m - s * yr - rad, s * m + c * yr, np.zeros_like(z)
            elif code == REP_INF:  # modifications.py:819-820
                d, h = A[o:o + 3], A[o + 3:o + 6]
                ux, uy, uz = np.mod(x + h[0], d[0]), np.mod(y + h[1], d[1]), np.mod(z + h[2], d[2])
                for u_, d_ in ((ux, d[0]), (uy, d[1]), (uz, d[2])):
                    margin = np.minimum(margin, np.minimum(np.abs(u_), np.abs(d_ - u_)))  # (np.mod takes the sign of d)
                x, y, z = ux - h[0], uy - h[1], uz - h[2]
            elif code == REP_FIN:  # modifications.py:847-868
                c, d, s, sh = A[o:o + 3], A[o + 3:o + 6], A[o + 6:o + 9], A[o + 9:o + 12]
                new = []
                for i, q in enumerate((x, y, z)):
                    inner = (q >= -d[i]) * (q <= d[i])
                    v = np.abs(q) - c[i]
                    v = v - 2 * v * (q < 0)
                    um = np.mod(q - d[i], s[i])
                    u = um - sh[i]
                    margin = np.minimum(margin, np.where(inner, np.minimum(um, s[i] - um), np.inf))
                    margin = np.minimum(margin, np.abs(np.abs(q) - d[i]))
                    new.append(np.where(inner, u, v))
                x, y, z = new
            elif code == LIN_INST:  # modifications.py:1069-1083
                lh, s, d, lo, hi, off = A[o:o + 6]
                v = np.abs(x) - lh
                v = v - 2 * v * (x < 0)
                if a:
                    um = np.mod(x - off, s)
                    inner = (x >= lo) * (x <= hi)
                    margin = np.minimum(margin, np.where(inner, np.minimum(um, s - um), np.inf))
                    margin = np.minimum(margin, np.minimum(np.abs(x - lo), np.abs(x - hi)))
                    v = np.where(inner, um - d, v)
                else:
                    margin = np.minimum(margin, np.abs(x))
                x = v
            elif code == CURVE_INST:  # modifications.py:1120-1127, 1183-1191, 1253-1261 (nearest instance)
                n = int(A[o])
                stride = 12 if a == 1 else 3
                rec = A[o + 4:o + 4 + n * stride].reshape(n, stride)
                best = np.full(x.shape, np.inf)
                second = np.full(x.shape, np.inf)
                idx = np.zeros(x.shape, dtype=np.int64)
                for i in range(n):
                    d2 = (x - rec[i, 0]) ** 2 + (y - rec[i, 1]) ** 2 + (z - rec[i, 2]) ** 2
                    upd = d2 < best
                    second = np.where(upd, best, np.minimum(second, d2))
                    best = np.where(upd, d2, best)
                    idx = np.where(upd, i, idx)
                margin = np.minimum(margin, np.sqrt(second) - np.sqrt(best))
                r = rec[idx]
                vx, vy, vz = x - r[:, 0], y - r[:, 1], z - r[:, 2]
                if a == 1:
                    x = r[:, 3] * vx + r[:, 4] * vy + r[:, 5] * vz
                    y = r[:, 6] * vx + r[:, 7] * vy + r[:, 8] * vz
                    z = r[:, 9] * vx + r[:, 10] * vy + r[:, 11] * vz
                else:
                    x, y, z = vx, vy, vz
            elif code == ZERO_Z:
                z = np.zeros_like(z)
            # ---- value ops ----
            elif code == ROUND:  # modifications.py:113
                acc = acc - A[o]
            elif code == ABS:  # :159
                acc = np.abs(acc)
            elif code == NEG:  # :295
                acc = -acc
            elif code == SIGN:  # :318
                margin = np.minimum(margin, np.abs(acc))
                acc = np.sign(acc)
            elif code == ONION:  # :384
                acc = np.abs(acc) - A[o]
            elif code == CONCENTRIC:  # :406
                acc = np.abs(acc - A[o])
            elif code == SCALE_V:  # transformations.py:242
                acc = acc * A[o]
            elif code == EXTRUDE_BEGIN:  # modifications.py:489-492
                V[a] = np.abs(z) - A[o]
                z = np.zeros_like(z)
            elif code == EXTRUDE_END:  # modifications.py:494-497
                w1 = V[a]
                acc = np.minimum(np.maximum(acc, w1), 0) + _norm(np.maximum(acc, 0), np.maximum(w1, 0))
            # ---- post-processing (post_processing.py:380-560) ----
            elif code == PP_SIGMOID:
                acc = A[o] * (1 / (1 + np.exp(4 * acc / A[o + 1])))
            elif code == PP_POS_SIGMOID:
                acc = A[o] * (1 / (1 + np.exp(4 * (acc - A[o + 1]) / A[o + 1])))
            elif code == PP_CAPPED_EXP:
                acc = A[o] * np.minimum(np.exp(-4 * acc / A[o + 1]), 1)
            elif code == PP_HARD_BIN:
                margin = np.minimum(margin, np.abs(acc - A[o]))
                acc = (acc <= A[o]).astype(np.float64)
            elif code == PP_LINEAR:
                acc = np.clip(1 - acc / A[o + 1], 0, 1) * A[o]
            elif code == PP_RELU:
                acc = np.maximum(acc / A[o], 0)
            elif code == PP_SMOOTH_RELU:
                v = acc / A[o + 1]
                acc = (v + np.sqrt(v ** 2 + A[o])) / 2
            elif code == PP_SLOWSTART:
                acc = np.sqrt(np.maximum(acc / A[o], 0) ** 2 + A[o + 1]) - A[o + 2]
            elif code == PP_GAUSS_BOUNDARY:
                acc = A[o] * np.exp(-4 * (acc / A[o + 1]) ** 2)
            elif code == PP_GAUSS_FALLOFF:
                acc = A[o] * np.exp(-4 * (np.maximum(acc, 0) / A[o + 1]) ** 2)
            # ---- combine (combine.py:51-78): left operand V[a], right operand acc ----
            elif code == C_UNION:
                acc = np.minimum(V[a], acc)
            elif code == C_INTERSECT:
                acc = np.maximum(V[a], acc)
            elif code == C_SUBTRACT:
                acc = np.maximum(V[a], -acc)
            elif code == C_SUM:
                acc = V[a] + acc
            elif code == C_DIFF:
                acc = V[a] - acc
            elif code == C_SMIN2:
                acc = _smin2(V[a], acc, A[o])
            elif code == C_SMIN3:
                acc = _smin3(V[a], acc, A[o])
            elif code == C_SMAX3:
                acc = -_smin3(-V[a], -acc, A[o])
            elif code == C_SSUB3:
                acc = -_smin3(-V[a], acc, A[o])
            elif code == C_BOLTZ_INT:
                acc = _boltz(V[a], acc, A[o])
            elif code == C_BOLTZ_SUB:
                acc = _boltz(V[a], -acc, A[o])
            # ---- 3D primitives (sdf_3D.py) ----
            elif code == P_SPHERE:  # :25-27
                acc = _norm(x, y, z) - A[o]
            elif code == P_CYLINDER:  # :30-37
                d0 = _norm(x, y) - A[o]
                d1 = np.abs(z) - A[o + 1]
                acc = np.minimum(np.maximum(d0, d1), 0) + _norm(np.maximum(d0, 0), np.maximum(d1, 0))
            elif code == P_BOX:  # :40-47
                q0, q1, q2 = np.abs(x) - A[o], np.abs(y) - A[o + 1], np.abs(z) - A[o + 2]
                acc = _norm(np.maximum(q0, 0.0), np.maximum(q1, 0.0), np.maximum(q2, 0.0)) + \
                    np.minimum(np.maximum(q0, np.maximum(q1, q2)), 0.0)
            elif code == P_TORUS:  # :50-53
                acc = _norm(_norm(x, y) - A[o], z) - A[o + 1]
            elif code == P_CHAINLINK:  # :56-61
                xx = x - np.clip(x, -A[o + 2], A[o + 2])
                acc = _norm(_norm(xx, y) - A[o], z) - A[o + 1]
            elif code == P_BRAID:  # :64-75
                hl, R, r, pitch = A[o:o + 4]
                c, s = np.cos(pitch * z), np.sin(pitch * z)
                xr, yr = c * x - s * y, s * x + c * y
                zz = z - np.clip(z, -hl, hl)
                acc = _norm(_norm(xr, zz) - R, yr) - r
            elif code == P_ARC3D:  # :78-96
                R, r, C, S, ea = A[o:o + 5]
                dx, dy = _arc_core(x, y, C, S, R, ea)
                acc = _norm(dx, dy, z) - r
            elif code == P_PLANE:  # :99-102
                acc = (x * A[o] + y * A[o + 1] + z * A[o + 2]) - A[o + 3]
            elif code == P_UPLANE:  # :105-108
                acc = np.abs(x * A[o] + y * A[o + 1] + z * A[o + 2]) - A[o + 3]
            elif code == P_SEGMENT:  # :111-118
                acc = _segment((x - A[o], y - A[o + 1], z - A[o + 2]), A[o + 3:o + 6], A[o + 6])
            elif code == P_CONE:  # :121-136
                q0, q1, zoff, qq = A[o:o + 4]
                w0, w1 = _norm(x, y), z - zoff
                t = np.clip((w0 * q0 + w1 * q1) / qq, 0.0, 1.0)
                a0, a1 = w0 - q0 * t, w1 - q1 * t
                b0, b1 = w0 - q0 * np.clip(w0 / q0, 0.0, 1.0), w1 - q1
                d = np.minimum(a0 * a0 + a1 * a1, b0 * b0 + b1 * b1)
                s = np.maximum(-(w0 * q1 - w1 * q0), -(w1 - q1))
                acc = np.sqrt(d) * np.sign(s)
            elif code in (P_OINF_CONE, P_INF_CONE):  # :139-157
                v0, v1 = A[o], A[o + 1]
                q0, q1 = _norm(x, y), -z
                t = np.maximum(q0 * v0 + q1 * v1, 0.0)
                d = _norm(q0 - v0 * t, q1 - v1 * t)
                if code == P_OINF_CONE:
                    d = d * (-2 * (q0 * v1 - q1 * v0 < 0.0) + 1)
                acc = d
            elif code == P_SOLID_ANGLE:  # :160-183
                radius, C, S, ad, cad, sad = A[o:o + 6]
                xr = C * x + S * y
                yr = _norm(-S * x + C * y, z)
                acc = _sector_core(xr, yr, radius, ad, cad, sad)
            elif code == P_TRIANGLE3D:  # :186-214
                g = A[o:o + 34]
                va, vb, vc, s1, s2, s3, nrm, c1, c2, c3 = (g[3 * i:3 * i + 3] for i in range(10))
                coa = (x - va[0], y - va[1], z - va[2])
                cob = (x - vb[0], y - vb[1], z - vb[2])
                coc = (x - vc[0], y - vc[1], z - vc[2])
                dot = lambda u, v: u[0] * v[0] + u[1] * v[1] + u[2] * v[2]
                msum = np.sign(dot(c1, coa)) + np.sign(dot(c2, cob)) + np.sign(dot(c3, coc))
                ex1 = np.minimum(np.minimum(_tri_edge_sq(coa, s1, g[30]), _tri_edge_sq(cob, s2, g[31])),
                                 _tri_edge_sq(coc, s3, g[32]))
                ex2 = dot(nrm, coa) ** 2 / g[33]
                acc = np.sqrt(np.where(msum < 2., ex1, ex2))
            elif code == P_QUAD3D:  # :217-250
                g = A[o:o + 44]
                va, vb, vc, vd, s1, s2, s3, s4, nrm, c1, c2, c3, c4 = (g[3 * i:3 * i + 3] for i in range(13))
                cos_ = [(x - v[0], y - v[1], z - v[2]) for v in (va, vb, vc, vd)]
                dot = lambda u, v: u[0] * v[0] + u[1] * v[1] + u[2] * v[2]
                msum = sum(np.sign(dot(c, cv)) for c, cv in zip((c1, c2, c3, c4), cos_))
                e = [_tri_edge_sq(cv, s, ss) for cv, s, ss in zip(cos_, (s1, s2, s3, s4), g[39:43])]
                ex1 = np.minimum(np.minimum(e[3], e[2]), np.minimum(e[0], e[1]))
                ex2 = dot(nrm, cos_[0]) ** 2 / g[43]
                acc = np.sqrt(np.where(msum < 3., ex1, ex2))
            elif code == P_SEGLINE:  # :264-271
                n = int(A[o])
                pts = A[o + 1:o + 1 + 3 * n].reshape(n, 3)
                out = np.ones(x.shape) * 1e16
                for i in range(n - 1):
                    ba = pts[i + 1] - pts[i]
                    out = np.minimum(out, _segment((x - pts[i, 0], y - pts[i, 1], z - pts[i, 2]), ba, np.dot(ba, ba)))
                acc = out
            elif code == P_AXIS:  # :13-22
                acc = (x, y, z)[a] - A[o]
            elif code == P_FIELD:  # modifications.py:1586-1637: the filtered field of an earlier stage, same point order
                acc = np.asarray(fields[b], dtype=np.float64)
                if acc.shape != x.shape:
                    raise ValueError("P_FIELD: the bound field does not match the evaluation points")
            elif code == P_POINT_CLOUD:  # sdf_3D.py:283-286 / sdf_2D.py:221-224 (exact NN distance)
                from scipy.spatial import cKDTree
                cloud = prog.blobs[b]
                if a == 3:
                    acc = cKDTree(cloud[:3].T).query(np.stack([x, y, z], axis=1))[0]
                else:
                    acc = cKDTree(cloud[:2].T).query(np.stack([x, y], axis=1))[0]
            # ---- 2D primitives (sdf_2D.py) ----
            elif code == P_CIRCLE:  # :12-14
                acc = _norm(x, y) - A[o]
            elif code == P_NEU_CIRCLE:  # :17-19
                p = A[o + 1]
                if np.isinf(p):
                    acc = np.maximum(np.abs(x), np.abs(y)) - A[o]
                else:
                    acc = (np.abs(x) ** p + np.abs(y) ** p) ** (1.0 / p) - A[o]
            elif code == P_BOX2D:  # :22-28
                d0, d1 = np.abs(x) - A[o], np.abs(y) - A[o + 1]
                acc = _norm(np.maximum(d0, 0), np.maximum(d1, 0)) + np.minimum(np.maximum(d0, d1), 0)
            elif code == P_SEGMENT2D:  # :31-38
                acc = _segment((x - A[o], y - A[o + 1]), A[o + 2:o + 4], A[o + 4])
            elif code == P_RBOX2D:  # :41-57
                hx, hy, r0, r1, r2, r3 = A[o:o + 6]
                r = np.where(y > 0, np.where(x < 0, r3, r2), np.where(x > 0, r1, r0))
                if len({r0, r1, r2, r3}) > 1:
                    margin = np.minimum(margin, np.minimum(np.abs(x), np.abs(y)))
                d0, d1 = np.abs(x) - hx + r, np.abs(y) - hy + r
                acc = _norm(np.maximum(d0, 0), np.maximum(d1, 0)) + (np.minimum(np.maximum(d0, d1), 0) - r)
            elif code == P_TRIANGLE2D:  # :60-82
                g = A[o:o + 16]
                pv = [g[0:2], g[2:4], g[4:6]]
                ev = [g[6:8], g[8:10], g[10:12]]
                s = g[15]
                d0 = d1 = None
                for i in range(3):
                    v0, v1 = x - pv[i][0], y - pv[i][1]
                    h = np.clip((v0 * ev[i][0] + v1 * ev[i][1]) / g[12 + i], 0, 1)
                    pq0, pq1 = v0 - ev[i][0] * h, v1 - ev[i][1] * h
                    dd = pq0 * pq0 + pq1 * pq1
                    cr = s * (v0 * ev[i][1] - v1 * ev[i][0])
                    d0 = dd if d0 is None else np.minimum(d0, dd)
                    d1 = cr if d1 is None else np.minimum(d1, cr)
                acc = -np.sqrt(d0) * np.sign(d1)
            elif code == P_ARC:  # :85-102
                R, C, S, ea = A[o:o + 4]
                dx, dy = _arc_core(x, y, C, S, R, ea)
                acc = _norm(dx, dy)
            elif code == P_SECTOR:  # :105-129
                radius, C, S, ad, cad, sad = A[o:o + 6]
                acc = _sector_core(C * x + S * y, np.abs(-S * x + C * y), radius, ad, cad, sad)
            elif code == P_INF_SECTOR:  # :132-150
                C, S, ad, cad, sad = A[o:o + 5]
                xr, yr = C * x + S * y, np.abs(-S * x + C * y)
                phi = np.arctan2(yr, xr)
                t = np.clip(xr * cad + yr * sad, 0, np.inf)
                acc = np.sign(phi - ad) * _norm(xr - cad * t, yr - sad * t)
            elif code == P_NGON:  # :153-177
                radius, alpha, tx, ty, nox, noy, l = A[o:o + 7]
                phi = np.arctan2(y, x)
                phi = np.where(phi < 0, 2 * np.pi + phi, phi)
                phi = np.mod(phi, alpha)
                rr = _norm(x, y)
                q0, q1 = np.cos(phi) * rr - radius, np.sin(phi) * rr
                h = np.clip(q0 * tx + q1 * ty, 0, l)
                acc = _norm(q0 - tx * h, q1 - ty * h) * np.sign(q0 * nox + q1 * noy)
            elif code == P_SEGLINE2D:  # :191-198
                n = int(A[o])
                pts = A[o + 1:o + 1 + 2 * n].reshape(n, 2)
                out = np.ones(x.shape) * 1e16
                for i in range(n - 1):
                    ba = pts[i + 1] - pts[i]
                    out = np.minimum(out, _segment((x - pts[i, 0], y - pts[i, 1]), ba, np.dot(ba, ba)))
                acc = out
            elif code == POLY_SIGN:  # d * interior at the coordinates saved in P[a]
                sx, sy, _ = P[a]
                n = int(A[o])
                if b == 0:  # .polygon(): geom_2d.py:530-555 / 601-626, interior_polygon == crossing number for simple polygons
                    pts = A[o + 1:o + 1 + 2 * n].reshape(n, 2)
                    inside = np.zeros(sx.shape, dtype=bool)
                    for i in range(n):
                        pa, pb = pts[i], pts[(i + 1) % n]
                        cond = (pa[1] > sy) != (pb[1] > sy)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            xi = (pb[0] - pa[0]) * (sy - pa[1]) / (pb[1] - pa[1]) + pa[0]
                        inside ^= cond & (sx < xi)
                    acc = acc * np.where(inside, -1.0, 1.0)
                else:  # .shape(): geom_2d.py:440-452 term by term
                    rec = A[o + 1:o + 1 + 6 * n].reshape(n, 6)
                    interior = np.ones(sx.shape)
                    for px, py, lx, ux, nx, ny in rec:
                        mask = (sx >= lx) * (sx < ux)
                        sgn = np.sign((sx - px) * nx + (sy - py) * ny)
                        interior = np.where(mask, interior * sgn, interior)
                        # a sample on an interval end switches segments: decided by rounding in fp32
                        margin = np.minimum(margin, np.minimum(np.abs(sx - lx), np.abs(sx - ux)))
                    acc = acc * interior
            elif code == P_POLYGON2D:  # sdf_2D.py:201-218; interior of a SIMPLE polygon (triangulation_functions.py:390-430
                # builds it as the union of the ear-clipping triangles, boundary included) == crossing-number rule
                n = int(A[o])
                pts = A[o + 1:o + 1 + 2 * n].reshape(n, 2)
                out = np.ones(x.shape) * 1e16
                inside = np.zeros(x.shape, dtype=bool)
                for i in range(n):
                    pa, pb = pts[i], pts[(i + 1) % n]
                    ba = pb - pa
                    out = np.minimum(out, _segment((x - pa[0], y - pa[1]), ba, np.dot(ba, ba)))
                    cond = (pa[1] > y) != (pb[1] > y)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        xi = (pb[0] - pa[0]) * (y - pa[1]) / (pb[1] - pa[1]) + pa[0]
                    inside ^= cond & (x < xi)
                    margin = np.minimum(margin, np.where(cond, np.abs(x - xi), np.inf))
                acc = np.where(inside, -out, out)
            else:
                raise ValueError(f"oracle: unknown opcode {code}")
            if C_UNION <= code <= C_BOLTZ_SUB and b:
                V[b - 1] = acc  # fused PUSH_V of a left-deep combine chain
    if return_state:
        return acc, (x, y, z)
    if return_margin:
        return acc, margin
    return acc


def run_grid(prog, size, res, x0=None, x1=None, chunk_planes=None, return_margin=False):
    """Evaluates on the generate_grid grid (helper_functions.py:23-93), plane range [x0, x1), chunked in x-slabs so
    that 513^3 / 1025^3 fit in host RAM (pointwise ops make slabbing exactly equivalent)."""
    dims = 3 if res[2] > 1 or size[2] != 0 else 2
    axes = [np.linspace(-size[i] / 2, size[i] / 2, res[i]) for i in range(dims)]
    if getattr(prog, "stages", None):
        return _run_grid_staged(prog, size, res, dims, axes, x0, x1, return_margin)
    x0 = 0 if x0 is None else x0
    x1 = res[0] if x1 is None else x1
    per_plane = res[1] * res[2]
    if chunk_planes is None:
        chunk_planes = max(1, (1 << 21) // per_plane)
    out = np.empty((x1 - x0) * per_plane)
    mar = np.empty((x1 - x0) * per_plane) if return_margin else None
    for s in range(x0, x1, chunk_planes):
        e = min(x1, s + chunk_planes)
        if dims == 3:
            co = np.asarray(np.meshgrid(axes[0][s:e], axes[1], axes[2], indexing="ij")).reshape(3, -1)
        else:
            c2 = np.asarray(np.meshgrid(axes[0][s:e], axes[1], indexing="ij")).reshape(2, -1)
            co = np.zeros((3, c2.shape[1]))
            co[:2] = c2
        sl = slice((s - x0) * per_plane, (e - x0) * per_plane)
        if return_margin:
            out[sl], mar[sl] = run(prog, co, return_margin=True)
        else:
            out[sl] = run(prog, co)
    return (out, mar) if return_margin else out


def _run_grid_staged(prog, size, res, dims, axes, x0, x1, return_margin):
    """Grid-stencil modifications (modifications.py:1586-1637: u = inner(co); u = smarter_reshape(u, res); stencil; flatten)
    restated over the op list: every stage filters the field accumulated by the ops before its P_FIELD op."""
    from . import fields_np
    if x0 not in (None, 0) or x1 not in (None, res[0]):
        raise ValueError("grid stencils need the whole grid")
    if dims == 3:
        co = np.asarray(np.meshgrid(axes[0], axes[1], axes[2], indexing="ij")).reshape(3, -1)
    else:
        c2 = np.asarray(np.meshgrid(axes[0], axes[1], indexing="ij")).reshape(2, -1)
        co = np.zeros((3, c2.shape[1]))
        co[:2] = c2
    shape = tuple(int(r) for r in res[:dims])
    fields = {}
    margin = np.full(co.shape[1], np.inf)
    for st in sorted(prog.stages, key=prog.stage_op_index):
        pre = prog.prefix(prog.stage_op_index(st))
        u, m = run(pre, co, return_margin=True, fields=fields)
        margin = np.minimum(margin, m)
        u = u.reshape(shape)
        if st["kind"] == 0:
            u = fields_np.conv_averaging(u, tuple(st["ksize"][:dims]), st["iterations"])
        elif st["kind"] == 2:  # signed (modifications.py:220-275); threshold = smallest grid step (:239-240)
            thr = min(abs(a[1] - a[0]) for a in axes[:dims])
            # a sample whose value sits on the threshold decides a whole row's crossing parity: reported as its margin
            margin = np.minimum(margin, np.abs(u.reshape(-1) - thr))
            u = fields_np.signed(u.reshape(-1), shape, thr)
        else:
            u = fields_np.conv_edge_detection(u)
        fields[st["blob"]] = u.reshape(-1)
    out, m = run(prog, co, return_margin=True, fields=fields)
    margin = np.minimum(margin, m)
    return (out, margin) if return_margin else out


def from_sdf(field, res, normalize=True):
    """vector_functions.py:130-139 + batch_normalize (vector_modification_functions.py:14-20): np.gradient with unit
    spacing on the reshaped grid, flattened to (dims, N), normalised where the norm is non-zero."""
    res = tuple(int(r) for r in res)
    g = np.asarray(np.gradient(np.asarray(field, dtype=np.float64).reshape(res)))
    vec = g.reshape(len(res), -1)
    if normalize:
        m = np.sqrt((vec * vec).sum(axis=0))
        mask = ~(m == 0)
        vec[:, mask] = vec[:, mask] / m[mask]
    return vec


def point_cloud_distance(co, points, dim=3):
    """sdf_3D.py:283-286 / sdf_2D.py:221-224 by exhaustive search (no KD-tree): the definition itself."""
    co = np.asarray(co, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64)[:dim]
    out = np.full(co.shape[1], np.inf)
    step = max(1, (1 << 24) // max(1, pts.shape[1]))
    for s in range(0, co.shape[1], step):
        q = co[:dim, s:s + step]
        d2 = ((q[:, :, None] - pts[:, None, :]) ** 2).sum(axis=0)
        out[s:s + step] = np.sqrt(d2.min(axis=1))
    return out


def point_cloud_distance_kdtree(co, points, dim=3):
    """The same field the way the reference computes it (sdf_3D.py:283-286): scipy's cKDTree, nearest neighbour.
    For clouds too large for the exhaustive definition above."""
    from scipy.spatial import cKDTree

    co = np.asarray(co, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64)[:dim]
    return cKDTree(pts.T).query(co[:dim].T)[0]
